"""Weight packing: reference parameter tensors -> operand images the kernels consume.

An operand image is the K-major SWIZZLE_128B layout described in csrc/common.cuh: 32 fp32 columns
(128 bytes) per row per chunk, the 16-byte piece p of row r stored at slot p ^ (r & 7)."""
from __future__ import annotations

import torch


def round_tf32(w: torch.Tensor) -> torch.Tensor:
    """Round fp32 to the nearest tf32 (10 explicit mantissa bits), ties away from zero like cvt.rna."""
    bits = w.contiguous().view(torch.int32)
    bits = (bits + 0x1000) & ~0x1FFF
    return bits.view(torch.float32)


def pad2(w: torch.Tensor, n: int, k: int) -> torch.Tensor:
    out = torch.zeros((n, k), dtype=torch.float32, device=w.device)
    out[:w.shape[0], :w.shape[1]] = w
    return out


def ceil_to(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def weight_image(w: torch.Tensor, nt: int | None = None, tf32: bool = True) -> torch.Tensor:
    """W [N, K] (N % nt == 0, K % 32 == 0) -> flat image [N/nt][K/32][nt][8 slots][4]."""
    n, k = w.shape
    nt = n if nt is None else nt
    assert n % nt == 0 and k % 32 == 0 and nt % 8 == 0, (n, k, nt)
    w = w.to(torch.float32)
    if tf32:
        w = round_tf32(w)
    x = w.reshape(n // nt, nt, k // 32, 8, 4).permute(0, 2, 1, 3, 4)          # [tile, chunk, row, piece, 4]
    rows = torch.arange(nt, device=w.device)
    slots = torch.arange(8, device=w.device)
    src_piece = slots[None, :] ^ (rows[:, None] & 7)                           # piece stored in (row, slot)
    idx = src_piece[None, None, :, :, None].expand(n // nt, k // 32, nt, 8, 4)
    return torch.gather(x, 3, idx).contiguous().reshape(-1)


def weight_image_h(w: torch.Tensor, nt: int | None = None) -> torch.Tensor:
    """fp16 operand image of W [N, K] (N % nt == 0, K % 64 == 0): [N/nt][K/64][nt][8 slots][8 halves], returned as a
    flat float32-typed tensor holding the fp16 bit patterns (so that it can live in the packed fp32 buffer)."""
    n, k = w.shape
    nt = n if nt is None else nt
    assert n % nt == 0 and k % 64 == 0 and nt % 8 == 0, (n, k, nt)
    h = w.to(torch.float32).clamp(-65504.0, 65504.0).to(torch.float16)
    x = h.reshape(n // nt, nt, k // 64, 8, 8).permute(0, 2, 1, 3, 4)          # [tile, chunk, row, piece, 8]
    rows = torch.arange(nt, device=w.device)
    slots = torch.arange(8, device=w.device)
    src_piece = slots[None, :] ^ (rows[:, None] & 7)
    idx = src_piece[None, None, :, :, None].expand(n // nt, k // 64, nt, 8, 8)
    return torch.gather(x, 3, idx).contiguous().reshape(-1).view(torch.float32)


def image_to_matrix_h(img: torch.Tensor, rows: int, k: int) -> torch.Tensor:
    """Inverse for ONE fp16 tile image (flat fp16 tensor [k/64][rows][8 slots][8]) -> fp32 [rows, k]."""
    x = img.reshape(k // 64, rows, 8, 8)
    r = torch.arange(rows, device=img.device)
    slots = torch.arange(8, device=img.device)
    slot_of_piece = slots[None, :] ^ (r[:, None] & 7)
    idx = slot_of_piece[None, :, :, None].expand(k // 64, rows, 8, 8)
    y = torch.gather(x, 2, idx)
    return y.permute(1, 0, 2, 3).reshape(rows, k).float()


def image_to_matrix(img: torch.Tensor, rows: int, k: int) -> torch.Tensor:
    """Inverse of the activation/edge image layout for ONE tile: [k/32][rows][8 slots][4] -> [rows, k]."""
    x = img.reshape(k // 32, rows, 8, 4)
    r = torch.arange(rows, device=img.device)
    slots = torch.arange(8, device=img.device)
    slot_of_piece = slots[None, :] ^ (r[:, None] & 7)                          # slot holding piece p of row r
    idx = slot_of_piece[None, :, :, None].expand(k // 32, rows, 8, 4)
    y = torch.gather(x, 2, idx)                                                # [chunk, row, piece, 4]
    return y.permute(1, 0, 2, 3).reshape(rows, k)


def matrix_to_image(m: torch.Tensor) -> torch.Tensor:
    """[rows, k] -> one tile image [k/32][rows][8 slots][4] (no tf32 rounding)."""
    rows, k = m.shape
    return weight_image(m, rows, tf32=False)


# =====================================================================================================
# Whole-model packing: reference state dict -> one flat fp32 device buffer + named offsets
# =====================================================================================================
TAB_HEAD = 16
# A/B switch: the coordinate kernel with coord_mlp.0 composed into input_lin under uniform conditioning (csrc/equi_lin.cu).
# Correct (parity-green), measured SLOWER than csrc/equi.cu on B200 (0.40 vs 0.37 ms per launch at QM9 B = 2500: it trades the
# K = 256 tensor-core product, which was never the bound, for a second 1 KB-per-edge gather, which is); off by default.
import os as _os
EQUI_LIN = _os.environ.get('JODO_EQUI_LIN') == '1'


def tab_layer_stride(D):
    return 6 * D + 6 * (D // 4) + 2 * D + 16


class Packed:
    """One flat fp32 device buffer with named, 256-byte aligned pieces, filled by ONE launch of jodo_pack_weights.

    The packing is *recorded*: every piece is declared with its destination format and the source sub-matrices that
    feed it (``dst = src * scale + add`` at a (row, column) offset; padding stays zero).  ``finish()`` zero-fills the
    buffer, uploads the item table and runs the kernel (CUDA tensors) or a torch emulation of the same items (CPU
    tensors: tests).  Pieces flagged ``host=True`` are also returned as ctypes float arrays (per-column constants the
    kernels take BY VALUE); they come back in one device-to-host copy."""

    F32, IMG_F16, IMG_TF32 = 0, 1, 2

    def __init__(self, device):
        self.device = torch.device(device)
        self._off = {}
        self._size = 0
        self._items = []        # (src2d, name, kind, dst_ld, nt, k_pad, row0, col0, scale, add)
        self._keep = []         # temporaries that must outlive the launch
        self._host_names = []
        self.buf = None
        self.meta = {}
        self.host = {}          # small per-column constant tables passed to kernels BY VALUE (ctypes float arrays)
        self.launches = 0       # kernels launched by finish() (0 on the CPU)

    # ---- declarations ----------------------------------------------------------------------------------
    def _alloc(self, name, n_floats):
        assert name not in self._off, name
        self._off[name] = (self._size, n_floats)
        self._size += n_floats + ((-n_floats) % 64)

    def _src(self, t):
        t = t.detach()
        if t.dtype != torch.float32 or t.device != self.device:
            t = t.to(self.device, torch.float32)
            self._keep.append(t)
        if t.dim() == 1:
            t = t.unsqueeze(0)
        assert t.dim() == 2 and (t.stride(1) == 1 or t.shape[1] == 1), (tuple(t.shape), t.stride())
        return t

    def _record(self, name, kind, dst_ld, nt, k_pad, pieces):
        for pc in pieces:
            src, row0, col0 = pc[0], pc[1], pc[2]
            scale = pc[3] if len(pc) > 3 else 1.0
            add = pc[4] if len(pc) > 4 else 0.0
            if src is None or src.numel() == 0:
                continue
            self._items.append((self._src(src), name, kind, dst_ld, nt, k_pad, int(row0), int(col0), float(scale), float(add)))

    def image_h(self, name, n_pad, k_pad, nt, pieces):
        """fp16 operand image of a zero-padded [n_pad, k_pad] matrix in N tiles of nt rows; pieces: (W2d, row0, col0[, scale])."""
        assert n_pad % nt == 0 and k_pad % 64 == 0 and nt % 8 == 0, (name, n_pad, k_pad, nt)
        for pc in pieces:
            assert pc[0].dim() == 2 and pc[1] + pc[0].shape[0] <= n_pad and pc[2] + pc[0].shape[1] <= k_pad, (name, tuple(pc[0].shape))
        self._alloc(name, n_pad * k_pad // 2)
        self._record(name, self.IMG_F16, 0, nt, k_pad, pieces)

    def image_tf32(self, name, n_pad, k_pad, nt, pieces):
        assert n_pad % nt == 0 and k_pad % 32 == 0 and nt % 8 == 0, (name, n_pad, k_pad, nt)
        self._alloc(name, n_pad * k_pad)
        self._record(name, self.IMG_TF32, 0, nt, k_pad, pieces)

    def mat(self, name, rows, ld, pieces, host=False):
        """fp32 row-major [rows, ld]; pieces: (src, row0, col0[, scale[, add]]) -- 1-D sources are one row."""
        self._alloc(name, rows * ld)
        self._record(name, self.F32, ld, 0, 0, pieces)
        if host:
            self._host_names.append(name)

    def vec(self, name, n, pieces, host=False):
        """fp32 vector of n entries; pieces: (src (flattened), offset[, scale[, add]])."""
        self.mat(name, 1, n, [(pc[0].reshape(-1), 0, pc[1]) + tuple(pc[2:]) for pc in pieces], host=host)

    def add(self, name, t):
        """A tensor copied as is (flattened)."""
        self.vec(name, t.numel(), [(t, 0)])

    def add_host(self, name, t, scale=1.0, n=None):
        """A small table that kernels take by value: packed like `add` and fetched back to the host with the others."""
        self.vec(name, t.numel() if n is None else n, [(t, 0, scale)], host=True)

    # ---- execution -----------------------------------------------------------------------------------
    def finish(self):
        self.buf = torch.zeros(max(self._size, 64), device=self.device, dtype=torch.float32)
        if self.device.type == 'cuda':
            self._run_cuda()
        else:
            self._run_emulated()
        if self._host_names:                              # one device-to-host copy for all by-value tables
            lo = min(self._off[n][0] for n in self._host_names)
            hi = max(self._off[n][0] + self._off[n][1] for n in self._host_names)
            hb = self.buf[lo:hi].cpu()
            import ctypes
            for n in self._host_names:
                o, k = self._off[n]
                v = hb[o - lo:o - lo + k].tolist()
                self.host[n] = (ctypes.c_float * k)(*v)
        self._items, self._keep = None, None
        return self

    def _dst_ptr(self, name):
        return self.buf.data_ptr() + 4 * self._off[name][0]

    def _run_cuda(self):
        import ctypes
        import numpy as np
        from . import _lib
        n = len(self._items)
        if n == 0:
            return
        arr = (_lib.PackItem * n)()
        nblk = np.zeros(n, dtype=np.int64)
        for i, (src, name, kind, dst_ld, nt, k_pad, row0, col0, scale, add) in enumerate(self._items):
            rows, cols = src.shape
            ld = src.stride(0) if rows > 1 else max(cols, 1)
            if src.shape[1] == 1 and src.stride(1) != 1:
                ld = src.stride(0)
            arr[i] = _lib.PackItem(src.data_ptr(), ld, rows, cols, self._dst_ptr(name), kind, dst_ld, nt, k_pad, row0, col0, scale, add)
            nblk[i] = (rows * cols + _lib.PACK_ELEMS_PER_BLOCK - 1) // _lib.PACK_ELEMS_PER_BLOCK
        first = np.zeros(n, dtype=np.int64)
        np.cumsum(nblk[:-1], out=first[1:])
        blk_item = np.repeat(np.arange(n, dtype=np.int32), nblk)
        raw = np.frombuffer(bytes(arr), dtype=np.uint8)
        table = torch.from_numpy(raw.copy()).to(self.device)
        bi = torch.from_numpy(blk_item).to(self.device)
        bf = torch.from_numpy(first.astype(np.int32)).to(self.device)
        _lib.call('jodo_pack_weights', ctypes.c_void_p(table.data_ptr()), _lib.ptr(bi), _lib.ptr(bf), ctypes.c_int(int(blk_item.shape[0])),
                  _lib.stream_ptr())
        self._keep += [table, bi, bf]
        torch.cuda.current_stream().synchronize()         # the sources may be temporaries; packing is not on the hot path
        self.launches = 2                                 # the zero fill and the pack kernel

    def _run_emulated(self):
        """The same items in torch (CPU tests; also the executable definition of the three destination formats)."""
        for src, name, kind, dst_ld, nt, k_pad, row0, col0, scale, add in self._items:
            o, nfl = self._off[name]
            rows, cols = src.shape
            v = src.float() * scale + add
            R = (row0 + torch.arange(rows))[:, None].expand(rows, cols)
            C = (col0 + torch.arange(cols))[None, :].expand(rows, cols)
            if kind == self.F32:
                self.buf[o:o + nfl][(R * dst_ld + C).reshape(-1)] = v.reshape(-1)
                continue
            tile, rr = R // nt, R % nt
            if kind == self.IMG_F16:
                off = ((tile * (k_pad // 64) + C // 64) * nt + rr) * 64 + ((((C % 64) // 8) ^ (rr % 8)) * 8) + C % 8      # in halves
                h = self.buf[o:o + nfl].view(torch.float16)
                h[off.reshape(-1)] = v.clamp(-65504.0, 65504.0).to(torch.float16).reshape(-1)
            else:
                off = ((tile * (k_pad // 32) + C // 32) * nt + rr) * 32 + ((((C % 32) // 4) ^ (rr % 8)) * 4) + C % 4      # in floats
                self.buf[o:o + nfl][off.reshape(-1)] = round_tf32(v.contiguous()).reshape(-1)

    # ---- access --------------------------------------------------------------------------------------
    def __getitem__(self, name):
        o, n = self._off[name]
        return self.buf[o:o + n]

    def ptr(self, name):
        return self.buf.data_ptr() + 4 * self._off[name][0]


def _gbf_all(sd, prefixes, dev):
    """GBF constants of several CondGaussianLayers at once (reference models/layers.py:291-295, 332-333):
    exp(-0.5 ((x-mu)/sg)^2) / (a sg) = 2^(-((x-mu) c1)^2) * c2 with c1 = sqrt(0.5 log2 e) / sg, c2 = 1 / (a sg), a = sqrt(2 * 3.14159),
    sg = |stds| + 1e-5.  Returns (mu, c1, c2), each [len(prefixes), ed - 1] -- a handful of launches for the whole model."""
    mu = torch.stack([sd[p + '.means.weight'].detach().to(dev, torch.float32).view(-1) for p in prefixes])
    sg = torch.stack([sd[p + '.stds.weight'].detach().to(dev, torch.float32).view(-1) for p in prefixes]).abs() + 1e-5
    a = (2 * 3.14159) ** 0.5
    return mu, (0.5 * 1.4426950408889634) ** 0.5 / sg, 1.0 / (a * sg)


def pack_model(sd, dims, device, fused=None):
    """sd: name -> tensor (reference names, no 'module.' prefix); dims: jodo_b200.params.Dims."""
    d = dims
    D, ed, T, L = d.D, d.ed, d.T, d.L
    if fused is None:
        fused = D == 256                   # fused edge-tile kernels; other sizes take the wide path (jodo_b200/wide.py)
    assert not fused or D == 256, 'the fused edge-tile kernels are built for nf = 256'
    ntb = 256 if D % 256 == 0 else 128     # N tile of the wide per-molecule / per-atom GEMMs
    pk = Packed(device)
    dev = pk.device
    W = lambda n: sd[n + '.weight']
    Bv = lambda n: sd[n + '.bias']

    def add_lin(name, w_pieces, b_pieces, nt, n, k, n_pad=None, k_pad=None):
        """Linear layer as an fp16 image + fp32 bias.  w_pieces: (W2d, row0, col0[, scale]); b_pieces: (b, offset[, scale[, add]])."""
        n_pad = ceil_to(n, nt) if n_pad is None else n_pad
        k_pad = ceil_to(k, 64) if k_pad is None else k_pad
        pk.image_h(name + '.img', n_pad, k_pad, nt, w_pieces)
        pk.vec(name + '.b', n_pad, b_pieces)
        pk.meta[name] = dict(N=n_pad, K=k_pad, NT=nt)

    def lin(name, wname, nt, n_pad=None, k_pad=None, bias=True):
        w = W(wname)
        add_lin(name, [(w, 0, 0)], [(Bv(wname), 0)] if bias else [], nt, w.shape[0], w.shape[1], n_pad, k_pad)

    # ---- molecule level
    pk.add('time.w8', sd['time_mlp.0.weights'])
    lin('time1', 'time_mlp.1', ntb)
    lin('time3', 'time_mlp.3', ntb)
    if d.cond_ch:
        pk.add('cond0.w', W('cond_mlp.0'))
        pk.add('cond0.b', Bv('cond_mlp.0'))
        lin('cond2', 'cond_mlp.2', ntb)
        lin('condlin', 'cond_lin', ntb)
    # per-molecule tables: one GEMM  [B, T] x [T, ld_tab]
    stride = tab_layer_stride(D)
    ld_tab = ceil_to(TAB_HEAD + L * stride, 256)
    wp, bp = [], []
    # Every "scale" column gets +1 on its bias: the kernels modulate with one FMA, x * (1 + scale) + shift.

    def tab_rows(name, o, n, scales):
        wp.append((W(name), o, 0))
        b = Bv(name)
        cuts = sorted({0, n} | {x for s0, s1 in scales for x in (s0, s1)})
        for c0, c1 in zip(cuts[:-1], cuts[1:]):
            plus = any(s0 <= c0 < s1 for s0, s1 in scales)
            bp.append((b[c0:c1], o + c0, 1.0, 1.0 if plus else 0.0))

    if not d.two_d:
        tab_rows('dist_layer.time_mlp.1', 0, 2, ((0, 1),))          # GBF time MLP chunks as (scale, shift)
    for l in range(L):
        b = f'e_block_{l}'
        o = TAB_HEAD + l * stride
        chunks = [(f'{b}.node_time_mlp.1', 6 * D, ((D, 2 * D), (4 * D, 5 * D))),
                  (f'{b}.edge_time_mlp.1', 6 * ed, ((ed, 2 * ed), (4 * ed, 5 * ed)))]
        if not d.two_d:                                       # the 2-D model has no coordinate branch / distance features
            chunks += [(f'{b}.equi_update.time_mlp.1', 2 * D, ((D, 2 * D),)),
                       (f'{b}.dist_layer.time_mlp.1', 2, ((0, 1),))]
        for name, n, scales in chunks:                        # chunk order: shift, scale, gate (AdaLN); scale, shift (GBF)
            tab_rows(name, o, n, scales)
            o += n
    add_lin('tab', wp, bp, 128, ld_tab, T)                    # ld_tab is a multiple of 256; 128-column tiles: 154 CTAs for the one-row-tile launch
    pk.meta['ld_tab'] = ld_tab
    # ---- atom level
    lin('node_emb', 'node_emb', ntb)
    cnp = ceil_to(d.cn, 4)
    k_ah = ceil_to(D + L * cnp, 64)
    pk.meta.update(cnp=cnp, k_ah=k_ah, ld_ah=k_ah + (64 if fused else 128))    # room for the last node_i GEMM's padded N tile
    w0 = W('node_pred_mlp.0')
    add_lin('npred0', [(w0[:, :D], 0, 0)] + [(w0[:, D + l * d.cn:D + (l + 1) * d.cn], 0, D + l * cnp) for l in range(L)],
            [(Bv('node_pred_mlp.0'), 0)], ntb, D, k_ah)
    lin('npred2', 'node_pred_mlp.2', 128 if (D // 2) % 128 == 0 else 64)
    lin('npred4', 'node_pred_mlp.4', 16)
    if not fused:
        from .wide import pack_wide
        pack_wide(pk, sd, d, add_lin, lin)
        pk.finish()
        pk.meta['coord_scale'] = [float(v) for v in pk.host['coord_scale']] if 'coord_scale' in pk.host else []
        return pk
    # ---- edge level (model)
    gb = [f'e_block_{l}.dist_layer' for l in range(L)]
    mu, c1, c2 = _gbf_all(sd, ['dist_layer'] + gb, dev)
    pk._keep += [mu, c1, c2]
    k63 = mu.shape[1]
    pk.mat('gbf', 3, 64, [(mu[0], 0, 0), (c1[0], 1, 0), (c2[0], 2, 0)])       # {mu, c1, c2} x 64 (entry 63 unused)
    we = W('edge_emb')                                         # [ed, 2ch + ed]: [edge_x | cond_edge_x | dist]
    pk.image_tf32('edge_emb.img', ed, 96, ed, [(we[:, 2 * d.ch:], 0, 0), (we[:, :2 * d.ch], 0, ed)])
    pk.add('edge_emb.b', Bv('edge_emb'))
    keh = ceil_to(ed + L * d.ce, 64)
    assert keh == 192, keh
    pk.meta['keh'] = keh
    pk.image_h('ehead0.img', 2 * ed, keh, 2 * ed, [(W('edge_exist_mlp.0'), 0, 0), (W('edge_type_mlp.0'), ed, 0)])
    pk.vec('ehead0.b', 2 * ed, [(Bv('edge_exist_mlp.0'), 0), (Bv('edge_type_mlp.0'), ed)])
    pk.image_h('ehead2.img', ed, 2 * ed, ed, [(W('edge_exist_mlp.2'), 0, 0), (W('edge_type_mlp.2'), ed // 2, ed)])
    pk.vec('ehead2.b', ed, [(Bv('edge_exist_mlp.2'), 0), (Bv('edge_type_mlp.2'), ed // 2)])
    pk.mat('ehead4.w', d.ch, ed // 2, [(W('edge_exist_mlp.4'), 0, 0), (W('edge_type_mlp.4'), 1, 0)])     # [ch, 32]
    pk.vec('ehead4.b', d.ch, [(Bv('edge_exist_mlp.4'), 0), (Bv('edge_type_mlp.4'), 1)])
    # ---- blocks
    h = d.qk // 2

    def split(w):
        """Rows of lin_query / lin_key / lin_edge0 in the split-head layout of csrc/attn.cu: heads 0..S/2-1 at rows
        [0, qk/2), heads S/2..S-1 at rows [D/2, D/2 + qk/2) of a D-row block, zeros elsewhere."""
        return [(w[:h], 0), (w[h:], D // 2)]

    for l in range(L):
        b = f'e_block_{l}'
        p = f'b{l}.'
        wq, wk, wv = W(f'{b}.attn_mpnn.lin_query'), W(f'{b}.attn_mpnn.lin_key'), W(f'{b}.attn_mpnn.lin_value')
        bq, bk, bv = Bv(f'{b}.attn_mpnn.lin_query'), Bv(f'{b}.attn_mpnn.lin_key'), Bv(f'{b}.attn_mpnn.lin_value')
        add_lin(p + 'qkv',
                [(t, r0, 0) for t, r0 in split(wq)] + [(t, D + r0, 0) for t, r0 in split(wk)] + [(wv, 2 * D, 0)],
                [(t, r0) for t, r0 in split(bq)] + [(t, D + r0) for t, r0 in split(bk)] + [(bv, 2 * D)], 256, 3 * D, D)
        lin(p + 'n2e', f'{b}.node2edge_lin', 64, bias=False)
        pk.add_host(p + 'n2e.bias', Bv(f'{b}.node2edge_lin'))
        lin(p + 'ff1', f'{b}.ff_linear1', 256)
        lin(p + 'ff2', f'{b}.ff_linear2', 128)                 # K = r D is deep: narrower tiles balance the SMs
        wi, bi = W(f'{b}.equi_update.input_lin'), Bv(f'{b}.equi_update.input_lin')      # [D, 2D + 2ed]: [h_row | h_col | e | dist]
        # the bias rides on the h[row] part
        if EQUI_LIN:
            # N tiles 2, 3 (and bias [2D, 3D)) hold the same parts composed with coord_mlp.0 for the uniform-conditioning
            # coordinate kernel; jodo_equi_compose rewrites them at every call (csrc/equi_lin.cu)
            add_lin(p + 'ab', [(wi[:, :D], 0, 0), (wi[:, D:2 * D], D, 0)], [(bi, 0)], 256, 2 * D, D, n_pad=4 * D)
            pk.mat(p + 'c0.w32', D, D, [(W(f'{b}.equi_update.coord_mlp.0'), 0, 0)])
            pk.add(p + 'c0.b32', Bv(f'{b}.equi_update.coord_mlp.0'))
            pk.mat(p + 'wi.w32', D, 2 * D + 2 * ed, [(wi, 0, 0)])
            pk.add(p + 'wi.b32', bi)
            pk.mat(p + 'w2.w32', 3, D, [(W(f'{b}.equi_update.coord_mlp.2'), 0, 0)])
            pk.image_h(p + 'wce.img', D, 2 * ed, D, [])            # composed per call
            pk.vec(p + 'eqc', 1040, [])                            # d / 2 | coord_mlp.2 rows | GBF pair of the step
        else:
            add_lin(p + 'ab', [(wi[:, :D], 0, 0), (wi[:, D:2 * D], D, 0)], [(bi, 0)], 256, 2 * D, D)
        lin(p + 'node_l', f'node_{l}', 64, n_pad=64)
        pk.mat(p + 'gbf', 3, 64, [(mu[1 + l], 0, 0), (c1[1 + l], 1, 0), (c2[1 + l], 2, 0)])
        pk.image_h(p + 'emb.img', ed, 2 * ed, ed, [(W(f'{b}.edge_emb'), 0, 0)])                    # [64, 128]: [dist | e]
        pk.add_host(p + 'emb.b', Bv(f'{b}.edge_emb'))
        w0e = W(f'{b}.attn_mpnn.lin_edge0')
        pk.image_h(p + 'e0.img', D, ed, D, [(t, r0, 0) for t, r0 in split(w0e)])
        pk.image_h(p + 'e1.img', D, ed, D, [(W(f'{b}.attn_mpnn.lin_edge1'), 0, 0)])
        # SiLU(x) = h + h tanh(h), h = x / 2: the factor (exact in fp16) is folded into ff_linear3's image and bias
        pk.image_h(p + 'ff3.img', ed * d.r, ed, 128, [(W(f'{b}.ff_linear3'), 0, 0, 0.5)])           # N tiles of 128 hidden units
        pk.add_host(p + 'ff3.b', Bv(f'{b}.ff_linear3'), scale=0.5, n=256)
        pk.image_h(p + 'ff4.img', ed, ed * d.r, ed, [(W(f'{b}.ff_linear4'), 0, 0)])
        pk.add_host(p + 'ff4.b', Bv(f'{b}.ff_linear4'))
        pk.image_h(p + 'edge_l.img', 16, ed, 16, [(W(f'edge_{l}'), 0, 0)])
        pk.add_host(p + 'edge_l.b', Bv(f'edge_{l}'), n=16)
        pk.image_h(p + 'win.img', D, 2 * ed, D, [(wi[:, 2 * D:], 0, 0)])                            # [256, 128]: [e | dist]
        # SiLU(x) = h + h tanh(h) with h = x / 2: the factor 1/2 is exact in fp16, so it is folded into the image and bias
        pk.image_h(p + 'wc0h.img', D, D, D, [(W(f'{b}.equi_update.coord_mlp.0'), 0, 0, 0.5)])
        pk.add_host(p + 'b0h', Bv(f'{b}.equi_update.coord_mlp.0'), scale=0.5)
        w2 = W(f'{b}.equi_update.coord_mlp.2')
        pk.image_h(p + 'w2.img', 16, D, 16, [(w2, 0, 0)])                                            # N = 16 (3 real)
        pk.image_h(p + 'w2x.img', 32, D, 32, [(w2, 0, 0)])                                           # N = 32 (CTA-pair kernel)
        # float4 {mu, c1, c2, 0} per feature COLUMN c = k + 1 (entry 0, the raw x column, is unused): by-value table
        pk.mat(p + 'gbf4', 64, 4, [(mu[1 + l].unsqueeze(1), 1, 0), (c1[1 + l].unsqueeze(1), 1, 1), (c2[1 + l].unsqueeze(1), 1, 2)],
               host=True)
    cs = torch.stack([sd[f'e_block_{l}.equi_update.coord_norm.scale'].detach().to(dev, torch.float32).reshape(()) for l in range(L)])
    pk.add_host('coord_scale', cs)
    pk.finish()
    pk.meta['coord_scale'] = [float(v) for v in pk.host['coord_scale']]
    return pk
