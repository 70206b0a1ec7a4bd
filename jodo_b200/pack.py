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


def tab_layer_stride(D):
    return 6 * D + 6 * (D // 4) + 2 * D + 16


class Packed:
    """Flat fp32 buffer with named, 256-byte aligned pieces."""

    def __init__(self, device):
        self.device = device
        self._pieces = []
        self._off = {}
        self._size = 0
        self.buf = None
        self.meta = {}
        self.host = {}          # small per-column constant tables passed to kernels BY VALUE (ctypes float arrays)

    def add(self, name, t):
        t = t.detach().to(self.device, torch.float32).reshape(-1)
        self._off[name] = (self._size, t.numel())
        self._pieces.append(t)
        pad = (-t.numel()) % 64
        if pad:
            self._pieces.append(torch.zeros(pad, device=self.device))
        self._size += t.numel() + pad

    def add_host(self, name, t):
        import ctypes
        v = t.detach().float().reshape(-1).cpu().tolist()
        self.host[name] = (ctypes.c_float * len(v))(*v)

    def finish(self):
        self.buf = torch.cat(self._pieces)
        self._pieces = None
        return self

    def __getitem__(self, name):
        o, n = self._off[name]
        return self.buf[o:o + n]

    def ptr(self, name):
        return self.buf.data_ptr() + 4 * self._off[name][0]


def split_heads(w, D, qk):
    """Rows [qk, ...] of lin_query / lin_key / lin_edge0 -> [D, ...]: heads 0..S/2-1 at rows [0, qk/2), heads
    S/2..S-1 at rows [D/2, D/2 + qk/2), zeros elsewhere, so that each half of the attention CTA (csrc/attn.cu) owns
    one 128-column half of the q / k / g0 rows."""
    h = qk // 2
    out = torch.zeros((D,) + tuple(w.shape[1:]), dtype=w.dtype, device=w.device)
    out[:h] = w[:h]
    out[D // 2:D // 2 + h] = w[h:]
    return out


def _gbf_consts(sd, prefix, dev):
    """{mu, sqrt(0.5 log2 e)/sg, 1/(a sg)} x 64 from the reference's fp32 mu / sg (models/layers.py:291-295,332-333):
    exp(-0.5 ((x-mu)/sg)^2) / (a sg) = 2^(-((x-mu) c1)^2) * c2."""
    mu = sd[prefix + '.means.weight'].float().view(-1)
    sg = sd[prefix + '.stds.weight'].float().view(-1).abs() + 1e-5
    a = (2 * 3.14159) ** 0.5
    asg = a * sg
    out = torch.zeros(192, device=dev)
    k = mu.numel()
    out[0:k] = mu
    out[64:64 + k] = (0.5 * 1.4426950408889634) ** 0.5 / sg
    out[128:128 + k] = 1.0 / asg
    return out


def _gbf_table4(sd, prefix, dev):
    """float4 {mu, sqrt(0.5 log2 e)/sg, 1/(a sg), 0} per feature COLUMN c = k + 1 (entry 0, the raw x column, is
    unused): the layout gbf_eval_cols (csrc/edge_common.cuh) reads with one 16-byte shared-memory load per feature."""
    c = _gbf_consts(sd, prefix, dev)
    out = torch.zeros(64, 4, device=dev)
    out[1:, 0] = c[0:63]
    out[1:, 1] = c[64:127]
    out[1:, 2] = c[128:191]
    return out.reshape(-1)


def pack_model(sd, dims, device, fused=None):
    """sd: name -> tensor (reference names, no 'module.' prefix); dims: jodo_b200.params.Dims."""
    d = dims
    D, ed, T, L = d.D, d.ed, d.T, d.L
    if fused is None:
        fused = D == 256                   # fused edge-tile kernels; other sizes take the wide path (jodo_b200/wide.py)
    assert not fused or D == 256, 'the fused edge-tile kernels are built for nf = 256'
    ntb = 256 if D % 256 == 0 else 128     # N tile of the wide per-molecule / per-atom GEMMs
    sd = {k: v.detach().to(device, torch.float32) for k, v in sd.items()}
    pk = Packed(device)
    W = lambda n: sd[n + '.weight']
    Bv = lambda n: sd[n + '.bias']
    z = lambda *s: torch.zeros(*s, device=device)

    def add_lin(name, w, b, nt, n_pad=None, k_pad=None):
        n, k = w.shape
        n_pad = ceil_to(n, nt) if n_pad is None else n_pad
        k_pad = ceil_to(k, 64) if k_pad is None else k_pad
        pk.add(name + '.img', weight_image_h(pad2(w, n_pad, k_pad), nt))
        bb = z(n_pad)
        if b is not None:
            bb[:n] = b
        pk.add(name + '.b', bb)
        pk.meta[name] = dict(N=n_pad, K=k_pad, NT=nt)

    # ---- molecule level
    pk.add('time.w8', sd['time_mlp.0.weights'])
    add_lin('time1', W('time_mlp.1'), Bv('time_mlp.1'), ntb)
    add_lin('time3', W('time_mlp.3'), Bv('time_mlp.3'), ntb)
    if d.cond_ch:
        pk.add('cond0.w', W('cond_mlp.0').reshape(-1))
        pk.add('cond0.b', Bv('cond_mlp.0'))
        add_lin('cond2', W('cond_mlp.2'), Bv('cond_mlp.2'), ntb)
        add_lin('condlin', W('cond_lin'), Bv('cond_lin'), ntb)
    # per-molecule tables: one GEMM  [B, T] x [T, ld_tab]
    stride = tab_layer_stride(D)
    ld_tab = ceil_to(TAB_HEAD + L * stride, 256)
    wt, bt = z(ld_tab, T), z(ld_tab)
    # Every "scale" column gets +1 on its bias: the kernels modulate with one FMA, x * (1 + scale) + shift.
    if not d.two_d:
        wt[0:2], bt[0:2] = W('dist_layer.time_mlp.1'), Bv('dist_layer.time_mlp.1')
        bt[0] += 1.0                                          # GBF time MLP chunks as (scale, shift)
    for l in range(L):
        b = f'e_block_{l}'
        o = TAB_HEAD + l * stride
        chunks = [(f'{b}.node_time_mlp.1', 6 * D, ((D, 2 * D), (4 * D, 5 * D))),
                  (f'{b}.edge_time_mlp.1', 6 * ed, ((ed, 2 * ed), (4 * ed, 5 * ed)))]
        if not d.two_d:                                       # the 2-D model has no coordinate branch / distance features
            chunks += [(f'{b}.equi_update.time_mlp.1', 2 * D, ((D, 2 * D),)),
                       (f'{b}.dist_layer.time_mlp.1', 2, ((0, 1),))]
        for name, n, scales in chunks:
            wt[o:o + n], bt[o:o + n] = W(name), Bv(name)
            for s0, s1 in scales:                             # chunk order: shift, scale, gate (AdaLN); scale, shift (GBF)
                bt[o + s0:o + s1] += 1.0
            o += n
    add_lin('tab', wt, bt, 256)                              # ld_tab is a multiple of 256
    pk.meta['ld_tab'] = ld_tab
    # ---- atom level
    add_lin('node_emb', W('node_emb'), Bv('node_emb'), ntb)
    cnp = ceil_to(d.cn, 4)
    k_ah = ceil_to(D + L * cnp, 64)
    pk.meta.update(cnp=cnp, k_ah=k_ah, ld_ah=k_ah + (64 if fused else 128))    # room for the last node_i GEMM's padded N tile
    w0 = W('node_pred_mlp.0')
    w0p = z(D, k_ah)
    w0p[:, :D] = w0[:, :D]
    for l in range(L):
        w0p[:, D + l * cnp:D + l * cnp + d.cn] = w0[:, D + l * d.cn:D + (l + 1) * d.cn]
    add_lin('npred0', w0p, Bv('node_pred_mlp.0'), ntb)
    add_lin('npred2', W('node_pred_mlp.2'), Bv('node_pred_mlp.2'), 128 if (D // 2) % 128 == 0 else 64)
    add_lin('npred4', W('node_pred_mlp.4'), Bv('node_pred_mlp.4'), 16)
    if not fused:
        from .wide import pack_wide
        pack_wide(pk, sd, d, add_lin)
        return pk.finish()
    # ---- edge level (model)
    pk.add('gbf', _gbf_consts(sd, 'dist_layer', device))
    we = W('edge_emb')                                         # [ed, 2ch + ed]: [edge_x | cond_edge_x | dist]
    wep = z(ed, 96)
    wep[:, :ed] = we[:, 2 * d.ch:]
    wep[:, ed:ed + 2 * d.ch] = we[:, :2 * d.ch]
    pk.add('edge_emb.img', weight_image(wep, ed))
    pk.add('edge_emb.b', Bv('edge_emb'))
    keh = ceil_to(ed + L * d.ce, 64)
    assert keh == 192, keh
    pk.meta['keh'] = keh
    wh0 = z(2 * ed, keh)
    wh0[:ed, :d.edge_cat] = W('edge_exist_mlp.0')
    wh0[ed:, :d.edge_cat] = W('edge_type_mlp.0')
    pk.add('ehead0.img', weight_image_h(wh0, 2 * ed))
    pk.add('ehead0.b', torch.cat([Bv('edge_exist_mlp.0'), Bv('edge_type_mlp.0')]))
    wh2 = z(ed, 2 * ed)
    wh2[:ed // 2, :ed] = W('edge_exist_mlp.2')
    wh2[ed // 2:, ed:] = W('edge_type_mlp.2')
    pk.add('ehead2.img', weight_image_h(wh2, ed))
    pk.add('ehead2.b', torch.cat([Bv('edge_exist_mlp.2'), Bv('edge_type_mlp.2')]))
    pk.add('ehead4.w', torch.cat([W('edge_exist_mlp.4'), W('edge_type_mlp.4')], dim=0))     # [ch, 32]
    pk.add('ehead4.b', torch.cat([Bv('edge_exist_mlp.4'), Bv('edge_type_mlp.4')]))
    # ---- blocks
    scales = []
    for l in range(L):
        b = f'e_block_{l}'
        p = f'b{l}.'
        wq = z(3 * D, D)
        bq = z(3 * D)
        wq[:D], bq[:D] = (split_heads(W(f'{b}.attn_mpnn.lin_query'), D, d.qk),
                          split_heads(Bv(f'{b}.attn_mpnn.lin_query'), D, d.qk))
        wq[D:2 * D], bq[D:2 * D] = (split_heads(W(f'{b}.attn_mpnn.lin_key'), D, d.qk),
                                    split_heads(Bv(f'{b}.attn_mpnn.lin_key'), D, d.qk))
        wq[2 * D:], bq[2 * D:] = W(f'{b}.attn_mpnn.lin_value'), Bv(f'{b}.attn_mpnn.lin_value')
        add_lin(p + 'qkv', wq, bq, 256)
        add_lin(p + 'n2e', W(f'{b}.node2edge_lin'), None, 64)
        pk.add_host(p + 'n2e.bias', Bv(f'{b}.node2edge_lin'))
        add_lin(p + 'ff1', W(f'{b}.ff_linear1'), Bv(f'{b}.ff_linear1'), 256)
        add_lin(p + 'ff2', W(f'{b}.ff_linear2'), Bv(f'{b}.ff_linear2'), 128)     # K = r D is deep: narrower tiles balance the SMs
        wi = W(f'{b}.equi_update.input_lin')                   # [D, 2D + 2ed]: [h_row | h_col | e | dist]
        add_lin(p + 'ab', torch.cat([wi[:, :D], wi[:, D:2 * D]], dim=0),
                torch.cat([Bv(f'{b}.equi_update.input_lin'), z(D)]), 256)      # input_lin bias rides on the h[row] part
        add_lin(p + 'node_l', W(f'node_{l}'), Bv(f'node_{l}'), 64, n_pad=64)
        pk.add(p + 'gbf', _gbf_consts(sd, f'{b}.dist_layer', device))
        pk.add(p + 'emb.img', weight_image_h(W(f'{b}.edge_emb'), ed))                       # [64, 128]: [dist | e]
        pk.add_host(p + 'emb.b', Bv(f'{b}.edge_emb'))
        pk.add(p + 'e0.img', weight_image_h(split_heads(W(f'{b}.attn_mpnn.lin_edge0'), D, d.qk), D))
        pk.add(p + 'e1.img', weight_image_h(W(f'{b}.attn_mpnn.lin_edge1'), D))
        w3, w4 = W(f'{b}.ff_linear3'), W(f'{b}.ff_linear4')    # [ed r, ed], [ed, ed r]
        # SiLU(x) = h + h tanh(h), h = x / 2: the factor (exact in fp16) is folded into ff_linear3's image and bias
        pk.add(p + 'ff3.img', weight_image_h(0.5 * w3, 128))                               # N tiles of 128 hidden units
        b3 = z(256)
        b3[:ed * d.r] = 0.5 * Bv(f'{b}.ff_linear3')
        pk.add_host(p + 'ff3.b', b3)
        pk.add(p + 'ff4.img', weight_image_h(w4, ed))
        pk.add_host(p + 'ff4.b', Bv(f'{b}.ff_linear4'))
        pk.add(p + 'edge_l.img', weight_image_h(pad2(W(f'edge_{l}'), 16, ed), 16))
        bl = z(16)
        bl[:d.ce] = Bv(f'edge_{l}')
        pk.add_host(p + 'edge_l.b', bl)
        pk.add(p + 'win.img', weight_image_h(wi[:, 2 * D:].contiguous(), D))                # [256, 128]: [e | dist]
        pk.add(p + 'win.b', Bv(f'{b}.equi_update.input_lin'))
        pk.add(p + 'wc0.img', weight_image_h(W(f'{b}.equi_update.coord_mlp.0'), D))
        pk.add(p + 'wc0.b', Bv(f'{b}.equi_update.coord_mlp.0'))
        pk.add(p + 'wc2', W(f'{b}.equi_update.coord_mlp.2'))                               # [3, 256]
        # SiLU(x) = h + h tanh(h) with h = x / 2: the factor 1/2 is exact in fp16, so it is folded into the image and bias
        pk.add(p + 'wc0h.img', weight_image_h(0.5 * W(f'{b}.equi_update.coord_mlp.0'), D))
        pk.add_host(p + 'b0h', 0.5 * Bv(f'{b}.equi_update.coord_mlp.0'))
        pk.add(p + 'w2.img', weight_image_h(pad2(W(f'{b}.equi_update.coord_mlp.2'), 16, D), 16))          # N = 16 (3 real)
        pk.add(p + 'w2x.img', weight_image_h(pad2(W(f'{b}.equi_update.coord_mlp.2'), 32, D), 32))         # N = 32 (CTA-pair kernel)
        pk.add_host(p + 'gbf4', _gbf_table4(sd, f'{b}.dist_layer', device))
        scales.append(sd[f'{b}.equi_update.coord_norm.scale'].reshape(()))
    pk.meta['coord_scale'] = [float(s) for s in torch.stack(scales).cpu()]
    return pk.finish()
