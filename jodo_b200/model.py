"""Drop-in replacements for the reference's ``DGT_concat`` / ``Cond_DGT_concat`` denoisers.

Boundary: same constructor (``__init__(config)`` reading the reference's config
keys), same parameter names / shapes / registration order (reference checkpoints load with
``strict=True`` and EMA's positional ``copy_to`` works), same call
``model(t, xh, node_mask, edge_mask, context=None, edge_x=..., noise_level=..., cond_x=...,
cond_edge_x=...) -> (x_hat [B,N,3+in], e_hat [B,N,N,ch])`` as reference
models/mol_gnn.py:491-594 / 687-794.  The forward itself is a fixed sequence of launches of the
hand-written sm_100a kernels in libjodo_b200.so through its C ABI; there is no PyTorch or CPU
fallback: without the library (or without a CUDA device) the call raises.
"""
from __future__ import annotations

import ctypes

import torch
from torch import nn

from . import _lib, wide
from . import pack as _pack
from .pack import TAB_HEAD, pack_model, tab_layer_stride
from .params import build_param_tree, check_supported, dims_from_config, param_spec, synth_state_dict
from .plan import Plan

_c = ctypes.c_int
_WEIGHT_EPOCH = [0]
FINGERPRINT_EVERY = 16      # calls between two parameter fingerprints (the backstop against writes through .data)


def invalidate_packed_weights():
    """Make every jodo_b200 module re-derive its packed weight images at its next call."""
    _WEIGHT_EPOCH[0] += 1


def watch_data_writers(cls, names=('copy_to', 'restore')):
    """Wrap methods that write parameters through ``.data`` (reference models/ema.py:44-55, 66-77:
    ``ExponentialMovingAverage.copy_to`` / ``restore``) so that the packed images are invalidated after each call."""
    for name in names:
        fn = getattr(cls, name)
        if getattr(fn, '_jodo_watched', False):
            continue

        def wrapped(*a, __fn=fn, **k):
            try:
                return __fn(*a, **k)
            finally:
                invalidate_packed_weights()
        wrapped._jodo_watched = True
        wrapped.__name__, wrapped.__doc__ = getattr(fn, '__name__', name), fn.__doc__
        setattr(cls, name, wrapped)


class _Workspace:
    """Device buffers for one plan (sizes depend on the packed atom / tile counts only)."""

    def __init__(self, plan: Plan, d, meta, dev):
        f = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
        zf = lambda *s: torch.zeros(*s, device=dev, dtype=torch.float32)
        B, Nn, nt = plan.B, plan.Nn, plan.n_pair_tiles      # the edge state lives on PAIR tiles (plan.py)
        D, T = d.D, d.T
        self.feat = f(B, 64)
        self.t1 = f(B, T)
        self.temb = f(B, T)
        if d.cond_ch:
            self.c1 = f(B * d.cond_ch, D)
            self.c2 = f(B * d.cond_ch, D)
            self.ctx = f(B, T)
        self.tab = f(B, meta['ld_tab'])
        self.temb_img = torch.zeros(((B + 127) // 128) * 128 * T, device=dev, dtype=torch.float16)
        self.kin = meta['node_emb']['K']
        self.xin = f(Nn, self.kin)                          # zero-padded to the GEMM's K by jodo_gather_nodes
        self.pos = [zf(Nn, 4), zf(Nn, 4)]
        self.ah = zf(Nn, meta['ld_ah'])                    # concatenated atom hiddens (pads stay 0)
        self.h = [f(Nn, D), f(Nn, D)]
        h16 = lambda *s: torch.empty(*s, device=dev, dtype=torch.float16)
        mt = (Nn + 127) // 128
        img = lambda k: torch.zeros(mt * 128 * k, device=dev, dtype=torch.float16)    # fp16 operand image [mt][k/64][128][64]
        self.hn_img, self.h2_img, self.hnode_img, self.hout_img = img(D), img(D), img(D), img(D)
        self.ff_img = img(d.r * D)
        self.qkv = h16(3 * D // 8, Nn, 8)                   # fp16 per-atom operands gathered by the edge kernels (piece-major)
        self.hnode = zf(Nn, D)                             # atoms without partners are never written: stay 0
        self.pbuf = h16(8, Nn, 8)                           # piece-major hoisted node2edge_lin part
        self.h2 = f(Nn, D)
        self.ab = h16((4 if _pack.EQUI_LIN else 2) * D // 8, Nn, 8)    # piece-major (csrc/edge_common.cuh): A | B (| composed YA | YB)
        self.n1 = f(Nn, D)
        self.n2 = f(Nn, D // 2)
        self.ap = f(Nn, meta['npred4']['N'])
        self.eh_tile_bytes = (meta['keh'] // 64) * 128 * 128                  # fp16 image, 16 KB per 64 columns
        self.eh = torch.zeros(nt * self.eh_tile_bytes // 2, device=dev, dtype=torch.float16)
        self.e = torch.zeros(nt * 8192, device=dev, dtype=torch.float32)      # fp32 edge state, 32 KB per tile
        self.e16 = torch.zeros(nt * 8192, device=dev, dtype=torch.float16)    # fp16 operand copy, 16 KB per tile
        self.extra = torch.zeros(nt * 128, device=dev, dtype=torch.uint8)
        self.flags = torch.zeros(4, device=dev, dtype=torch.int32)            # [0] dist flag, [1] nan flag, [2] non-uniform
        self.mol_bad = torch.zeros(B, device=dev, dtype=torch.int32)          # molecules with a non-finite input (NaN isolation)


class _DGTBase(nn.Module):
    MAX_PLANS = 4       # cached (plan, workspace) entries, least recently used evicted first
    VARIANT = None      # set by the subclasses: the variant is the class's, whatever name the registry knows it under

    def __init__(self, config):
        super().__init__()
        check_supported(config, self.VARIANT)
        self.dims = dims_from_config(config, self.VARIANT)
        # nf = 256 with the 14 + 2 head layout of every reference config: fused edge-tile kernels (csrc/attn.cu and
        # equi.cu hard-code 7 learned heads of 18 channels per half, 16-column value heads, the /3 adjacency mean);
        # other sizes, other head layouts and the 2-D model: GEMM + row-kernel path (wide.py), which takes H / X / sc
        # as arguments
        d_ = self.dims
        self.wide = d_.two_d or not (d_.D == 256 and d_.H == 16 and d_.X == 2 and d_.S == 14 and d_.sc == 18 and d_.C == 16)
        if self.wide:
            why = wide.supported(self.dims)
            if why:
                raise NotImplementedError('jodo_b200: ' + why)
        else:
            if self.dims.r not in (2, 4):
                raise NotImplementedError(f'jodo_b200 edge kernels are built for model.mlp_ratio 2 or 4 (got {self.dims.r})')
            if self.dims.ce % 4 or self.dims.ce > 16:
                raise NotImplementedError('unsupported n_layers (edge hidden slice must be a multiple of 4 <= 16)')
        self.edge_th = float(config.model.edge_quan_th)
        self.spatial_cut_off = float(getattr(config.model, 'spatial_cut_off', 0.))
        self.n_layers = self.dims.L
        self._spec = param_spec(config, self.VARIANT)
        build_param_tree(self, self._spec)
        self.load_state_dict(synth_state_dict(self._spec, seed=int(getattr(config, 'seed', 0))))
        self._packed = {}
        self._packed_key = None
        self._fp_pending = None
        self._fp_tick = 0
        self.force_wide = False    # tests: run an nf = 256 model through the wide path
        self._plans = {}
        self.debug = None          # set to a dict to capture intermediates (tests)

    # ---- caches -------------------------------------------------------------------------------------
    def refresh_weights(self):
        """Drop the packed weight images; the next call re-derives them from the parameters."""
        self._packed = {}
        self._fp_pending = None

    @staticmethod
    def _fingerprint(params):
        return torch.stack(torch._foreach_norm(params))          # one 2-norm per parameter tensor, on the device

    def _weights(self, use_wide=None):
        """Packed fp16 operand images of the parameters (one set per path: fused edge-tile kernels / wide path).  They
        are re-derived when a parameter is replaced or written in place through the tensor itself (optimizer steps,
        ``load_state_dict``, ``p.copy_``: ``_version`` changes) or after ``refresh_weights()`` /
        ``invalidate_packed_weights()``.  Writes through ``p.data`` (the reference's
        ``ExponentialMovingAverage.copy_to`` / ``restore``, models/ema.py:55, 77) bypass the version counter; wrap those
        with ``watch_data_writers``.  As a backstop every FINGERPRINT_EVERY-th call enqueues a fingerprint of the
        parameters (no host sync); a later call that finds it different from the one taken at pack time raises."""
        use_wide = self.wide if use_wide is None else use_wide
        params = list(self.parameters())
        key = tuple((p.data_ptr(), p._version) for p in params) + (_WEIGHT_EPOCH[0],)
        if key != self._packed_key:
            self._packed, self._packed_key, self._fp_pending = {}, key, None
        capturing = torch.cuda.is_current_stream_capturing()      # CUDA-graph capture: no event queries, no D2H copies
        pend = None if capturing else self._fp_pending
        if pend is not None and pend.query():
            self._fp_pending = None
            if self._packed and not torch.equal(self._fp_host, self._packed_fp):
                self._packed = {}
                raise _lib.JodoError('parameters were modified through .data after the weight images were packed, so '
                                     'earlier calls used stale weights; call model.refresh_weights() after such writes '
                                     '(or wrap the writer with jodo_b200.model.watch_data_writers)')
        self._fp_tick += 1
        if use_wide not in self._packed:
            sd = {k: v for k, v in self.state_dict().items()}
            if not self._packed:
                self._packed_fp = self._fingerprint(params).cpu()
                self._fp_host = torch.empty_like(self._packed_fp).pin_memory()
                self._fp_pending = None
            self._packed[use_wide] = pack_model(sd, self.dims, params[0].device, fused=not use_wide)
        elif self._fp_pending is None and self._fp_tick % FINGERPRINT_EVERY == 0 and not capturing:
            self._fp_host.copy_(self._fingerprint(params), non_blocking=True)
            self._fp_pending = torch.cuda.Event()
            self._fp_pending.record()
        return self._packed[use_wide]

    def _compose_items(self, pk):
        """Device table of jodo_equi_compose_item, one per block (pointers into the packed buffer), built once per packing."""
        t = getattr(pk, '_compose_table', None)
        if t is None:
            import numpy as np
            d = self.dims
            stride = tab_layer_stride(d.D)
            arr = (_lib.EquiComposeItem * d.L)()
            for l in range(d.L):
                p = f'b{l}.'
                arr[l] = _lib.EquiComposeItem(pk.ptr(p + 'c0.w32'), pk.ptr(p + 'c0.b32'), pk.ptr(p + 'wi.w32'), pk.ptr(p + 'wi.b32'),
                                              pk.ptr(p + 'w2.w32'), TAB_HEAD + l * stride + 6 * d.D + 6 * d.ed, pk.ptr(p + 'wce.img'),
                                              pk.ptr(p + 'ab.img'), pk.ptr(p + 'ab.b'), pk.ptr(p + 'eqc'))
            t = torch.from_numpy(np.frombuffer(bytes(arr), dtype=np.uint8).copy()).to(pk.buf.device)
            pk._compose_table = t
        return t

    def graph_token(self, node_mask, edge_mask):
        """What a CUDA graph captured over this model's launches depends on: the (plan, workspace, masks) cache entry and
        the packed weight images.  The holder keeps the returned objects alive (raw pointers into them are baked into
        the graph) and compares `graph_token(...)` identity-wise before each replay (sampler.GraphedAncestralStep)."""
        hit = self._plan(node_mask, edge_mask)
        return hit, self._weights(hit[5])

    def _plan(self, node_mask, edge_mask):
        key = (node_mask.data_ptr(), tuple(node_mask.shape), node_mask._version, edge_mask.data_ptr(),
               tuple(edge_mask.shape), edge_mask._version, self.force_wide)
        hit = self._plans.get(key)
        if hit is not None:
            self._plans[key] = self._plans.pop(key)          # most recently used last
        if hit is None:
            plan = Plan(node_mask)
            B, N = plan.B, plan.N
            nm = (node_mask.reshape(B, N) > 0).float()
            want = nm[:, :, None] * nm[:, None, :] * (1 - torch.eye(N, device=nm.device))[None]
            if not torch.equal((edge_mask.reshape(B, N, N) > 0).float(), want):
                raise ValueError('edge_mask must be node_mask x node_mask without the diagonal '
                                 '(reference sampling.py:197-199)')
            use_wide = self.wide or self.force_wide or plan.loose
            if use_wide and not self.wide:
                why = wide.supported(self.dims)
                if why:
                    raise NotImplementedError('jodo_b200 wide path (molecules with more than 129 atoms): ' + why)
            ws = (wide.WideWorkspace if use_wide else _Workspace)(plan, self.dims, self._weights(use_wide).meta, node_mask.device)
            while len(self._plans) >= self.MAX_PLANS:        # each entry pins a workspace (GBs at B = 2500): evict the
                self._plans.pop(next(iter(self._plans)))     # least recently used one, never the whole cache
            # the masks are kept alive with the entry so that the allocator cannot hand their addresses to new masks
            hit = self._plans[key] = (plan, ws, _lib.plan_struct(plan), node_mask, edge_mask, use_wide,
                                      _lib.pair_plan_struct(plan))
        return hit

    # ---- forward ------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, t, xh, node_mask, edge_mask, context=None, *args, **kwargs):
        if self.training:
            raise RuntimeError('jodo_b200 implements the inference path (call model.eval())')
        if getattr(self, '_is_replica', False):
            # torch.nn.DataParallel over several devices (reference models/utils.py:27 with > 1 visible GPU) calls
            # per-device replicas from worker threads; replicas have no parameters() and would share this module's
            # packed images, plans and workspaces across devices
            raise _lib.JodoError('jodo_b200 modules cannot be replicated by torch.nn.DataParallel over several devices: '
                                 'run one process per GPU (torchrun; sampler.shard_molecules / gather_samples) or '
                                 'restrict the process to one device (CUDA_VISIBLE_DEVICES)')
        edge_x = kwargs['edge_x']
        noise_level = kwargs['noise_level']
        cond_x = kwargs.get('cond_x')
        cond_edge_x = kwargs.get('cond_edge_x')
        if not xh.is_cuda:
            raise _lib.JodoError('jodo_b200 runs on CUDA tensors only (no CPU fallback)')
        d = self.dims
        if d.cond_ch and context is None:
            raise ValueError('cond_DGT_concat needs a context')
        hit = self._plan(node_mask, edge_mask)
        plan, ws, ps, use_wide, pps = hit[0], hit[1], hit[2], hit[5], hit[6]
        pk = self._weights(use_wide)
        meta = pk.meta
        B, N = plan.B, plan.N
        st = _lib.stream_ptr()
        L = _lib.lib()
        dbg = self.debug
        c32 = lambda x: x.contiguous().float()
        xh, edge_x, noise_level = c32(xh), c32(edge_x), c32(noise_level)
        if cond_x is not None:
            cond_x, cond_edge_x = c32(cond_x), c32(cond_edge_x)
        if use_wide:
            return wide.forward_wide(self, pk, plan, ws, ps, pps, xh, edge_x, noise_level, cond_x, cond_edge_x, context)
        D, T, ld_tab = d.D, d.T, meta['ld_tab']

        def lin(name, A, C, M=None, **kw):
            m = meta[name]
            _lib.rowlinear(A, m['K'], pk[name + '.img'], pk[name + '.b'], C, m['N'], m['NT'], M=M, stream=st,
                           tag='jodo_rowlinear:' + name.split('.')[-1], **kw)

        def ilin(name, Aimg, tag=None, **kw):
            m = meta[name]
            _lib.imglinear(Aimg, plan.Nn, m['K'], pk[name + '.img'], pk[name + '.b'], m['N'], m['NT'], stream=st,
                           tag=tag or ('jodo_imglinear:' + name.split('.')[-1]), **kw)

        # ---- per molecule: noise-level embedding (+ context) and all AdaLN tables
        # flags[2] = 1 unless every molecule carries the same conditioning row (the samplers broadcast one noise level)
        nonuni = ws.flags.data_ptr() + 8
        _lib.call('jodo_time_features', _lib.ptr(noise_level), _lib.ptr(pk['time.w8']), _lib.ptr(ws.feat), _c(B), st)
        if d.cond_ch:
            lin('time1', ws.feat, ws.t1, epi=_lib.EPI_ACT, act_out=_lib.ACT_GELU)
            ctx = c32(context).reshape(B * d.cond_ch)
            _lib.call('jodo_cond_in', _lib.ptr(ctx), _lib.ptr(pk['cond0.w']), _lib.ptr(pk['cond0.b']), _lib.ptr(ws.c1),
                      _c(B * d.cond_ch), _c(D), st)
            lin('cond2', ws.c1, ws.c2)
            lin('condlin', ws.c2.view(B, d.cond_ch * D), ws.ctx)
            lin('time3', ws.t1, ws.temb, epi=_lib.EPI_ADD, aux=ws.ctx)
            _lib.call('jodo_uniform_flag', _lib.ptr(ws.temb), _c(B), _c(T), ctypes.c_void_p(nonuni), st)
        else:
            # No context: the rows depend on the noise level alone, so the flag is taken from the INPUT and, when it reads
            # uniform, the two layers of time_mlp run for row 0 only, as matrix-vector products on the GEMMs' weight images
            # (a 128-row tensor-core tile per layer is a 16-chunk latency chain: 0.03 + 0.05 ms); the GEMMs return at once.
            _lib.call('jodo_uniform_flag', _lib.ptr(noise_level), _c(B), _c(1), ctypes.c_void_p(nonuni), st)
            for name, src, dst, act in (('time1', ws.feat, ws.t1, _lib.ACT_GELU), ('time3', ws.t1, ws.temb, _lib.ACT_NONE)):
                m = meta[name]
                _lib.call('jodo_row0_linear', _lib.ptr(src), _c(m['K']), ctypes.c_void_p(pk.ptr(name + '.img')), _c(m['NT']),
                          _c(m['N']), _lib.ptr(pk[name + '.b']), _c(_lib.ACT_NONE), _c(act), None, _lib.ptr(dst),
                          ctypes.c_void_p(nonuni), st, tag='jodo_row0_linear:' + name)
                if act:
                    lin(name, src, dst, epi=_lib.EPI_ACT, act_out=act, skip_if_zero=nonuni)
                else:
                    lin(name, src, dst, skip_if_zero=nonuni)
        # all AdaLN rows.  Uniform conditioning: row 0 as a matrix-vector product (every consumer reads row 0); otherwise every
        # molecule's row through the persistent GEMM.  Each tests the device flag, one of them returns at once.
        m = meta['tab']
        _lib.call('jodo_row0_linear', _lib.ptr(ws.temb), _c(T), ctypes.c_void_p(pk.ptr('tab.img')), _c(m['NT']), _c(m['N']),
                  _lib.ptr(pk['tab.b']), _c(_lib.ACT_SILU), _c(_lib.ACT_NONE), None, _lib.ptr(ws.tab), ctypes.c_void_p(nonuni), st,
                  tag='jodo_row0_linear:tab')
        _lib.call('jodo_act_image', _lib.ptr(ws.temb), _c(T), _c(B), _c(T), _c(_lib.ACT_SILU), _lib.ptr(ws.temb_img), st)
        _lib.imglinear(ws.temb_img, B, m['K'], pk['tab.img'], pk['tab.b'], m['N'], m['NT'], C32=ws.tab, stream=st,
                       tag='jodo_imglinear:tab', skip_if_zero=nonuni)
        if _pack.EQUI_LIN:   # uniform conditioning: coord_mlp.0 composed into input_lin for every block from the step's table row
            _lib.call('jodo_equi_compose', ctypes.c_void_p(self._compose_items(pk).data_ptr()), _c(d.L), _lib.ptr(ws.tab),
                      ctypes.c_void_p(nonuni), st)
        # ---- per atom: packed inputs, node embedding (slice 0 of the concatenated atom hiddens)
        _lib.call('jodo_gather_nodes', _lib.ptr(xh), _lib.ptr(cond_x), ctypes.byref(ps), _c(d.inn), _c(ws.kin),
                  _lib.ptr(ws.xin), _lib.ptr(ws.pos[0]), _lib.ptr(ws.mol_bad), st)
        lin('node_emb', ws.xin, ws.ah[:, :D])
        # ---- per edge: model-level embedding + adjacency heads
        # ---- per unordered pair (the edge state is symmetric): model-level embedding + adjacency bits
        ea = _lib.EdgeEmbedArgs(pps, _lib.dp(edge_x), _lib.dp(cond_edge_x), _lib.dp(cond_x), d.ch, d.inn, self.edge_th,
                                self.spatial_cut_off, _lib.dp(ws.flags), _lib.dp(ws.tab), ld_tab, pk.ptr('gbf'),
                                pk.ptr('edge_emb.img'), pk.ptr('edge_emb.b'), _lib.dp(ws.e), _lib.dp(ws.e16),
                                _lib.dp(ws.eh), ws.eh_tile_bytes, _lib.dp(ws.extra), ws.flags.data_ptr() + 8,
                                _lib.dp(ws.mol_bad))
        _lib.call('jodo_edge_embed', ctypes.byref(ea), st)

        h = ws.ah[:, :D]
        stride = tab_layer_stride(D)
        for l in range(d.L):
            p = f'b{l}.'
            off = TAB_HEAD + l * stride
            pin, pout = ws.pos[l & 1], ws.pos[(l + 1) & 1]
            hout = ws.h[l & 1]
            # norm1_node + modulate -> fp16 operand image, q/k/v
            _lib.call('jodo_ln_mod_img', _lib.ptr(h), _c(h.stride(0)), None, _c(0), _lib.ptr(ws.tab), _c(ld_tab),
                      _c(0), _c(off), _c(off + D), ctypes.byref(ps), None, _c(0), _lib.ptr(ws.hn_img), None,
                      ctypes.c_void_p(nonuni), st)
            ilin(p + 'qkv', ws.hn_img, C16=ws.qkv)
            aa = _lib.AttnArgs(ps, _lib.dp(ws.e16), _lib.dp(pin), _lib.dp(ws.qkv), plan.Nn, _lib.dp(ws.tab), ld_tab,
                               off, _lib.dp(ws.extra), pk.ptr(p + 'emb.img'), pk.ptr(p + 'e0.img'), pk.ptr(p + 'e1.img'),
                               _lib.dp(ws.hnode), ws.flags.data_ptr() + 8, pk.host[p + 'gbf4'], pk.host[p + 'emb.b'])
            _lib.call('jodo_attn', ctypes.byref(aa), st)
            # node path: gated residual + norm2 (+ image of hnode), hoisted node2edge, FFN, hoisted input_lin parts, node_l
            _lib.call('jodo_ln_mod_img', _lib.ptr(h), _c(h.stride(0)), _lib.ptr(ws.hnode), _c(D), _lib.ptr(ws.tab),
                      _c(ld_tab), _c(off + 2 * D), _c(off + 3 * D), _c(off + 4 * D), ctypes.byref(ps), _lib.ptr(ws.h2),
                      _c(D), _lib.ptr(ws.h2_img), _lib.ptr(ws.hnode_img), ctypes.c_void_p(nonuni), st)
            ilin(p + 'n2e', ws.hnode_img, C16=ws.pbuf)
            ilin(p + 'ff1', ws.h2_img, epi=_lib.EPI_ACT, act_out=_lib.ACT_SILU, Cimg=ws.ff_img)
            ilin(p + 'ff2', ws.ff_img, epi=_lib.EPI_GATED_RES, aux=ws.h2, gate=ws.tab[:, off + 5 * D:],
                 row_mol=plan.node_mol, C32=hout, Cimg=ws.hout_img, nonuni=nonuni)
            ilin(p + 'ab', ws.hout_img, C16=ws.ab)
            ilin(p + 'node_l', ws.hout_img, C32=ws.ah[:, D + l * meta['cnp']:])
            # edge path
            ua = _lib.EdgeUpdateArgs(pps, _lib.dp(ws.e), _lib.dp(ws.e16), _lib.dp(ws.pbuf), plan.Nn,
                                     _lib.dp(ws.tab), ld_tab, off, d.r, pk.ptr(p + 'ff3.img'), pk.ptr(p + 'ff4.img'),
                                     pk.ptr(p + 'edge_l.img'), _lib.dp(ws.eh), ws.eh_tile_bytes, d.ed + l * d.ce, d.ce,
                                     ws.flags.data_ptr() + 8, pk.host[p + 'n2e.bias'], pk.host[p + 'ff3.b'],
                                     pk.host[p + 'ff4.b'], pk.host[p + 'edge_l.b'])
            _lib.call('jodo_edge_update', ctypes.byref(ua), st)
            # coordinate update (JODO_EQUI_LIN=1: the composed kernel under uniform conditioning, the general one
            # otherwise; each tests the device flag and one of them returns at once)
            if _pack.EQUI_LIN:
                la = _lib.EquiLinArgs(ps, _lib.dp(ws.e16), _lib.dp(pin), _lib.dp(pout), _lib.dp(ws.ab), plan.Nn, _lib.dp(ws.extra),
                                      pk.ptr(p + 'win.img'), pk.ptr(p + 'wce.img'), pk.ptr(p + 'eqc'), meta['coord_scale'][l],
                                      nonuni, pk.host[p + 'gbf4'])
                _lib.call('jodo_equi_lin', ctypes.byref(la), st)
            qa = _lib.EquiArgs(ps, _lib.dp(ws.e16), _lib.dp(pin), _lib.dp(pout), _lib.dp(ws.ab), plan.Nn,
                               _lib.dp(ws.tab), ld_tab, off, _lib.dp(ws.extra), pk.ptr(p + 'win.img'),
                               pk.ptr(p + 'wc0h.img'), pk.ptr(p + 'w2.img'), pk.ptr(p + 'w2x.img'), meta['coord_scale'][l],
                               nonuni, pk.host[p + 'gbf4'], pk.host[p + 'b0h'], 1 if _pack.EQUI_LIN else 0)
            _lib.call('jodo_equi', ctypes.byref(qa), st)
            _lib.call('jodo_com', _lib.ptr(pout), ctypes.byref(ps), st)
            if dbg is not None:
                dbg.setdefault('blocks', []).append(dict(hnode=ws.hnode.clone(), h=hout.clone(), e=ws.e.clone(),
                                                         pos=pout.clone(), qkv=ws.qkv.clone(), hn_img=ws.hn_img.clone()))
            h = hout
        # ---- heads
        lin('npred0', ws.ah, ws.n1, epi=_lib.EPI_ACT, act_out=_lib.ACT_SILU)
        lin('npred2', ws.n1, ws.n2, epi=_lib.EPI_ACT, act_out=_lib.ACT_SILU)
        lin('npred4', ws.n2, ws.ap)
        out_x = torch.zeros(B, N, 3 + d.inn, device=xh.device, dtype=torch.float32)
        _lib.call('jodo_node_out', _lib.ptr(ws.pos[d.L & 1]), _lib.ptr(ws.ap), _c(ws.ap.stride(0)), ctypes.byref(ps),
                  ctypes.c_void_p(ws.flags.data_ptr() + 4), _lib.ptr(ws.mol_bad), _c(d.inn), _lib.ptr(out_x), st)
        # every pair row writes both orientations: the reference's 0.5 (e + e^T) (mol_gnn.py:579) of two identical values
        out_e = torch.zeros(B, N, N, d.ch, device=xh.device, dtype=torch.float32)
        ha = _lib.EdgeHeadArgs(pps, _lib.dp(ws.eh), ws.eh_tile_bytes, meta['keh'], pk.ptr('ehead0.img'), pk.ptr('ehead0.b'),
                               pk.ptr('ehead2.img'), pk.ptr('ehead2.b'), pk.ptr('ehead4.w'), pk.ptr('ehead4.b'), d.ch,
                               _lib.dp(out_e), _lib.dp(ws.mol_bad))
        _lib.call('jodo_edge_head', ctypes.byref(ha), st)
        if dbg is not None:
            dbg.update(tab=ws.tab.clone(), temb=ws.temb.clone(), ah=ws.ah.clone(), eh=ws.eh.clone(),
                       extra=ws.extra.clone(), plan=plan, flags=ws.flags.clone())
        return out_x, out_e


class DGT_concat(_DGTBase):
    """B200-native drop-in for the reference ``DGT_concat`` (models/mol_gnn.py:410-594)."""
    VARIANT = 'uncond'


class Cond_DGT_concat(_DGTBase):
    """B200-native drop-in for the reference ``Cond_DGT_concat`` (models/mol_gnn.py:597-794)."""
    VARIANT = 'cond'


class DGT_concat_2D(_DGTBase):
    """B200-native drop-in for the reference ``DGT_concat_2D`` (models/mol_gnn.py:797-947; MOSES / ZINC250k configs):
    atom features and bonds only -- ``xh`` is ``[B, N, in]``, the return is ``(atom_pred [B,N,in], e_hat)``.  Runs
    on the wide path without the distance features and the coordinate branch, with one adjacency head."""
    VARIANT = '2d'


class DGT_concat_sim(_DGTBase):
    """B200-native drop-in for the reference ``DGT_concat_sim`` (models/mol_gnn.py:949-1124): ``EquivariantBlock`` (:97) with
    ``Trans_Layer`` attention (models/layers.py:13: all heads learned, no adjacency heads) and ``CondEquiUpdate`` (:16: one
    coord_mlp output).  Runs on the wide path with zero extra heads."""
    VARIANT = 'sim'


MODELS = {'DGT_concat': DGT_concat, 'cond_DGT_concat': Cond_DGT_concat, 'DGT_concat_2D': DGT_concat_2D,
          'DGT_concat_sim': DGT_concat_sim}


def create_model(config, device='cuda'):
    """Same role as reference models/utils.py:24-28 (without the DataParallel wrapper: multi-GPU is one
    process per GPU here)."""
    return MODELS[config.model.name](config).to(device).eval()


def register_into_reference(models_utils_module, suffix='_b200'):
    """Make ``config.model.name = 'DGT_concat_b200'`` / ``'cond_DGT_concat_b200'`` resolve to this
    implementation inside the reference's registry (reference models/utils.py:5-21)."""
    for name, cls in MODELS.items():
        key = name + suffix
        if key not in models_utils_module._MODELS:
            models_utils_module.register_model(cls, name=key)
