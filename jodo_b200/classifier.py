"""Drop-in for the reference's property classifier (cond_gen/model.py:26-220): the EGNN that scores conditional
samples at the end of ``get_cond_sampling_eval_fn`` (reference sampling.py:363-367), and for its edge-list builder
``get_adj_matrix_fn`` (cond_gen/utils.py:18-40, a Python triple loop over batch x n x n).

Boundary: same constructor arguments, same parameter names / shapes / registration order (the reference's
``best_checkpoint.npy`` state dicts load with ``strict=True``), same call
``classifier(h0=[B*N, in], x=[B*N, 3], edges=..., edge_attr=None, node_mask=[B*N, 1], edge_mask=[B*N*N, 1], n_nodes=N)
-> pred [B]``.  The forward is a fixed sequence of launches through the C ABI (jodo_rowlinear on packed atoms,
jodo_imglinear on the plan's directed edge rows, the row kernels of csrc/egnn.cu); no CPU / PyTorch fallback.
``edges`` is accepted and ignored: the graph is always the full graph of every molecule, which the varlen plan encodes
(only real ordered pairs; the reference multiplies the padded and diagonal rows by zero, model.py:207).
"""
from __future__ import annotations

import ctypes
import math
import pickle

import torch
from torch import nn

from . import _lib
from .pack import Packed, ceil_to
from .plan import Plan

_c = ctypes.c_int


def egnn_param_spec(in_node_nf, hidden_nf, n_layers, attention, node_attr, in_edge_nf=0):
    """[(name, shape)] in the reference's registration order (cond_gen/model.py:35-53, 93-122; E_GCL_mask deletes
    coord_mlp, :196)."""
    H = hidden_nf
    spec = [('embedding.weight', (H, in_node_nf)), ('embedding.bias', (H,))]
    na = in_node_nf if node_attr else 0
    for i in range(n_layers):
        p = f'gcl_{i}.'
        spec += [(p + 'edge_mlp.0.weight', (H, 2 * H + 1 + in_edge_nf)), (p + 'edge_mlp.0.bias', (H,)),
                 (p + 'edge_mlp.2.weight', (H, H)), (p + 'edge_mlp.2.bias', (H,)),
                 (p + 'node_mlp.0.weight', (H, 2 * H + na)), (p + 'node_mlp.0.bias', (H,)),
                 (p + 'node_mlp.2.weight', (H, H)), (p + 'node_mlp.2.bias', (H,))]
        if attention:
            spec += [(p + 'att_mlp.0.weight', (1, H)), (p + 'att_mlp.0.bias', (1,))]
    spec += [('node_dec.0.weight', (H, H)), ('node_dec.0.bias', (H,)), ('node_dec.2.weight', (H, H)), ('node_dec.2.bias', (H,)),
             ('graph_dec.0.weight', (H, H)), ('graph_dec.0.bias', (H,)), ('graph_dec.2.weight', (1, H)), ('graph_dec.2.bias', (1,))]
    return spec


def egnn_synth_state_dict(spec, seed=0, gain=1.0):
    """Seeded nn.Linear-style init (uniform +-1/sqrt(fan_in)); tests and benches (no checkpoints are available offline)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    fan = {}
    for name, shape in spec:
        if name.endswith('.weight'):
            fan[name[:-7]] = shape[1]
    for name, shape in spec:
        bound = gain / math.sqrt(fan[name.rsplit('.', 1)[0]])
        sd[name] = (torch.rand(shape, generator=g) * 2 - 1) * bound
    return sd


class _Named(nn.Module):
    """Container whose children are registered under numeric names (the reference's nn.Sequential indices)."""


def _build_tree(root, spec):
    for name, shape in spec:
        parts = name.split('.')
        mod = root
        for p in parts[:-1]:
            if p not in mod._modules:
                mod.add_module(p, _Named())
            mod = mod._modules[p]
        mod.register_parameter(parts[-1], nn.Parameter(torch.empty(shape), requires_grad=False))


class EGNN(nn.Module):
    """B200-native drop-in for the reference ``EGNN`` (cond_gen/model.py:26-70)."""

    def __init__(self, in_node_nf, in_edge_nf, hidden_nf, device='cuda', act_fn=None, n_layers=4, coords_weight=1.0,
                 attention=False, node_attr=1):
        super().__init__()
        if in_edge_nf:
            raise NotImplementedError('jodo_b200 EGNN: edge attributes are not used by the reference sampler (in_edge_nf = 0)')
        if act_fn is not None and not isinstance(act_fn, nn.SiLU):
            raise NotImplementedError('jodo_b200 EGNN: the activation is SiLU (cond_gen/model.py:27)')
        if hidden_nf % 64 or hidden_nf > 256:
            raise NotImplementedError(f'jodo_b200 EGNN: hidden_nf must be a multiple of 64 up to 256 (got {hidden_nf})')
        self.hidden_nf, self.n_layers, self.in_node_nf = hidden_nf, n_layers, in_node_nf
        self.attention, self.node_attr = bool(attention), bool(node_attr)
        self.device = device
        self._spec = egnn_param_spec(in_node_nf, hidden_nf, n_layers, self.attention, self.node_attr)
        _build_tree(self, self._spec)
        self.load_state_dict(egnn_synth_state_dict(self._spec, seed=0))
        self._packed, self._packed_key = None, None
        self._plans = {}
        self.to(device)

    # ---- packed weights ----------------------------------------------------------------------------------
    def _weights(self):
        params = list(self.parameters())
        key = tuple((p.data_ptr(), p._version) for p in params)
        if key == self._packed_key:
            return self._packed
        sd = dict(self.state_dict())
        H, L, inn = self.hidden_nf, self.n_layers, self.in_node_nf
        pk = Packed(params[0].device)
        nt = 128 if H % 128 == 0 else 64

        def add_lin(name, w_pieces, b_pieces, n, k, nt_=nt):
            n_pad, k_pad = ceil_to(n, nt_), ceil_to(k, 64)
            pk.image_h(name + '.img', n_pad, k_pad, nt_, w_pieces)
            pk.vec(name + '.b', n_pad, b_pieces)
            pk.meta[name] = dict(N=n_pad, K=k_pad, NT=nt_)

        W = lambda n: sd[n + '.weight']
        Bv = lambda n: sd[n + '.bias']
        add_lin('embedding', [(W('embedding'), 0, 0)], [(Bv('embedding'), 0)], H, inn)
        # node_mlp.0 reads cat[h, agg(, h0)]; the workspace row is [h | agg | h0 | 0] with kcat columns
        self.kcat = ceil_to(2 * H + (inn if self.node_attr else 0), 64)
        for l in range(L):
            p, q = f'gcl_{l}.', f'l{l}.'
            w0 = W(p + 'edge_mlp.0')                               # [H, 2H + 1]: [h_row | h_col | radial]
            add_lin(q + 'pq', [(w0[:, :H], 0, 0), (w0[:, H:2 * H], H, 0)], [(Bv(p + 'edge_mlp.0'), 0)], 2 * H, H)
            pk.vec(q + 'wr', H, [(w0[:, 2 * H].contiguous(), 0)])
            add_lin(q + 'e2', [(W(p + 'edge_mlp.2'), 0, 0)], [(Bv(p + 'edge_mlp.2'), 0)], H, H)
            add_lin(q + 'n0', [(W(p + 'node_mlp.0'), 0, 0)], [(Bv(p + 'node_mlp.0'), 0)], H, self.kcat)
            add_lin(q + 'n2', [(W(p + 'node_mlp.2'), 0, 0)], [(Bv(p + 'node_mlp.2'), 0)], H, H)
            if self.attention:
                pk.vec(q + 'wa', H, [(W(p + 'att_mlp.0'), 0)])
                pk.add_host(q + 'ba', Bv(p + 'att_mlp.0'))
        add_lin('nd0', [(W('node_dec.0'), 0, 0)], [(Bv('node_dec.0'), 0)], H, H)
        add_lin('nd2', [(W('node_dec.2'), 0, 0)], [(Bv('node_dec.2'), 0)], H, H)
        add_lin('gd0', [(W('graph_dec.0'), 0, 0)], [(Bv('graph_dec.0'), 0)], H, H)
        add_lin('gd2', [(W('graph_dec.2'), 0, 0)], [(Bv('graph_dec.2'), 0)], 1, H, nt_=16)
        pk.finish()
        self._packed, self._packed_key = pk, key
        return pk

    def _plan(self, node_mask, n_nodes):
        key = (node_mask.data_ptr(), tuple(node_mask.shape), node_mask._version, int(n_nodes))
        hit = self._plans.get(key)
        if hit is None:
            plan = Plan(node_mask.reshape(-1, n_nodes), loose=True)     # row kernels only: groups may straddle tiles
            if len(self._plans) >= 4:
                self._plans.pop(next(iter(self._plans)))
            hit = self._plans[key] = (plan, _lib.plan_struct(plan), node_mask)
        return hit

    # ---- forward -----------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, h0, x, edges=None, edge_attr=None, node_mask=None, edge_mask=None, n_nodes=None):
        if self.training:
            raise RuntimeError('jodo_b200 EGNN implements the inference path (call classifier.eval())')
        if not h0.is_cuda:
            raise _lib.JodoError('jodo_b200 runs on CUDA tensors only (no CPU fallback)')
        if edge_attr is not None:
            raise NotImplementedError('jodo_b200 EGNN: edge_attr must be None (reference sampling.py:366)')
        H, L, inn, N = self.hidden_nf, self.n_layers, self.in_node_nf, int(n_nodes)
        B = h0.shape[0] // N
        plan, ps, _ = self._plan(node_mask, N)
        if edge_mask is not None:
            nm = (node_mask.reshape(B, N) > 0).float()
            want = nm[:, :, None] * nm[:, None, :] * (1 - torch.eye(N, device=nm.device))[None]
            if not torch.equal((edge_mask.reshape(B, N, N) > 0).float(), want):
                raise ValueError('edge_mask must be node_mask x node_mask without the diagonal (reference sampling.py:333-337)')
        pk = self._weights()
        meta = pk.meta
        dev = h0.device
        st = _lib.stream_ptr()
        P = _lib.ptr
        Nn, R = plan.Nn, plan.n_tiles * 128
        f = lambda *s: torch.zeros(*s, device=dev, dtype=torch.float32)
        idx = plan.node_dense.long()
        kin = meta['embedding']['K']
        xin = f(Nn, kin)
        xin[:, :inn] = h0.reshape(B * N, inn).float()[idx]
        pos = f(Nn, 4)
        pos[:, :3] = x.reshape(B * N, 3).float()[idx]
        kcat = self.kcat
        cat = f(Nn, kcat)                                          # [h | agg | h0 | 0]
        if self.node_attr:
            cat[:, 2 * H:2 * H + inn] = xin[:, :inn]
        pq, t1 = f(Nn, 2 * H), f(Nn, H)
        a_img = torch.zeros(R * H, device=dev, dtype=torch.float16)
        m16 = torch.zeros(R, H, device=dev, dtype=torch.float16)

        def lin(name, A, C, **kw):
            m = meta[name]
            _lib.rowlinear(A, m['K'], pk[name + '.img'], pk[name + '.b'], C, m['N'], m['NT'], stream=st,
                           tag='jodo_rowlinear:' + name.split('.')[-1], **kw)

        h = cat[:, :H]
        lin('embedding', xin, h)
        for l in range(L):
            q = f'l{l}.'
            lin(q + 'pq', h, pq)
            _lib.call('jodo_egnn_edge_in', ctypes.byref(ps), P(pos), P(pq), _c(2 * H), _c(H), P(pk[q + 'wr']), P(a_img), st)
            m = meta[q + 'e2']
            _lib.imglinear(a_img, R, m['K'], pk[q + 'e2.img'], pk[q + 'e2.b'], m['N'], m['NT'], epi=_lib.EPI_ACT,
                           act_out=_lib.ACT_SILU, C16=m16, stream=st, tag='jodo_imglinear:egnn_e2')
            wa = pk[q + 'wa'] if self.attention else None
            ba = float(pk.host[q + 'ba'][0]) if self.attention else 0.0
            _lib.call('jodo_egnn_agg', P(plan.grp_row0), P(plan.grp_len), P(m16), _c(H), _c(H), P(wa), ctypes.c_float(ba),
                      ctypes.c_void_p(cat.data_ptr() + 4 * H), _c(kcat), _c(Nn), st)
            lin(q + 'n0', cat, t1, epi=_lib.EPI_ACT, act_out=_lib.ACT_SILU)
            lin(q + 'n2', t1, h, epi=_lib.EPI_ADD, aux=h)          # recurrent: h = h + node_mlp(..) (model.py:146-147)
        t2, g0, g1 = f(Nn, H), f(B, H), f(B, H)
        lin('nd0', h, t1, epi=_lib.EPI_ACT, act_out=_lib.ACT_SILU)
        lin('nd2', t1, t2)
        _lib.call('jodo_mol_sum', P(t2), _c(H), _c(H), P(plan.mol_start), _c(B), P(g0), _c(H), st)
        lin('gd0', g0, g1, epi=_lib.EPI_ACT, act_out=_lib.ACT_SILU)
        out = f(B, 16)
        lin('gd2', g1, out)
        return out[:, 0].contiguous()


def get_model(args):
    """Reference cond_gen/model.py:6-13."""
    if args.model_name == 'egnn':
        return EGNN(in_node_nf=5, in_edge_nf=0, hidden_nf=args.nf, device=args.device, n_layers=args.n_layers,
                    coords_weight=1.0, attention=args.attention, node_attr=args.node_attr)
    raise Exception('Wrong model name %s' % args.model_name)


def get_classifier(classifier_path, args_classifier_path, device='cuda'):
    """Reference cond_gen/model.py:15-23: the pickled training arguments select the sizes, the state dict loads strictly."""
    with open(args_classifier_path, 'rb') as fh:
        args_classifier = pickle.load(fh)
    args_classifier.device = device
    args_classifier.model_name = 'egnn'
    classifier = get_model(args_classifier)
    classifier.load_state_dict(torch.load(classifier_path, map_location=torch.device('cpu')))
    return classifier


def get_adj_matrix_fn():
    """Reference cond_gen/utils.py:18-40 without the Python triple loop: rows[k] = i + b n, cols[k] = j + b n for every
    (b, i, j) in lexicographic order (self-loops included), built with three aranges and cached per (n, batch, device)."""
    cache = {}

    def get_adj_matrix(n_nodes, batch_size, device):
        key = (int(n_nodes), int(batch_size), str(device))
        if key not in cache:
            b = torch.arange(batch_size, device=device).view(-1, 1, 1) * n_nodes
            i = torch.arange(n_nodes, device=device).view(1, -1, 1)
            j = torch.arange(n_nodes, device=device).view(1, 1, -1)
            shape = (batch_size, n_nodes, n_nodes)
            cache[key] = [(i + b).expand(shape).reshape(-1), (j + b).expand(shape).reshape(-1)]
        return cache[key]

    return get_adj_matrix
