"""GPU unit tests of the wide path's attention kernels through the C ABI: the molecule-staged kernel (k | v of a molecule
in shared memory, online softmax; csrc/wide_attn.cu) against the per-target kernel (csrc/wide.cu) and against a plain
fp64 restatement of TransMixLayer.message + aggregation (reference models/layers.py:157-186) on the same fp16 operands."""
import ctypes

import pytest
import torch

from jodo_b200 import _lib
from jodo_b200.plan import Plan

pytestmark = pytest.mark.gpu


def _case(n_list, D, H, X, seed, loose=False):
    g = torch.Generator().manual_seed(seed)
    B, N = len(n_list), max(n_list)
    nm = torch.zeros(B, N)
    for i, n in enumerate(n_list):
        nm[i, :n] = 1
    plan = Plan(nm.cuda(), loose=loose)
    S = H - X
    sc = D // S
    qk = S * sc
    qkp = (qk + 127) // 128 * 128
    ldq, ldg = 2 * qkp + D, qkp + D
    Nn, RP = plan.Nn, plan.n_pair_tiles * 128
    qkv = torch.zeros(Nn, ldq)
    qkv[:, :qk] = torch.randn(Nn, qk, generator=g)
    qkv[:, qkp:qkp + qk] = torch.randn(Nn, qk, generator=g)
    qkv[:, 2 * qkp:] = torch.randn(Nn, D, generator=g)
    G = torch.zeros(RP, ldg)
    G[:, :qk] = torch.tanh(torch.randn(RP, qk, generator=g))
    G[:, qkp:] = torch.tanh(torch.randn(RP, D, generator=g))
    extra = torch.randint(0, 1 << max(X, 1), (RP,), generator=g, dtype=torch.uint8)
    if X:
        extra[::7] = 0                                      # rows without any adjacency (quirk 3: uniform attention if all are)
    return plan, dict(D=D, H=H, X=X, sc=sc, qk=qk, qkp=qkp, ldq=ldq, ldg=ldg), qkv.half().cuda(), G.half().cuda(), extra.cuda()


def _run(plan, d, qkv, G, extra, mol):
    hnode = torch.full((plan.Nn, d['D']), float('nan'), device='cuda')
    a = _lib.WideAttnArgs(plan.Nn, d['D'], d['H'], d['X'], d['sc'], _lib.dp(plan.grp_row0), _lib.dp(plan.grp_len), _lib.dp(plan.row_j),
                          _lib.dp(qkv), d['ldq'], d['qkp'], 2 * d['qkp'], _lib.dp(G), d['ldg'], d['qkp'], _lib.dp(extra),
                          _lib.dp(plan.row_pair), _lib.dp(hnode), max(plan.max_group, 1),
                          _lib.dp(plan.mol_start) if mol else 0, plan.B if mol else 0, int(plan.n_nodes.max()) if mol else 0)
    _lib.call('jodo_wide_attn', ctypes.byref(a), _lib.stream_ptr())
    torch.cuda.synchronize()
    return hnode


def _reference(plan, d, qkv, G, extra):
    """fp64 on the fp16 operands: logits q[c] k[r] g0 / sqrt(C), adjacency heads first, PyG softmax, v[r] g1 alpha summed."""
    D, H, X, sc, qk, qkp = d['D'], d['H'], d['X'], d['sc'], d['qk'], d['qkp']
    S, C = H - X, D // H
    q, k, v = qkv[:, :qk].double().cpu(), qkv[:, qkp:qkp + qk].double().cpu(), qkv[:, 2 * qkp:].double().cpu()
    g0, g1 = G[:, :qk].double().cpu(), G[:, qkp:].double().cpu()
    ex = extra.cpu()
    out = torch.zeros(plan.Nn, D, dtype=torch.float64)
    r0, gl, rj, rp = plan.grp_row0.cpu(), plan.grp_len.cpu(), plan.row_j.cpu(), plan.row_pair.cpu()
    for t in range(plan.Nn):
        n = int(gl[t])
        if n == 0:
            continue
        rows = torch.arange(int(r0[t]), int(r0[t]) + n)
        src, pr = rj[rows].long(), rp[rows].long()
        a = (q[t][None] * k[src] * g0[pr]).reshape(n, S, sc).sum(-1) / C ** 0.5
        adj = torch.stack([torch.where(((ex[pr] >> x) & 1).bool(), torch.tensor(1.0, dtype=torch.float64), torch.tensor(-1e10, dtype=torch.float64))
                           for x in range(X)], dim=1) if X else torch.zeros(n, 0, dtype=torch.float64)
        lg = torch.cat([adj, a], dim=1)
        e = torch.exp(lg - lg.max(0, keepdim=True).values)
        alpha = e / (e.sum(0, keepdim=True) + 1e-16)
        out[t] = ((v[src] * g1[pr]).reshape(n, H, C) * alpha[:, :, None]).sum(0).reshape(D)
    return out


@pytest.mark.parametrize('n_list,D,H,X,loose', [([9, 3, 14, 1, 12, 29, 2], 384, 16, 2, False), ([40, 33, 80, 5], 384, 16, 2, False),
                                                 ([18, 7, 29], 256, 16, 2, False), ([12, 20, 5], 256, 16, 1, False),
                                                 ([10, 16], 256, 16, 0, False), ([150, 9], 256, 16, 2, True)])
def test_molecule_staged_attention_matches_per_target_kernel_and_fp64(n_list, D, H, X, loose):
    plan, d, qkv, G, extra = _case(n_list, D, H, X, seed=sum(n_list) + D, loose=loose)
    want = _reference(plan, d, qkv, G, extra)
    per_target = _run(plan, d, qkv, G, extra, mol=False)
    staged = _run(plan, d, qkv, G, extra, mol=True)
    scale = float(want.abs().max())
    e1 = float((per_target.double().cpu() - want).abs().max()) / scale
    e2 = float((staged.double().cpu() - want).abs().max()) / scale
    print(f'n={n_list} D={D} X={X}: per-target {e1:.2e}  molecule-staged {e2:.2e} (of max |hnode| {scale:.2f})')
    assert e1 < 2e-5 and e2 < 2e-5                      # fp32 math on identical fp16 operands
    assert bool(torch.isfinite(staged).all())
