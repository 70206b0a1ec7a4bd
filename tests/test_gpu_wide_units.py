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


# ---- fused edge FFN (csrc/wide_ffn.cu): every compiled size against fp64 on the fp16-rounded operands
def _ffn_reference(e, P, pi, pj, mol, nb, tab, offs, w3, b3, w4, b4, ed):
    """reference models/mol_gnn.py:304-305, 313-317 in fp64; the two GEMM operands (e2, the SiLU output, both weights) are
    rounded to fp16 as the kernel stores them."""
    h = lambda x: x.half().double()
    og, osh, osc, og2 = offs
    valid = pi >= 0
    i, j = pi.clamp(min=0).long(), pj.clamp(min=0).long()
    t = tab.double()[mol.long()]
    v = e.double()[:, :ed] + t[:, og:og + ed] * (P.double()[i, :ed] + P.double()[j, :ed] + nb.double())
    mu = v.mean(1, keepdim=True)
    var = (v * v).mean(1, keepdim=True) - mu * mu
    e2 = (v - mu) / torch.sqrt(var.clamp(min=0) + 1e-6) * t[:, osc:osc + ed] + t[:, osh:osh + ed]      # the table stores 1 + scale
    hid = h(e2) @ h(w3).t() + b3.double()
    act = hid * torch.sigmoid(hid)
    y = h(act) @ h(w4).t() + b4.double()
    out = e2 + t[:, og2:og2 + ed] * y
    out[~valid] = 0
    return out


@pytest.mark.parametrize('ed,r,tiles', [(32, 2, 3), (64, 2, 5), (96, 2, 4), (32, 4, 2), (64, 4, 7), (96, 4, 1), (96, 4, 13), (96, 4, 700)])
def test_fused_edge_ffn_matches_fp64(ed, r, tiles):
    from jodo_b200.pack import weight_image_h
    g = torch.Generator(device='cuda').manual_seed(ed * 100 + r * 10 + tiles)
    rn = lambda *s: torch.randn(*s, device='cuda', generator=g)
    M, H, Nn, B, EDP = tiles * 128, r * ed, 300, 7, 128
    e32 = torch.zeros(M, EDP, device='cuda')
    e32[:, :ed] = rn(M, ed)
    P = torch.zeros(Nn, EDP, device='cuda')
    P[:, :ed] = rn(Nn, ed)
    pi = torch.randint(0, Nn, (M,), device='cuda', generator=g, dtype=torch.int32)
    pj = torch.randint(0, Nn, (M,), device='cuda', generator=g, dtype=torch.int32)
    pi[torch.rand(M, device='cuda', generator=g) < 0.05] = -1                       # padding rows anywhere
    pi[-17:] = -1
    mol = torch.randint(0, B, (M,), device='cuda', generator=g, dtype=torch.int32).sort().values.int()
    nb = rn(ed)
    ld_tab = 16 + 6 * ed
    tab = rn(B, ld_tab)
    tab[:, 16 + 2 * ed:16 + 3 * ed] += 1.0                                          # (1 + scale) column
    offs = (16, 16 + ed, 16 + 2 * ed, 16 + 3 * ed)                                  # gate, shift, scale, gate2
    w3, b3 = rn(H, ed) / ed ** 0.5, rn(H)
    w4, b4 = rn(ed, H) / H ** 0.5, rn(ed)
    k3 = (ed + 63) // 64 * 64
    w3p = torch.zeros(H, k3, device='cuda')
    w3p[:, :ed] = 0.5 * w3
    w3i, w4i = weight_image_h(w3p, H), weight_image_h(w4, ed)
    b3h = (0.5 * b3).contiguous()
    k1, c1, k2, c2 = 192, 0, 384, 128
    img1 = torch.full((tiles * 128 * k1,), 7.0, device='cuda', dtype=torch.float16)
    img2 = torch.full((tiles * 128 * k2,), 7.0, device='cuda', dtype=torch.float16)
    want = _ffn_reference(e32, P, pi, pj, mol, nb, tab, offs, w3, b3, w4, b4, ed)
    e_in = e32.clone()
    dp = _lib.dp
    a = _lib.WideFfnArgs(M, ed, H, dp(e32), EDP, dp(P), EDP, dp(pi), dp(pj), dp(mol), dp(nb), dp(tab), ld_tab, offs[0], offs[1], offs[2],
                         offs[3], dp(w3i), dp(b3h), dp(w4i), dp(b4), dp(img1), k1, c1, dp(img2), k2, c2)
    _lib.call('jodo_wide_edge_ffn', ctypes.byref(a), _lib.stream_ptr())
    torch.cuda.synchronize()
    scale = float(want.abs().max())
    err = float((e32[:, :ed].double() - want).abs().max()) / scale
    print(f'ed={ed} r={r} tiles={tiles}: {err:.2e} of max |e_out| {scale:.2f}')
    assert err < 2e-3                                                               # two fp16-operand GEMMs, fp32 accumulation
    assert bool((e32[:, ed:] == e_in[:, ed:]).all())                                # the padding columns are not touched
    from test_gpu_imglinear import image_rows
    for img, k, c0 in ((img1, k1, c1), (img2, k2, c2)):
        rows = image_rows(img, k)
        assert float((rows[:, c0:c0 + ed].double() - e32[:, :ed].double()).abs().max()) <= 1e-3 * max(1.0, scale)     # fp16 copy of the new state
        keep = torch.ones(k, dtype=torch.bool)
        keep[c0:c0 + ed] = False
        assert bool((rows[:, keep] == 7.0).all())                                   # nothing outside the placed columns
