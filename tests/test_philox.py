"""In-kernel sampler noise ("one kernel (Philox in-kernel)", replaces the torch.randn draws of
reference models/utils.py:67-99 inside the fused update).  A counter-based generator is a different random stream than
torch's by construction, so the parity chain is its own:

  CPU  oracle/philox_ref.py (numpy Philox4x32-10 + Box-Muller) against the published known-answer vectors;
  GPU  the kernel's normals against the oracle; the fused update with in-kernel noise against the SAME update kernel
       fed with the oracle's draws; exact noise properties (masked, symmetric, CoM-free); distribution moments;
       graph-replayed chain == eager chain bit for bit."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import philox_ref as P

SEED = 0x1234567890ABCDEF


def test_philox_known_answer_vectors():
    """Random123 kat_vectors, philox4x32-10."""
    kat = [([0, 0, 0, 0], [0, 0], '6627e8d5 e169c58d bc57ac4c 9b00dbd8'),
           ([0xffffffff] * 4, [0xffffffff] * 2, '408f276d 41c83b0e a20bc7c6 6d5451fd'),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0], 'd16cfe09 94fdcceb 5001e420 24126ea1')]
    for ctr, key, want in kat:
        got = P.philox4x32_10(np.array(ctr), np.array(key))
        assert ' '.join('%08x' % int(v) for v in got) == want


def test_oracle_normals_and_layout():
    z = P.normals4(np.arange(250000), 3, P.STREAM_POS, SEED).reshape(-1)
    assert abs(z.mean()) < 5e-3 and abs(z.var() - 1) < 5e-3 and abs((z ** 4).mean() - 3) < 5e-2
    # different steps / streams / seeds decorrelate
    a = P.normals4(np.arange(50000), 3, 0, SEED).reshape(-1)
    for other in (P.normals4(np.arange(50000), 4, 0, SEED), P.normals4(np.arange(50000), 3, 1, SEED),
                  P.normals4(np.arange(50000), 3, 0, SEED + 1)):
        assert abs(np.corrcoef(a, other.reshape(-1))[0, 1]) < 2e-2
    pos, feat = P.node_raw(3, 5, 6, 9, SEED)
    assert pos.shape == (3, 5, 3) and feat.shape == (3, 5, 6)
    assert np.array_equal(feat[1, 2, 4:], P.normals4(np.array([(1 * 5 + 2) * 2 + 1]), 9, P.STREAM_FEAT, SEED)[0, :2])
    e = P.edge_raw(2, 2, 4, 9, SEED)
    lin = ((1 * 2 + 1) * 4 + 3) * 4 + 2
    assert e[1, 1, 3, 2] == P.normals4(np.array([lin >> 2]), 9, P.STREAM_EDGE, SEED)[0, lin & 3]


# ---------------------------------------------------------------------------------------------------- GPU
def _normals_gpu(n4, step, stream):
    from jodo_b200 import _lib
    out = torch.empty(n4, 4, device='cuda')
    _lib.call('jodo_philox_normal', ctypes.c_ulonglong(n4), ctypes.c_ulonglong(SEED), ctypes.c_uint(step), ctypes.c_uint(stream),
              _lib.ptr(out), _lib.stream_ptr())
    torch.cuda.synchronize()
    return out


@pytest.mark.gpu
def test_kernel_normals_match_oracle():
    got = _normals_gpu(8192, 7, P.STREAM_FEAT).double().cpu().numpy()
    want = P.normals4(np.arange(8192), 7, P.STREAM_FEAT, SEED)
    err = np.abs(got - want).max()
    print(f'philox normals: max abs diff vs float64 oracle {err:.2e}')
    assert err < 2e-5                     # fp32 log / sincospi against float64
    z = _normals_gpu(1 << 20, 11, P.STREAM_EDGE).reshape(-1).double()
    assert abs(float(z.mean())) < 2e-3 and abs(float(z.var()) - 1) < 3e-3 and abs(float((z ** 4).mean()) - 3) < 2e-2


@pytest.mark.gpu
def test_update_with_inkernel_noise_equals_update_with_oracle_draws():
    from jodo_b200 import _lib
    B, N, F, ch, step = 5, 9, 9, 2, 123
    g = torch.Generator().manual_seed(4)
    n = [9, 1, 4, 7, 2]
    nm = torch.zeros(B, N, 1)
    for i, k in enumerate(n):
        nm[i, :k] = 1
    em = (nm[:, :, None, 0] * nm[:, None, :, 0] * (1 - torch.eye(N))[None]).reshape(B, N, N, 1)
    x, pred = torch.randn(B, N, F, generator=g) * nm, torch.randn(B, N, F, generator=g) * nm
    ex = torch.randn(B, N, N, ch, generator=g)
    ex = (ex + ex.transpose(1, 2)) * em
    ep = torch.randn(B, N, N, ch, generator=g)
    ep = (ep + ep.transpose(1, 2)) * em
    raw_pos, raw_feat = P.node_raw(B, N, F - 3, step, SEED)
    raw_edge = P.edge_raw(B, ch, N, step, SEED)
    c = lambda t: t.float().contiguous().cuda()
    dx, dp, dnm, dex, dep, dem = c(x), c(pred), c(nm), c(ex), c(ep), c(em)
    f = ctypes.c_float
    outs = []
    for philox in (True, False):
        o = [torch.empty_like(dx), torch.empty_like(dx), torch.empty_like(dex), torch.empty_like(dex)]
        if philox:
            _lib.call('jodo_ancestral_update_philox', _lib.ptr(dx), _lib.ptr(dp), _lib.ptr(dnm), _lib.ptr(dex), _lib.ptr(dep), _lib.ptr(dem),
                      ctypes.c_int(B), ctypes.c_int(N), ctypes.c_int(F), ctypes.c_int(ch), f(0.9), f(0.1), f(0.7), None,
                      ctypes.c_ulonglong(SEED), ctypes.c_uint(step), _lib.ptr(o[0]), _lib.ptr(o[1]), _lib.ptr(o[2]), _lib.ptr(o[3]),
                      _lib.stream_ptr())
        else:
            rp, rf, re_ = c(torch.from_numpy(raw_pos)), c(torch.from_numpy(raw_feat)), c(torch.from_numpy(raw_edge))
            _lib.call('jodo_ancestral_update', _lib.ptr(dx), _lib.ptr(dp), _lib.ptr(rp), _lib.ptr(rf), _lib.ptr(dnm), _lib.ptr(dex),
                      _lib.ptr(dep), _lib.ptr(re_), _lib.ptr(dem), ctypes.c_int(B), ctypes.c_int(N), ctypes.c_int(F), ctypes.c_int(ch),
                      f(0.9), f(0.1), f(0.7), None, _lib.ptr(o[0]), _lib.ptr(o[1]), _lib.ptr(o[2]), _lib.ptr(o[3]), _lib.stream_ptr())
        torch.cuda.synchronize()
        outs.append(o)
    a, r = outs
    assert torch.equal(a[1], r[1]) and torch.equal(a[3], r[3])                 # posterior means: no noise involved
    assert float((a[0] - r[0]).abs().max()) < 3e-5 and float((a[2] - r[2]).abs().max()) < 3e-5
    x_new, e_new = a[0], a[2]
    assert float((x_new * (1 - dnm)).abs().max()) == 0.0 and float((e_new * (1 - dem)).abs().max()) == 0.0
    assert float((e_new - e_new.permute(0, 2, 1, 3)).abs().max()) == 0.0      # symmetric by construction (same counter)
    zpos = (x_new - a[1])[..., :3] / 0.7
    assert float(zpos.sum(1).abs().max()) < 1e-5                              # CoM-free position noise
    assert float((x_new - a[1]).abs().max()) > 0.1                            # and it is noise


@pytest.mark.gpu
def test_philox_chain_graph_equals_eager_and_differs_by_seed():
    from jodo_b200 import configs, synth, sampler as S
    from jodo_b200.model import MODELS
    cfg = configs.NAMED['qm9_uncond']()
    model = MODELS[cfg.model.name](cfg).cuda().eval()
    b = synth.make_batch(cfg, 24, seed=21)
    d = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()}
    grid = torch.linspace(0.9946, 1e-3, 1000)[::100]            # 10 reverse steps
    res = []
    for graph, seed in ((False, 5), (True, 5), (False, 6)):
        smp = S.AncestralSampler(S.CosineVP(), grid, noise='philox', seed=seed)
        res.append(smp.sampling(model, d['xh'], d['node_mask'], d['edge_mask'], d['edge_x'], None, graph=graph))
    torch.cuda.synchronize()
    assert torch.isfinite(res[0][0]).all()
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])       # same counters, same kernels
    assert not torch.equal(res[0][0], res[2][0])
    with pytest.raises(ValueError):
        S.AncestralSampler(S.CosineVP(), grid, noise='philox', fused=False).step(
            lambda *a, **k: (d['xh'], d['edge_x']), 0, d['xh'], d['edge_x'], d['node_mask'], d['edge_mask'], None, None)
