"""Observable quirks of the reference forward that had no GPU test in round 1, against the oracle:

* quirk 2 -- batch-global NaN guard (reference models/mol_gnn.py:587-589): a NaN in ONE molecule's positions zeroes
  ALL positions of the batch (then CoM), the atom / edge logits of the other molecules are untouched;
* quirk 1 fallback -- `cond_x` given but all conditioning distances zero (`distances.sum() == 0`,
  mol_gnn.py:544-545) selects zero distance features, not GBF(0);
* the fused kernels are never reached with a head layout they are not built for (nf = 256, n_heads = 8 -> wide path);
* a second CUDA stream driving the same model concurrently gives the same results (constant-table guard, kernels.h).
"""
import pytest
import torch

from helpers import golden_weights, load_golden, oracle_forward
from jodo_b200.model import MODELS
from jodo_b200.params import param_spec, synth_state_dict

pytestmark = pytest.mark.gpu
TOL = 2e-3


def _model(cfg, sd):
    m = MODELS[cfg.model.name](cfg)
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval()


def _call(model, inp):
    d = lambda v: v.cuda() if torch.is_tensor(v) else v
    return model(d(inp['t']), d(inp['xh']), d(inp['node_mask']), d(inp['edge_mask']), context=d(inp['context']),
                 edge_x=d(inp['edge_x']), noise_level=d(inp['noise_level']), cond_x=d(inp['cond_x']),
                 cond_edge_x=d(inp['cond_edge_x']))


def _rel(a, b):
    return float((a.double().cpu() - b).abs().max() / b.abs().max().clamp_min(1e-30))


def test_nan_guard_is_batch_global():
    g, cfg = load_golden('qm9_selfcond')
    sd = golden_weights(g, cfg)
    model = _model(cfg, sd)
    inp = dict(g['inputs'])
    x_clean, e_clean = _call(model, inp)
    bad = dict(inp)
    bad['xh'] = inp['xh'].clone()
    bad['xh'][2, 1, 0] = float('nan')                      # one coordinate of molecule 2
    ox, oe = oracle_forward(sd, cfg, bad, torch.float64)
    x, e = _call(model, bad)
    x, e = x.cpu(), e.cpu()
    # every position of the batch is zero (reference: pos = zeros_like(pos), then CoM of zeros)
    assert float(x[..., :3].abs().max()) == 0.0 and float(ox[..., :3].abs().max()) == 0.0
    # molecules that never saw the NaN keep their logits (the oracle agrees, and so does the clean run)
    keep = [b for b in range(x.shape[0]) if b != 2]
    for b in keep:
        assert _rel(x[b, :, 3:], ox[b, :, 3:]) < TOL
        assert _rel(e[b], oe[b]) < TOL
        assert torch.equal(x[b, :, 3:], x_clean.cpu()[b, :, 3:])
        assert torch.equal(e[b], e_clean.cpu()[b])
    # the poisoned molecule's features are NaN in the reference as well (NaN propagates through its distances)
    assert bool(torch.isnan(ox[2, :, 3:]).any()) == bool(torch.isnan(x[2, :, 3:]).any())


def test_all_conditioning_distances_zero_takes_the_zero_branch():
    """cond_x given with coincident conditioning positions: distances.sum() == 0 -> dist feature := zeros."""
    g, cfg = load_golden('qm9_selfcond')
    sd = golden_weights(g, cfg)
    model = _model(cfg, sd)
    inp = dict(g['inputs'])
    cx = inp['cond_x'].clone()
    cx[..., :3] = 0.0                                       # every molecule's conditioning positions coincide
    inp['cond_x'] = cx
    ox, oe = oracle_forward(sd, cfg, inp, torch.float64)
    x, e = _call(model, inp)
    # an out-of-distribution input (every pair spatially adjacent, no distance features): measured 1.9e-3 .. 2.1e-3 of
    # max |x|, at the edge of the usual 2e-3 and moving with the summation order -- this branch is held to 3e-3
    assert _rel(x, ox) < 3e-3 and _rel(e, oe) < TOL
    # and it is NOT what GBF(0) would give: moving one atom by a hair leaves the zero branch
    cx2 = cx.clone()
    cx2[0, 0, 0] = 1e-3
    inp2 = dict(inp)
    inp2['cond_x'] = cx2
    ox2, oe2 = oracle_forward(sd, cfg, inp2, torch.float64)
    x2, e2 = _call(model, inp2)
    assert _rel(x2, ox2) < TOL and _rel(e2, oe2) < TOL
    assert float((ox2 - ox).abs().max()) > 1e-3             # the two branches are observably different
    assert float((x2.cpu() - x.cpu()).abs().max()) > 1e-3


def test_nf256_with_other_head_layout_runs_on_the_wide_path():
    """nf = 256 with n_heads = 8 (6 learned heads of 42 channels, 32-column value heads) must not reach
    the fused kernels, which hard-code the 14 + 2 layout."""
    g, cfg = load_golden('qm9_selfcond')
    cfg.model.n_heads = 8
    sd = synth_state_dict(param_spec(cfg), seed=3, perturb=True)
    model = _model(cfg, sd)
    assert model.wide
    ox, oe = oracle_forward(sd, cfg, g['inputs'], torch.float64)
    x, e = _call(model, g['inputs'])
    assert _rel(x, ox) < 5e-3 and _rel(e, oe) < 5e-3


def test_two_streams_drive_the_model_concurrently():
    """The uniform-conditioning rows are per-device __constant__ tables; launches from different streams are ordered
    by the guard in kernels.h, so two streams with DIFFERENT noise levels must each get their own result."""
    g, cfg = load_golden('qm9_selfcond')
    sd = golden_weights(g, cfg)
    ma, mb = _model(cfg, sd), _model(cfg, sd)
    inp = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in g['inputs'].items()}
    nla = torch.full_like(inp['noise_level'], -2.0)
    nlb = torch.full_like(inp['noise_level'], 3.0)
    kw = dict(edge_x=inp['edge_x'], cond_x=inp['cond_x'], cond_edge_x=inp['cond_edge_x'])
    ref_a = ma(inp['t'], inp['xh'], inp['node_mask'], inp['edge_mask'], noise_level=nla, **kw)
    ref_b = mb(inp['t'], inp['xh'], inp['node_mask'], inp['edge_mask'], noise_level=nlb, **kw)
    torch.cuda.synchronize()
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    outs_a, outs_b = [], []
    for _ in range(6):                                      # interleave the launches of the two streams
        with torch.cuda.stream(sa):
            outs_a.append(ma(inp['t'], inp['xh'], inp['node_mask'], inp['edge_mask'], noise_level=nla, **kw))
        with torch.cuda.stream(sb):
            outs_b.append(mb(inp['t'], inp['xh'], inp['node_mask'], inp['edge_mask'], noise_level=nlb, **kw))
    torch.cuda.synchronize()
    for xa, ea in outs_a:
        assert torch.equal(xa, ref_a[0]) and torch.equal(ea, ref_a[1])
    for xb, eb in outs_b:
        assert torch.equal(xb, ref_b[0]) and torch.equal(eb, ref_b[1])
