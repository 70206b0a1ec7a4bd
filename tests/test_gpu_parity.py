"""GPU parity: the CUDA path (through the C ABI, driven by the drop-in module) against the golden
outputs of the unmodified reference and against the oracle.

Tolerance: the kernels feed the tensor cores tf32 operands (10-bit mantissa) with fp32 accumulation;
everything else is fp32.  One forward stays within 2e-3 of the fp64 reference, relative to the largest
magnitude of the compared tensor (measured: 2e-4 .. 1e-3; 3e-3 for the bonds of geom_l8).  Masked entries must be exactly zero and the edge
output exactly symmetric."""
import pytest
import torch

from helpers import load_golden
from stage_diag import run_case

pytestmark = pytest.mark.gpu

TOL = 2e-3            # measured 2e-4 .. 1e-3 on one forward
TOL_CASE = {'geom_l8': 3e-3}          # bonds of the r = 4, 31-atom fixture: 2.4e-3
CASES = ['qm9_first', 'qm9_first_default_init', 'qm9_selfcond', 'qm9_cond_ctx', 'geom_l8', 'geom_l10_first',
         'geom_large',          # nf = 384: the wide path (jodo_b200/wide.py)
         'qm9_cond_multi',      # cond_DGT_concat with two properties (cond_ch = 2)
         'moses_2d', 'moses_2d_first',          # DGT_concat_2D (no coordinates) on the wide path
         'qm9_sim']                             # DGT_concat_sim (no adjacency heads) on the wide path


@pytest.mark.parametrize('name', CASES)
def test_forward_matches_reference(name):
    x, e, rep = run_case(name, verbose=True)
    g, _ = load_golden(name)
    rx, re_ = g['ref_fp64']
    inp = g['inputs']
    B, N = rx.shape[:2]
    x, e = x.cpu(), e.cpu()
    worst = {k: v for k, v in rep if k.startswith('out.')}
    assert all(v < TOL_CASE.get(name, TOL) for v in worst.values()), (worst, rep)
    # integer-exact properties
    nm = inp['node_mask']
    em = inp['edge_mask'].reshape(B, N, N, 1)
    assert float((x * (1 - nm)).abs().max()) == 0.0
    assert float((e * (1 - em)).abs().max()) == 0.0
    assert float((e - e.permute(0, 2, 1, 3)).abs().max()) == 0.0
    # CoM-free positions (reference assert_mean_zero_with_mask, models/utils.py:59-64)
    if not name.startswith('moses'):
        assert float(x[..., :3].sum(1).abs().max()) < 1e-4


@pytest.mark.parametrize('name', ['qm9_selfcond', 'geom_large'])
def test_uniform_conditioning_fast_path_matches_general_path(name):
    """All molecules at one noise level (what the samplers feed, sampling.py:549) take the device-detected fast path
    (row 0 of the AdaLN table through constant memory, the table as one matrix-vector product); perturbing one molecule's noise
    level forces the general per-molecule path; the other molecules must come out the same on both paths."""
    from helpers import golden_weights
    from jodo_b200.model import MODELS
    g, cfg = load_golden(name)            # geom_large: the wide path, whose row kernels read table row 0 under the same flag
    model = MODELS[cfg.model.name](cfg)
    model.load_state_dict(golden_weights(g, cfg), strict=True)
    model = model.cuda().eval()
    inp = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in g['inputs'].items()}
    nl = torch.full_like(inp['noise_level'], 1.25)
    kw = dict(edge_x=inp['edge_x'], cond_x=inp['cond_x'], cond_edge_x=inp['cond_edge_x'])
    xa, ea = model(inp['t'], inp['xh'], inp['node_mask'], inp['edge_mask'], noise_level=nl, **kw)
    flags_uni = int(next(iter(model._plans.values()))[1].flags[2])
    nl2 = nl.clone()
    nl2[-1] = 1.5
    xb, eb = model(inp['t'], inp['xh'], inp['node_mask'], inp['edge_mask'], noise_level=nl2, **kw)
    flags_gen = int(next(iter(model._plans.values()))[1].flags[2])
    assert flags_uni == 0 and flags_gen == 1
    B = xa.shape[0]
    # Molecules 0..B-2 see the same conditioning on both paths.  Their AdaLN rows come from two kernels -- row 0 as a matrix-vector
    # product on the uniform path (jodo_row0_linear), every row through the tensor-core GEMM on the general one: same fp16-rounded
    # operands, different fp32 summation order -- so the rows agree to ~1e-6 and the outputs to the occasional flipped fp16
    # operand rounding downstream (measured 5e-5 of max |x|); with JODO_EQUI_LIN=1 the uniform path also composes coord_mlp.0
    # into input_lin (csrc/equi_lin.cu).  Either way each path agrees with the fp64 oracle.
    from jodo_b200 import pack as _pack
    tol_ab = 1e-3 if _pack.EQUI_LIN else 5e-4
    assert float((xa[:B - 1] - xb[:B - 1]).abs().max()) < tol_ab * float(xa.abs().max())
    assert float((ea[:B - 1] - eb[:B - 1]).abs().max()) < tol_ab * float(ea.abs().max())
    from helpers import oracle_forward
    sd = golden_weights(g, cfg)
    for nl_, (x_, e_), tag in ((nl, (xa, ea), 'uniform'), (nl2, (xb, eb), 'general')):
        oi = dict(g['inputs'])
        oi['noise_level'] = nl_.cpu()
        ox, oe = oracle_forward(sd, cfg, oi, torch.float64)
        ex = float((x_.double().cpu() - ox).abs().max() / ox.abs().max())
        ee = float((e_.double().cpu() - oe).abs().max() / oe.abs().max())
        px = float((x_.double().cpu() - ox)[..., :3].abs().max() / ox[..., :3].abs().max())
        print(f'{tag} path vs fp64 oracle: x {ex:.2e} (positions {px:.2e})  e {ee:.2e}')
        assert ex < TOL and ee < TOL and px < TOL


def test_ab_variants_of_the_coordinate_kernels_stay_parity_green(monkeypatch):
    """The two coordinate-branch variants that are kept for A/B runs but are off by default (measured slower on B200,
    DESIGN.md section 5) compute the same function: coord_mlp.0 composed into input_lin under uniform conditioning
    (csrc/equi_lin.cu, JODO_EQUI_LIN=1) on the nf = 256 path, and the fused LayerNorm -> coord_mlp.0 kernel of the wide path
    (csrc/wide_equi.cu, JODO_WIDE_EQUI_FUSED=1) -- each against the fp64 oracle and against the default kernels.  Also the
    unfused edge FFN of the wide path (LayerNorm row kernel + two GEMMs, JODO_WIDE_FFN_UNFUSED=1) against the fused kernel
    (csrc/wide_ffn.cu)."""
    from helpers import golden_weights, oracle_forward
    from jodo_b200 import pack as _pack, wide as _wide
    from jodo_b200.model import MODELS

    def run(name, uniform):
        g, cfg = load_golden(name)
        sd = golden_weights(g, cfg)
        model = MODELS[cfg.model.name](cfg)
        model.load_state_dict(sd, strict=True)
        model = model.cuda().eval()
        inp = dict(g['inputs'])
        if uniform:
            inp['noise_level'] = torch.full_like(inp['noise_level'], 1.25)
        dev = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in inp.items()}
        x, e = model(dev['t'], dev['xh'], dev['node_mask'], dev['edge_mask'], noise_level=dev['noise_level'], edge_x=dev['edge_x'],
                     cond_x=dev['cond_x'], cond_edge_x=dev['cond_edge_x'], context=dev['context'])
        ox, oe = oracle_forward(sd, cfg, inp, torch.float64)
        return x.double().cpu(), e.double().cpu(), ox, oe

    for name, uniform, mod, flag in (('qm9_selfcond', True, _pack, 'EQUI_LIN'), ('geom_large', False, _wide, 'FUSED_EQUI'),
                                     ('geom_large', False, _wide, 'FFN_UNFUSED')):
        monkeypatch.setattr(mod, flag, False)
        x0, e0, ox, oe = run(name, uniform)
        monkeypatch.setattr(mod, flag, True)
        x1, e1, _, _ = run(name, uniform)
        ex, ee = float((x1 - ox).abs().max() / ox.abs().max()), float((e1 - oe).abs().max() / oe.abs().max())
        dx = float((x1 - x0).abs().max() / ox.abs().max())
        print(f'{name} [{flag}]: vs fp64 oracle x {ex:.2e} e {ee:.2e}; vs the default kernels x {dx:.2e}')
        assert ex < TOL and ee < TOL and dx < TOL
        assert dx > 0.0 or flag == 'FFN_UNFUSED'                         # the variant really ran (different rounding)
