"""What fp16 tensor-core operands (10-bit mantissa, saturating) cost over what the metric measures:

* a 200-step ancestral chain on 8 QM9 molecules with replayed noise -- CUDA vs the fp32 and fp64 oracles, compared on the
  INTEGER end products of post_process (atom types, formal charges, bond orders);
* a direct CUDA-vs-oracle comparison on a subsample of the full-size batches (QM9 B = 2500, GEOM-Drugs B = 512);
* operand range: weights scaled until fp16 operands saturate are flagged by jodo_saturation_count, moderately scaled
  weights are not flagged and stay within tolerance.
"""
import pytest
import torch

from jodo_b200 import _lib, configs, postprocess, sampler as S, synth
from jodo_b200.model import MODELS
from jodo_b200.params import param_spec, synth_state_dict
from oracle.dgt_dense import dgt_forward

pytestmark = pytest.mark.gpu
TOL = 2e-3


def _oracle_model(sd, cfg, dtype):
    sd = {k: v.to(dtype) for k, v in sd.items()}

    def model(t, xh, node_mask, edge_mask, **kw):
        c = lambda v: None if v is None else v.to(dtype)
        x, e = dgt_forward(sd, cfg, c(t), c(xh), c(node_mask), c(edge_mask), **{k: c(v) for k, v in kw.items()})
        return x, e
    return model


def _ints(cfg, x, e, nm, em):
    pos, one_hot, fc, bonds = postprocess.post_process(cfg, x.float(), nm.float(), e.float(), em.float())
    return one_hot.argmax(-1) * nm[..., 0].long(), fc[..., 0].long() if fc.dim() == 3 else fc.long(), bonds.long()


def test_long_chain_integer_end_products():
    cfg = configs.NAMED['qm9_uncond']()
    sd = synth_state_dict(param_spec(cfg), seed=11, perturb=True)
    b = synth.make_batch(cfg, 8, seed=77)
    nm, em = b['node_mask'], b['edge_mask']
    B, N, F_ = b['xh'].shape
    ch = b['edge_x'].shape[-1]
    steps = 200
    grid = torch.linspace(0.9946, 1e-3, 1000)[::1000 // steps]
    g = torch.Generator().manual_seed(5)
    noise = [(synth.node_noise(B, N, F_ - 3, nm, g), synth.edge_noise(B, N, ch, em, g)) for _ in range(len(grid))]

    def chain(model, dev, dtype):
        nf = lambda i, kind: noise[i][0 if kind == 'node' else 1].to(dev, dtype)
        smp = S.AncestralSampler(S.CosineVP(), grid, noise_fn=nf)
        x, e = smp.sampling(model, b['xh'].to(dev, dtype), nm.to(dev, dtype), em.to(dev, dtype), b['edge_x'].to(dev, dtype))
        return x.detach().cpu().double(), e.detach().cpu().double()

    cuda_model = MODELS[cfg.model.name](cfg)
    cuda_model.load_state_dict(sd, strict=True)
    xc, ec = chain(cuda_model.cuda().eval(), 'cuda', torch.float32)
    x32, e32 = chain(_oracle_model(sd, cfg, torch.float32), 'cpu', torch.float32)
    x64, e64 = chain(_oracle_model(sd, cfg, torch.float64), 'cpu', torch.float64)
    ref = _ints(cfg, x64, e64, nm, em)
    n_atoms, n_bonds = int(nm.sum()), int(em.sum())
    rep = {}
    for name, (x, e) in (('cuda', (xc, ec)), ('oracle_fp32', (x32, e32))):
        got = _ints(cfg, x, e, nm, em)
        rep[name] = dict(atom_types=int((got[0] != ref[0]).sum()), charges=int((got[1] != ref[1]).sum()),
                         bonds=int((got[2] != ref[2]).sum()), pos=float((x[..., :3] - x64[..., :3]).abs().max()),
                         feat=float((x[..., 3:] - x64[..., 3:]).abs().max()), edge=float((e - e64).abs().max()))
    print(f'{steps}-step chain, 8 molecules ({n_atoms} atoms, {n_bonds} ordered pairs), mismatches against the fp64 oracle:', rep)
    c, f = rep['cuda'], rep['oracle_fp32']
    # integer end products: the CUDA path may not be worse than a small multiple of what plain fp32 does to the same chain
    assert c['atom_types'] + c['charges'] <= 2 + 3 * (f['atom_types'] + f['charges']) + n_atoms // 50
    assert c['bonds'] <= 4 + 3 * f['bonds'] + n_bonds // 50
    assert torch.isfinite(xc).all() and torch.isfinite(ec).all()


@pytest.mark.parametrize('cfg_name,batch,max_n,nsub', [('qm9_uncond', 2500, None, 32), ('geom_l8', 512, 80, 12)])
def test_subsample_of_full_size_batch_matches_oracle(cfg_name, batch, max_n, nsub):
    """The BASELINE-sized batches directly against the oracle on a subsample: molecules are independent, so molecule b of
    the big batch must equal the oracle's output for a small batch that contains only the sampled molecules."""
    cfg = configs.NAMED[cfg_name]()
    sd = synth_state_dict(param_spec(cfg), seed=2, perturb=True)
    model = MODELS[cfg.model.name](cfg)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    b = synth.make_batch(cfg, batch, seed=42, max_n=max_n, self_cond=True, noise_level=1.25)
    d = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()}
    x, e = model(d['t'], d['xh'], d['node_mask'], d['edge_mask'], edge_x=d['edge_x'], noise_level=d['noise_level'],
                 cond_x=d['cond_x'], cond_edge_x=d['cond_edge_x'])
    x, e = x.cpu(), e.cpu()
    idx = torch.randperm(batch, generator=torch.Generator().manual_seed(1))[:nsub]
    idx[0] = int(b['n_nodes'].argmax())                        # always include a largest molecule
    n = int(b['n_nodes'][idx].max())
    N = b['xh'].shape[1]
    sub = dict(t=b['t'][idx], xh=b['xh'][idx, :n], node_mask=b['node_mask'][idx, :n], edge_x=b['edge_x'][idx, :n, :n],
               edge_mask=b['edge_mask'].reshape(batch, N, N, 1)[idx, :n, :n].reshape(-1, 1), noise_level=b['noise_level'][idx],
               cond_x=b['cond_x'][idx, :n], cond_edge_x=b['cond_edge_x'][idx, :n, :n])
    c = lambda v: v.double()
    ox, oe = dgt_forward({k: v.double() for k, v in sd.items()}, cfg, c(sub['t']), c(sub['xh']), c(sub['node_mask']), c(sub['edge_mask']),
                         edge_x=c(sub['edge_x']), noise_level=c(sub['noise_level']), cond_x=c(sub['cond_x']),
                         cond_edge_x=c(sub['cond_edge_x']))
    ex = float((x[idx, :n].double() - ox).abs().max() / ox.abs().max())
    ee = float((e[idx, :n, :n].double() - oe).abs().max() / oe.abs().max())
    print(f'{cfg_name} B={batch}: {nsub} sampled molecules (n up to {n}) vs fp64 oracle: x {ex:.2e}  e {ee:.2e}')
    assert ex < 3e-3 and ee < 3e-3
    assert float(x[idx, n:].abs().max() if n < N else 0.0) == 0.0


def test_operand_saturation_is_counted():
    """fp16 operands clamp at +-65504 without a trap.  Weights scaled by 8 stay in range: no clamped store, result within
    tolerance of the oracle on the same weights.  The hoisted input_lin / node2edge_lin parts scaled by 3e5 leave the
    range: the counter reports it (the result is then NOT trustworthy, which is what the counter is for)."""
    from helpers import load_golden, oracle_forward
    g, cfg = load_golden('qm9_selfcond')

    def run(scale):
        sd = synth_state_dict(param_spec(cfg), seed=0, perturb=True)
        for k in sd:
            if 'equi_update.input_lin.weight' in k or 'node2edge_lin.weight' in k:
                sd[k] = sd[k] * scale
        m = MODELS[cfg.model.name](cfg)
        m.load_state_dict(sd, strict=True)
        m = m.cuda().eval()
        inp = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in g['inputs'].items()}
        _lib.saturation_count(reset=True)
        x, e = m(inp['t'], inp['xh'], inp['node_mask'], inp['edge_mask'], edge_x=inp['edge_x'], noise_level=inp['noise_level'],
                 cond_x=inp['cond_x'], cond_edge_x=inp['cond_edge_x'])
        return sd, x, e, _lib.saturation_count(reset=True)

    sd, x, e, cnt = run(8.0)
    ox, oe = oracle_forward(sd, cfg, g['inputs'], torch.float64)
    rx = float((x.double().cpu() - ox).abs().max() / ox.abs().max())
    re_ = float((e.double().cpu() - oe).abs().max() / oe.abs().max())
    print(f'weights x8: clamped stores {cnt}, rel err x {rx:.2e} e {re_:.2e}')
    assert cnt == 0 and rx < 5e-3 and re_ < 5e-3
    _, x, e, cnt = run(3e5)
    print(f'weights x3e5: clamped stores {cnt}')
    assert cnt > 0
