"""ORACLE tooling (test infrastructure): generate tests/golden/*.pt from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):   python -m oracle.make_golden
The reference modules are imported through oracle/shim (see oracle/README.md); weights come
from jodo_b200.params.synth_state_dict (deterministic from name/shape/seed, so they never need
to be stored) and inputs from jodo_b200.synth.make_batch.  Each fixture stores the inputs,
the reference's fp32 outputs, its fp64 outputs (module cast to double, same fp32-valued inputs)
and, for some cases, the per-block (h, pos) the reference produced.
"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from jodo_b200 import configs, synth  # noqa: E402
from jodo_b200.params import param_spec, synth_state_dict  # noqa: E402
from oracle import ref_loader  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')

# name -> (config name, reference config file, overrides, batch kwargs, weight kwargs)
CASES = {
    'qm9_first': ('qm9_uncond', 'vpsde_qm9_uncond_jodo', {}, dict(n_nodes=[3, 9, 14, 18, 5, 12], seed=1),
                  dict(seed=0, perturb=True)),
    'qm9_first_default_init': ('qm9_uncond', 'vpsde_qm9_uncond_jodo', {}, dict(n_nodes=[4, 11, 7], seed=2),
                               dict(seed=1, perturb=False)),
    'qm9_selfcond': ('qm9_uncond', 'vpsde_qm9_uncond_jodo', {},
                     dict(n_nodes=[3, 9, 14, 18, 5, 29], seed=3, self_cond=True), dict(seed=0, perturb=True)),
    'qm9_cond_ctx': ('qm9_cond', 'vpsde_qm9_cond_jodo', {},
                     dict(n_nodes=[6, 13, 10, 17], seed=4, self_cond=True, context=True), dict(seed=2, perturb=True)),
    'geom_l8': ('geom_l8', 'vpsde_geom_uncond_jodo', {'n_layers': 8},
                dict(n_nodes=[20, 7, 31], seed=5, self_cond=True), dict(seed=3, perturb=True)),
    'geom_l10_first': ('geom_l10', 'vpsde_geom_uncond_jodo', {}, dict(n_nodes=[9, 23], seed=6),
                       dict(seed=4, perturb=True)),
    'geom_large': ('geom_large', 'vpsde_geom_uncond_jodo', {'nf': 384},
                   dict(n_nodes=[12, 6], seed=7, self_cond=True), dict(seed=5, perturb=True)),
    'qm9_cond_multi': ('qm9_cond_multi', 'vpsde_qm9_cond_multi_jodo', {},
                       dict(n_nodes=[7, 15, 3], seed=10, self_cond=True, context=True), dict(seed=8, perturb=True)),
    'qm9_sim': ('qm9_sim', 'vpsde_qm9_uncond_jodo', {'name': 'DGT_concat_sim'},
                dict(n_nodes=[5, 16, 9, 22], seed=12, self_cond=True), dict(seed=9, perturb=True)),
    'moses_2d': ('moses_2d', 'vpsde_moses_2d_jodo', {}, dict(n_nodes=[8, 27, 19, 23], seed=8, self_cond=True),
                 dict(seed=6)),
    'moses_2d_first': ('moses_2d', 'vpsde_moses_2d_jodo', {}, dict(n_nodes=[20, 11], seed=9), dict(seed=7)),
}


def _isolate_atom(batch, b=1, i=0):
    """Quirk 3: make atom i of molecule b a target with NO adjacent source in either
    adjacency head (cond bond channel < th everywhere, cond position far away) -> uniform attention."""
    batch['cond_edge_x'][b, i, :, 0] = -1.0
    batch['cond_edge_x'][b, :, i, 0] = -1.0
    batch['cond_x'][b, i, :3] += 25.0


def run_case(ref, name):
    cfg_name, ref_cfg_file, overrides, bkw, wkw = CASES[name]
    ours = configs.NAMED[cfg_name]()
    rcfg = ref_loader.load_config(ref_cfg_file)
    for k, v in overrides.items():
        rcfg.model[k] = v
    for k in ours.model:                       # our Config restates the reference's keys: check, don't trust
        if k in rcfg.model:
            assert rcfg.model[k] == ours.model[k], (k, rcfg.model[k], ours.model[k])
    for k in ('atom_types', 'max_node', 'fc_scale', 'info_name'):
        if k in ours.data:
            assert rcfg.data[k] == ours.data[k], k
    model = ref.model_utils._MODELS[rcfg.model.name](rcfg)
    spec = param_spec(ours)
    ref_spec = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
    assert spec == ref_spec, 'parameter tree differs from the reference'
    with open(os.path.join(GOLD, f'param_tree_{cfg_name}.json'), 'w') as f:
        json.dump([[k, list(s)] for k, s in ref_spec], f)
    sd = synth_state_dict(spec, **wkw)
    model.load_state_dict(sd, strict=True)
    model.eval()
    batch = synth.make_batch(ours, len(bkw['n_nodes']), **bkw)
    if name == 'qm9_selfcond':
        _isolate_atom(batch)
    inter = []
    hooks = []
    if name in ('qm9_selfcond', 'geom_l8'):
        for i in range(rcfg.model.n_layers):
            hooks.append(model._modules[f'e_block_{i}'].register_forward_hook(
                lambda mod, inp, out: inter.append((out[0].detach().clone(), out[2].detach().clone()))))

    def call(mdl, dt):
        cast = lambda x: None if x is None else x.to(dt)
        with torch.no_grad():
            return mdl(cast(batch['t']), cast(batch['xh']), cast(batch['node_mask']), cast(batch['edge_mask']),
                       edge_x=cast(batch['edge_x']), noise_level=cast(batch['noise_level']),
                       cond_x=cast(batch['cond_x']), cond_edge_x=cast(batch['cond_edge_x']),
                       context=cast(batch['context']))

    x32, e32 = call(model, torch.float32)
    inter.clear()
    x64, e64 = call(model.double(), torch.float64)
    for h in hooks:
        h.remove()
    out = dict(case=name, config=cfg_name, weights=wkw, batch_kwargs=bkw)
    out['inputs'] = {k: v for k, v in batch.items()}
    out['ref_fp32'] = (x32, e32)
    out['ref_fp64'] = (x64, e64)
    if inter:
        B, N = batch['xh'].shape[:2]
        out['blocks_fp64'] = [(h.reshape(B, N, -1).float(), p.reshape(B, N, 3)) for h, p in inter]
    torch.save(out, os.path.join(GOLD, f'{name}.pt'))
    print(f'{name}: x {tuple(x32.shape)} e {tuple(e32.shape)} |x32-x64|max={float((x32 - x64).abs().max()):.2e} '
          f'|x|max={float(x64.abs().max()):.3f} |e|max={float(e64.abs().max()):.3f}')


def run_sampler_case(ref):
    """3 ancestral steps of the reference AncestralSampler (sampling.py:518-596) on the real
    1000-step grid, recording every noise draw so the chain can be replayed elsewhere."""
    ours = configs.qm9_uncond()
    rcfg = ref_loader.load_config('vpsde_qm9_uncond_jodo')
    model = ref.model_utils._MODELS[rcfg.model.name](rcfg)
    model.load_state_dict(synth_state_dict(param_spec(ours), seed=0, perturb=True), strict=True)
    model.eval()
    ns = ref.noise_schedule.NoiseScheduleVP(rcfg.sde.schedule, continuous_beta_0=rcfg.sde.continuous_beta_0,
                                            continuous_beta_1=rcfg.sde.continuous_beta_1)
    steps = 1000
    time_steps = torch.linspace(ns.T, 1e-3, steps)
    sel = torch.tensor([0, 1, 500, 998, 999])
    sampler = ref.sampling.AncestralSampler(ns, time_steps[sel], True, True, True, ref.utils.get_self_cond_fn(rcfg))
    # keep the true s (next grid point) for the chosen t's
    s_full = torch.cat([time_steps[1:], torch.zeros(1)])
    sampler.s_array = s_full[sel]
    batch = synth.make_batch(ours, 3, seed=11, n_nodes=[5, 12, 8])
    rec = {'node': [], 'edge': []}
    orig_n, orig_e = ref.sampling.sample_combined_position_feature_noise, ref.sampling.sample_symmetric_edge_feature_noise

    def rec_n(*a, **k):
        v = orig_n(*a, **k)
        rec['node'].append(v.clone())
        return v

    def rec_e(*a, **k):
        v = orig_e(*a, **k)
        rec['edge'].append(v.clone())
        return v

    ref.sampling.sample_combined_position_feature_noise = rec_n
    ref.sampling.sample_symmetric_edge_feature_noise = rec_e
    torch.manual_seed(123)
    try:
        with torch.no_grad():
            x, ex = sampler.sampling(model, batch['xh'], batch['node_mask'], batch['edge_mask'], batch['edge_x'], None)
    finally:
        ref.sampling.sample_combined_position_feature_noise = orig_n
        ref.sampling.sample_symmetric_edge_feature_noise = orig_e
    alpha_sigma = [tuple(float(v) for v in ns.marginal_prob(t)) for t in time_steps[sel]]
    out = dict(case='qm9_ancestral_chain', config='qm9_uncond', weights=dict(seed=0, perturb=True),
               inputs=batch, t=time_steps[sel], s=s_full[sel], noise_node=rec['node'], noise_edge=rec['edge'],
               x_mean=x, edge_x_mean=ex, alpha_sigma=alpha_sigma, T=ns.T)
    torch.save(out, os.path.join(GOLD, 'qm9_ancestral_chain.pt'))
    print('qm9_ancestral_chain:', tuple(x.shape), tuple(ex.shape), float(x.abs().max()))


def run_sampler2d_case(ref):
    """5 steps of the reference AncestralSampler_2D (sampling.py:599-661) with DGT_concat_2D, every noise draw recorded."""
    ours = configs.NAMED['moses_2d']()
    rcfg = ref_loader.load_config('vpsde_moses_2d_jodo')
    model = ref.model_utils._MODELS[rcfg.model.name](rcfg)
    model.load_state_dict(synth_state_dict(param_spec(ours), seed=6), strict=True)
    model.eval()
    ns = ref.noise_schedule.NoiseScheduleVP(rcfg.sde.schedule, continuous_beta_0=rcfg.sde.continuous_beta_0,
                                            continuous_beta_1=rcfg.sde.continuous_beta_1)
    time_steps = torch.linspace(ns.T, 1e-3, 1000)
    sel = torch.tensor([0, 1, 500, 998, 999])
    sampler = ref.sampling.AncestralSampler_2D(ns, time_steps[sel], True, True)
    s_full = torch.cat([time_steps[1:], torch.zeros(1)])
    sampler.s_array = s_full[sel]
    batch = synth.make_batch(ours, 3, seed=12, n_nodes=[9, 20, 14])
    rec = {'node': [], 'edge': []}
    orig_n, orig_e = ref.sampling.sample_gaussian_with_mask, ref.sampling.sample_symmetric_edge_feature_noise

    def rec_n(*a, **k):
        v = orig_n(*a, **k)
        rec['node'].append(v.clone())
        return v

    def rec_e(*a, **k):
        v = orig_e(*a, **k)
        rec['edge'].append(v.clone())
        return v

    ref.sampling.sample_gaussian_with_mask = rec_n
    ref.sampling.sample_symmetric_edge_feature_noise = rec_e
    torch.manual_seed(321)
    try:
        with torch.no_grad():
            x, ex = sampler.sampling(model, batch['xh'], batch['node_mask'], batch['edge_mask'], batch['edge_x'], None)
    finally:
        ref.sampling.sample_gaussian_with_mask = orig_n
        ref.sampling.sample_symmetric_edge_feature_noise = orig_e
    out = dict(case='moses_2d_chain', config='moses_2d', weights=dict(seed=6), inputs=batch, t=time_steps[sel],
               s=s_full[sel], noise_node=rec['node'], noise_edge=rec['edge'], x_mean=x, edge_x_mean=ex, seed=321)
    torch.save(out, os.path.join(GOLD, 'moses_2d_chain.pt'))
    print('moses_2d_chain:', tuple(x.shape), tuple(ex.shape), float(x.abs().max()), len(rec['node']), len(rec['edge']))


def run_dpm_case(ref):
    """3 outer steps (6 model evaluations) of the reference DPM_Solver_hybrid ('singlestep_fixed', order 2,
    mix_dpm_solver.py:285-335) on the conditional QM9 model with a context, recording every position-noise draw."""
    ours = configs.qm9_cond()
    rcfg = ref_loader.load_config('vpsde_qm9_cond_jodo')
    rcfg.sampling.steps = 6
    rcfg.sampling.dpm_solver_order = 2
    rcfg.sampling.dpm_solver_method = 'singlestep_fixed'
    model = ref.model_utils._MODELS[rcfg.model.name](rcfg)
    model.load_state_dict(synth_state_dict(param_spec(ours), seed=3, perturb=True), strict=True)
    model.eval()
    ns = ref.noise_schedule.NoiseScheduleVP(rcfg.sde.schedule, continuous_beta_0=rcfg.sde.continuous_beta_0,
                                            continuous_beta_1=rcfg.sde.continuous_beta_1)
    solver = ref.mix_dpm_solver.DPM_Solver_hybrid(ns, rcfg)
    batch = synth.make_batch(ours, 3, seed=21, n_nodes=[6, 11, 4], context=True)
    rec = []
    orig = ref.mix_dpm_solver.sample_center_gravity_zero_gaussian_with_mask

    def rec_fn(*a, **k):
        v = orig(*a, **k)
        rec.append(v.clone())
        return v

    ref.mix_dpm_solver.sample_center_gravity_zero_gaussian_with_mask = rec_fn
    torch.manual_seed(77)
    try:
        x, ex = solver.sampling(model, batch['xh'], batch['node_mask'], batch['edge_mask'], batch['edge_x'], batch['context'])
    finally:
        ref.mix_dpm_solver.sample_center_gravity_zero_gaussian_with_mask = orig
    out = dict(case='qm9_cond_dpm_chain', config='qm9_cond', weights=dict(seed=3, perturb=True), inputs=batch,
               steps=6, order=2, noise_pos=rec, x=x, edge_x=ex)
    torch.save(out, os.path.join(GOLD, 'qm9_cond_dpm_chain.pt'))
    print('qm9_cond_dpm_chain:', tuple(x.shape), tuple(ex.shape), len(rec), 'noise draws', float(x.abs().max()))


def run_postprocess_case(ref):
    """post_process + mol_process of the reference (sampling.py:12-97) on a QM9-shaped and a GEOM-shaped final state."""
    import importlib
    out = {}
    for cfg_name, fname in (('qm9_uncond', 'vpsde_qm9_uncond_jodo'), ('geom_l8', 'vpsde_geom_uncond_jodo')):
        ours = configs.NAMED[cfg_name]()
        rcfg = ref_loader.load_config(fname)
        b = synth.make_batch(ours, 6, seed=31, max_n=40)
        xh = b['xh'] * 0.8
        ex = b['edge_x'] * 0.9
        inv = ref.utils.get_data_inverse_scaler(rcfg)
        pos, one_hot, fc, edge = ref.sampling.post_process(xh.clone(), rcfg.data.atom_types, rcfg.model.include_fc_charge,
                                                           b['node_mask'], inv, ex.clone(), b['edge_mask'],
                                                           rcfg.data.compress_edge)
        mols = ref.sampling.mol_process(one_hot, pos, fc, b['n_nodes'], edge)
        out[cfg_name] = dict(xh=xh, edge_x=ex, node_mask=b['node_mask'], edge_mask=b['edge_mask'], n_nodes=b['n_nodes'],
                             pos=pos, one_hot=one_hot, fc=fc, edge=edge, mols=mols)
    ours = configs.NAMED['moses_2d']()                         # 2-D: post_process_2D + mol_process_2D (sampling.py:35-50, 100-144)
    rcfg = ref_loader.load_config('vpsde_moses_2d_jodo')
    b = synth.make_batch(ours, 6, seed=32)
    xh, ex = b['xh'] * 0.8, b['edge_x'] * 0.9
    inv = ref.utils.get_data_inverse_scaler(rcfg)
    one_hot, fc, edge = ref.sampling.post_process_2D(xh.clone(), rcfg.data.atom_types, rcfg.model.include_fc_charge,
                                                     b['node_mask'], inv, ex.clone(), b['edge_mask'], rcfg.data.compress_edge)
    mols = ref.sampling.mol_process_2D(one_hot, fc, b['n_nodes'], edge)
    out['moses_2d'] = dict(xh=xh, edge_x=ex, node_mask=b['node_mask'], edge_mask=b['edge_mask'], n_nodes=b['n_nodes'],
                           one_hot=one_hot, fc=fc, edge=edge, mols=mols)
    torch.save(out, os.path.join(GOLD, 'postprocess.pt'))
    print('postprocess:', {k: tuple(v['edge'].shape) for k, v in out.items()})


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    ref = ref_loader.load()
    only = sys.argv[1:]
    for name in CASES:
        if not only or name in only:
            run_case(ref, name)
    if not only or 'chain' in only:
        run_sampler_case(ref)
    if not only or 'chain2d' in only:
        run_sampler2d_case(ref)
    if not only or 'dpm' in only:
        run_dpm_case(ref)
    if not only or 'post' in only:
        run_postprocess_case(ref)


if __name__ == '__main__':
    main()
