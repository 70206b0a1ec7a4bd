"""The caller of the hot path: one ancestral reverse-SDE step (reference sampling.py:530-596).

CPU: the oracle restatement (oracle/sampler_ref.py) and the product-side host sampler
(jodo_b200/sampler.py), both driving the oracle denoiser, replay the chain recorded from the
unmodified reference sampler (tests/golden/qm9_ancestral_chain.pt) with identical noise draws.
GPU: the same chain with the CUDA denoiser behind the product sampler."""
import pytest
import torch

from helpers import golden_weights, load_golden
from jodo_b200 import sampler as S


def _oracle_model(sd, cfg, dtype):
    from oracle.dgt_dense import dgt_forward
    sd = {k: v.to(dtype) for k, v in sd.items()}

    def model(t, xh, node_mask, edge_mask, **kw):
        return dgt_forward(sd, cfg, t, xh, node_mask, edge_mask, **kw)
    return model


def _cast(d, dtype):
    return {k: (v.to(dtype) if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in d.items()}


def test_schedule_matches_reference():
    g, _ = load_golden('qm9_ancestral_chain')
    from oracle import sampler_ref as R
    sch = S.CosineVP()
    assert sch.T == g['T'] == R.T_END
    for t, (a, s) in zip(g['t'], g['alpha_sigma']):
        a1, s1 = sch.marginal_prob(t)
        a2, s2 = R.marginal_prob(t)
        assert float(a1) == a == float(a2) and float(s1) == s == float(s2)      # same fp32 operation order: bit exact


# The recorded chain ran in fp32.  Replaying it in fp64 shows how far the reference's own fp32 rounding moves a
# free-running chain: the self-conditioning adjacency heads threshold the previous prediction (cond_edge_x >= 0,
# |dpos|^2 <= cut-off; models/mol_gnn.py:523-525, models/utils.py:111-119), so a 1e-6 difference can flip a bit and
# change the next call discretely (measured: 1.4e-3 of max |x|, 1.4e-2 of max |e| after 5 calls).  Free-running
# comparisons therefore carry a loose tolerance; the tight check is the teacher-forced one (tests/test_gpu_parity.py
# and test_cuda_chain_teacher_forced below).
@pytest.mark.parametrize('dtype,tol', [(torch.float32, 2e-4), (torch.float64, 3e-2)])
def test_oracle_sampler_replays_reference_chain(dtype, tol):
    from oracle import sampler_ref as R
    g, cfg = load_golden('qm9_ancestral_chain')
    model = _oracle_model(golden_weights(g, cfg), cfg, dtype)
    b = _cast(g['inputs'], dtype)
    nn_ = [z.to(dtype) for z in g['noise_node']]
    ne = [z.to(dtype) for z in g['noise_edge']]
    with torch.no_grad():
        xm, em = R.replay_chain(model, g['t'].to(dtype), g['s'].to(dtype), b['xh'], b['edge_x'], b['node_mask'],
                                b['edge_mask'], nn_, ne)
    sx, se = float(g['x_mean'].abs().max()), float(g['edge_x_mean'].abs().max())
    assert float((xm.float() - g['x_mean']).abs().max()) < tol * sx
    assert float((em.float() - g['edge_x_mean']).abs().max()) < tol * se


def test_product_sampler_replays_reference_chain():
    g, cfg = load_golden('qm9_ancestral_chain')
    model = _oracle_model(golden_weights(g, cfg), cfg, torch.float32)
    b = g['inputs']
    noise = lambda i, kind: g['noise_node'][i] if kind == 'node' else g['noise_edge'][i]
    smp = S.AncestralSampler(S.CosineVP(), g['t'], noise_fn=noise, s_array=g['s'])
    xm, em = smp.sampling(model, b['xh'], b['node_mask'], b['edge_mask'], b['edge_x'])
    assert float((xm - g['x_mean']).abs().max()) < 2e-4 * float(g['x_mean'].abs().max())
    assert float((em - g['edge_x_mean']).abs().max()) < 2e-4 * float(g['edge_x_mean'].abs().max())
    # CoM invariant the reference asserts at the end of the chain (sampling.py:591, models/utils.py:59-64)
    assert float(xm[..., :3].sum(1).abs().max()) < 1e-4


def test_noise_draw_order_matches_reference():
    """node noise then edge noise per step, from one generator: the stream the reference draws
    (models/utils.py:83-99) -- recorded under torch.manual_seed(123) by make_golden."""
    g, cfg = load_golden('qm9_ancestral_chain')
    b = g['inputs']
    B, N = b['xh'].shape[:2]
    gen = torch.Generator().manual_seed(123)
    for i in range(len(g['t'])):
        zn = S.node_noise(B, N, b['xh'].shape[2] - 3, b['node_mask'], gen)
        ze = S.edge_noise(B, N, b['edge_x'].shape[-1], b['edge_mask'], gen)
        assert torch.equal(zn, g['noise_node'][i])
        assert torch.equal(ze, g['noise_edge'][i])


def test_default_s_array_is_shifted_grid():
    grid = torch.linspace(0.9946, 1e-3, 1000)
    c = S.ancestral_coefficients(S.CosineVP(), grid)
    c2 = S.ancestral_coefficients(S.CosineVP(), grid, torch.cat([grid[1:], torch.zeros(1)]))
    assert torch.equal(c, c2)
    assert torch.isfinite(c).all()
    assert float(c[-1, 2]) < 1e-3             # last step: s = 0 -> sigma_s ~ 0 (fp32 rounding leaves 2.4e-4, as in the reference)


@pytest.mark.gpu
def test_cuda_chain_teacher_forced():
    """SURVEY.md 8d parity protocol (1): at every step of the recorded chain the CUDA denoiser and the fp32 oracle
    get the identical (x, edge_x, cond_x, cond_edge_x, noise_level) -- the oracle's own chain state -- and their
    predictions must agree within the one-call tolerance (5e-3 of the largest magnitude, see test_gpu_parity.py)."""
    from jodo_b200.model import MODELS
    g, cfg = load_golden('qm9_ancestral_chain')
    sd = golden_weights(g, cfg)
    ref = _oracle_model(sd, cfg, torch.float64)
    model = MODELS[cfg.model.name](cfg)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    b = g['inputs']
    nm_d, em_d = b['node_mask'].cuda(), b['edge_mask'].cuda()
    coef = S.ancestral_coefficients(S.CosineVP(), g['t'], g['s'])
    x, ex, cx, cex = b['xh'].double(), b['edge_x'].double(), None, None
    nm, em = b['node_mask'].double(), b['edge_mask'].double()
    f = lambda v: None if v is None else v.float().cuda()
    for i in range(len(g['t'])):
        c_x, c_p, sigma, nl = (float(v) for v in coef[i])
        nlv = torch.full((x.shape[0],), nl, dtype=torch.float64)
        px, pe = ref(g['t'][i].double().expand(x.shape[0]), x, nm, em, edge_x=ex, noise_level=nlv, cond_x=cx,
                     cond_edge_x=cex)
        gx, ge = model(f(nlv), f(x), nm_d, em_d, edge_x=f(ex), noise_level=f(nlv), cond_x=f(cx), cond_edge_x=f(cex))
        assert float((gx.cpu().double() - px).abs().max()) < 5e-3 * float(px.abs().max()), i
        assert float((ge.cpu().double() - pe).abs().max()) < 5e-3 * float(pe.abs().max()), i
        x = c_x * x + c_p * px + sigma * g['noise_node'][i].double()
        ex = c_x * ex + c_p * pe + sigma * g['noise_edge'][i].double()
        cx, cex = px, pe


@pytest.mark.gpu
def test_cuda_chain_free_running():
    """Protocol (2): free-running 5-step chain with replayed noise under the product sampler.  Loose tolerance (see
    the note above the oracle replay test: adjacency bits of the self-conditioning can flip); exact invariants."""
    from jodo_b200.model import MODELS
    g, cfg = load_golden('qm9_ancestral_chain')
    model = MODELS[cfg.model.name](cfg)
    model.load_state_dict(golden_weights(g, cfg), strict=True)
    model = model.cuda().eval()
    b = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in g['inputs'].items()}
    noise = lambda i, kind: (g['noise_node'][i] if kind == 'node' else g['noise_edge'][i]).cuda()
    smp = S.AncestralSampler(S.CosineVP(), g['t'], noise_fn=noise, s_array=g['s'])
    xm, em = smp.sampling(model, b['xh'], b['node_mask'], b['edge_mask'], b['edge_x'])
    xm, em = xm.cpu(), em.cpu()
    assert float((xm - g['x_mean']).abs().max()) < 5e-2 * float(g['x_mean'].abs().max())
    assert float((em - g['edge_x_mean']).abs().max()) < 1e-1 * float(g['edge_x_mean'].abs().max())
    nm = g['inputs']['node_mask']
    assert float((xm * (1 - nm)).abs().max()) == 0.0
    assert float((em - em.permute(0, 2, 1, 3)).abs().max()) == 0.0
    assert float(xm[..., :3].sum(1).abs().max()) < 1e-4


# ---- DPM-Solver++ singlestep, order 2: the driver of conditional QM9 sampling (BASELINE config 5) ---------------------
def test_oracle_dpm_solver_replays_reference_chain():
    from oracle import sampler_ref as R
    g, cfg = load_golden('qm9_cond_dpm_chain')
    model = _oracle_model(golden_weights(g, cfg), cfg, torch.float32)
    b = g['inputs']
    with torch.no_grad():
        x, e = R.dpm_singlestep2_chain(model, b['xh'], b['edge_x'], b['node_mask'], b['edge_mask'], b['context'], g['steps'],
                                       g['noise_pos'])
    assert float((x - g['x']).abs().max()) < 5e-4 * float(g['x'].abs().max())
    assert float((e - g['edge_x']).abs().max()) < 5e-4 * float(g['edge_x'].abs().max())


def test_product_dpm_solver_replays_reference_chain():
    g, cfg = load_golden('qm9_cond_dpm_chain')
    model = _oracle_model(golden_weights(g, cfg), cfg, torch.float32)
    b = g['inputs']
    sol = S.DPMSolverSinglestep(S.CosineVP(), g['steps'], order=g['order'], noise_fn=lambda i: g['noise_pos'][i])
    x, e = sol.sampling(model, b['xh'], b['node_mask'], b['edge_mask'], b['edge_x'], b['context'])
    assert sol.n_evals == g['steps'] and sol.n_noise == len(g['noise_pos'])
    assert float((x - g['x']).abs().max()) < 5e-4 * float(g['x'].abs().max())
    assert float((e - g['edge_x']).abs().max()) < 5e-4 * float(g['edge_x'].abs().max())
    assert float(x[..., :3].sum(1).abs().max()) < 1e-4                  # assert_mean_zero_with_mask, mix_dpm_solver.py:374


def test_position_noise_matches_reference_draws():
    g, cfg = load_golden('qm9_cond_dpm_chain')
    b = g['inputs']
    gen = torch.Generator().manual_seed(77)
    for z in g['noise_pos']:
        assert torch.equal(S.position_noise(b['xh'].shape[0], b['xh'].shape[1], b['node_mask'], gen), z)


@pytest.mark.gpu
def test_cuda_dpm_solver_chain():
    """Conditional model (context path) under the product DPM-Solver with replayed noise, free-running for 6 model
    evaluations: loose tolerance for the same reason as the ancestral chain, exact invariants."""
    from jodo_b200.model import MODELS
    g, cfg = load_golden('qm9_cond_dpm_chain')
    model = MODELS[cfg.model.name](cfg)
    model.load_state_dict(golden_weights(g, cfg), strict=True)
    model = model.cuda().eval()
    b = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in g['inputs'].items()}
    sol = S.DPMSolverSinglestep(S.CosineVP(), g['steps'], order=g['order'], noise_fn=lambda i: g['noise_pos'][i].cuda())
    x, e = sol.sampling(model, b['xh'], b['node_mask'], b['edge_mask'], b['edge_x'], b['context'])
    x, e = x.cpu(), e.cpu()
    assert sol.n_evals == 6
    assert float((x - g['x']).abs().max()) < 5e-2 * float(g['x'].abs().max())
    assert float((e - g['edge_x']).abs().max()) < 1e-1 * float(g['edge_x'].abs().max())
    nm = g['inputs']['node_mask']
    assert float((x * (1 - nm)).abs().max()) == 0.0
    assert float((e - e.permute(0, 2, 1, 3)).abs().max()) == 0.0
    assert float(x[..., :3].sum(1).abs().max()) < 1e-4


@pytest.mark.gpu
def test_fused_ancestral_update_matches_torch_ops():
    """jodo_ancestral_update (posterior mean + CoM-free node noise + symmetric edge noise in two launches) against the
    torch-op form of the same step, same generator seed: the raw draws are identical, so the results agree to the
    rounding of the CoM reduction order."""
    cfg = __import__('jodo_b200.configs', fromlist=['NAMED']).NAMED['qm9_uncond']()
    from jodo_b200 import synth
    b = synth.make_batch(cfg, 37, seed=12, self_cond=True)
    d = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()}

    def fake_model(t, xh, node_mask, edge_mask, **kw):
        return d['cond_x'], d['cond_edge_x']

    grid = torch.linspace(0.9946, 1e-3, 1000)
    outs = []
    for fused in (True, False):
        gen = torch.Generator(device='cuda').manual_seed(99)
        smp = S.AncestralSampler(S.CosineVP(), grid, generator=gen, fused=fused)
        outs.append(smp.step(fake_model, 400, d['xh'], d['edge_x'], d['node_mask'], d['edge_mask'], None, None))
    for a, r in zip(outs[0][:4], outs[1][:4]):
        assert a.shape == r.shape
        assert float((a - r).abs().max()) < 2e-6 * max(1.0, float(r.abs().max()))
    # exact properties of the fused noise: masked, symmetric, CoM-free
    x_new, e_new, x_mean, e_mean = outs[0][:4]
    assert float((x_new * (1 - d['node_mask'])).abs().max()) == 0.0
    assert float((e_new - e_new.permute(0, 2, 1, 3)).abs().max()) == 0.0
    assert torch.equal(x_mean, outs[1][2]) and torch.equal(e_mean, outs[1][3])


@pytest.mark.gpu
@pytest.mark.parametrize('cfg_name', ['qm9_uncond', 'qm9_cond'])
def test_graph_captured_chain_equals_eager_chain(cfg_name):
    """the reverse chain with one self-conditioned step captured into a CUDA graph and replayed
    (sampler.sampling(graph=True)) draws the same torch.randn stream and runs the same kernels as the eager loop, so
    the two chains agree bit for bit; the graphed loop issues two host calls per step."""
    import time
    from jodo_b200 import configs, synth
    from jodo_b200.model import MODELS
    cfg = configs.NAMED[cfg_name]()
    model = MODELS[cfg.model.name](cfg).cuda().eval()
    b = synth.make_batch(cfg, 48, seed=21)
    d = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()}
    ctx = None
    if cfg_name == 'qm9_cond':                                   # normalised property values, one per molecule
        ctx = torch.randn(48, int(cfg.model.cond_ch), generator=torch.Generator().manual_seed(5)).cuda()
    grid = torch.linspace(0.9946, 1e-3, 1000)[::40]            # 25 reverse steps
    res, dt = [], []
    for graph in (False, True):
        torch.manual_seed(1234)
        smp = S.AncestralSampler(S.CosineVP(), grid)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        x, e = smp.sampling(model, d['xh'], d['node_mask'], d['edge_mask'], d['edge_x'], ctx, graph=graph)
        torch.cuda.synchronize()
        dt.append(time.perf_counter() - t0)
        res.append((x, e))
    print(f'{cfg_name}: eager {dt[0] * 1e3:.1f} ms, graphed {dt[1] * 1e3:.1f} ms for {len(grid)} steps of 48 molecules')
    assert torch.isfinite(res[0][0]).all()
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])


def test_product_2d_sampler_replays_reference_chain():
    """AncestralSampler2D driving the 2-D oracle against the chain recorded from the reference's AncestralSampler_2D +
    DGT_concat_2D (tests/golden/moses_2d_chain.pt): replayed noise, then the same generator stream."""
    g, cfg = load_golden('moses_2d_chain')
    model = _oracle_model(golden_weights(g, cfg), cfg, torch.float32)
    b = g['inputs']
    noise = lambda i, kind: g['noise_node'][i] if kind == 'node' else g['noise_edge'][i]
    smp = S.AncestralSampler2D(S.CosineVP(), g['t'], noise_fn=noise, s_array=g['s'])
    xm, em = smp.sampling(model, b['xh'], b['node_mask'], b['edge_mask'], b['edge_x'])
    assert float((xm - g['x_mean']).abs().max()) < 2e-4 * float(g['x_mean'].abs().max())
    assert float((em - g['edge_x_mean']).abs().max()) < 2e-4 * float(g['edge_x_mean'].abs().max())
    # own draws: masked Gaussian node noise then symmetric edge noise per step, one generator (sampling.py:644-659)
    gen = torch.Generator().manual_seed(g['seed'])
    smp = S.AncestralSampler2D(S.CosineVP(), g['t'], generator=gen, s_array=g['s'])
    B, N, F_ = b['xh'].shape
    for i in range(len(g['t'])):
        assert torch.equal(smp._node_noise(B, N, F_, b['node_mask']), g['noise_node'][i])
        assert torch.equal(S.edge_noise(B, N, b['edge_x'].shape[-1], b['edge_mask'], gen), g['noise_edge'][i])


@pytest.mark.gpu
def test_cuda_2d_chain():
    """The same 2-D chain with the CUDA DGT_concat_2D behind the product sampler (replayed noise, 5 free-running
    steps: loose tolerance as for the 3-D chain, exact invariants)."""
    from jodo_b200.model import MODELS
    g, cfg = load_golden('moses_2d_chain')
    model = MODELS[cfg.model.name](cfg)
    model.load_state_dict(golden_weights(g, cfg), strict=True)
    model = model.cuda().eval()
    b = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in g['inputs'].items()}
    noise = lambda i, kind: (g['noise_node'][i] if kind == 'node' else g['noise_edge'][i]).cuda()
    smp = S.AncestralSampler2D(S.CosineVP(), g['t'], noise_fn=noise, s_array=g['s'])
    xm, em = smp.sampling(model, b['xh'], b['node_mask'], b['edge_mask'], b['edge_x'])
    xm, em = xm.cpu(), em.cpu()
    assert float((xm - g['x_mean']).abs().max()) < 2e-2 * float(g['x_mean'].abs().max())
    assert float((em - g['edge_x_mean']).abs().max()) < 2e-2 * float(g['edge_x_mean'].abs().max())
    nm = g['inputs']['node_mask']
    assert float((xm * (1 - nm)).abs().max()) == 0.0
    assert float((em - em.permute(0, 2, 1, 3)).abs().max()) == 0.0


@pytest.mark.gpu
def test_fused_dpm_outer_step_equals_torch_ops():
    """jodo_dpm_update (both updates of the order-2 singlestep, reference mix_dpm_solver.py:93-150 + :44-59) against the
    torch expressions of DPMSolverSinglestep._second_update on the same model outputs and the same normal draws."""
    from jodo_b200 import configs, synth
    from jodo_b200.model import MODELS
    cfg = configs.NAMED['qm9_cond']()
    model = MODELS[cfg.model.name](cfg).cuda().eval()
    b = synth.make_batch(cfg, 24, seed=5)
    d = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()}
    ctx = torch.randn(24, 1, generator=torch.Generator().manual_seed(3)).cuda()
    outs = {}
    for steps in (1, 3):
        for fused in (False, True):
            torch.manual_seed(99)
            sol = S.DPMSolverSinglestep(S.CosineVP(), 10, order=2)
            grid = sol.outer_grid('cuda')
            x, e = d['xh'], d['edge_x']
            for step in range(steps):
                x, e = sol.outer_step(model, step, grid, x, d['node_mask'], d['edge_mask'], e, ctx, fused=fused)
            outs[steps, fused] = (x, e)
    # one outer step: the two paths differ by the summation order of the CoM projection of the noise (1e-7), seen once
    # through the second model evaluation; three free-running steps amplify that through five more evaluations
    for steps, tol in ((1, 5e-6), (3, 1e-4)):
        (xa, ea), (xb, eb) = outs[steps, False], outs[steps, True]
        assert float((xa - xb).abs().max()) < tol * max(1.0, float(xa.abs().max())), steps
        assert float((ea - eb).abs().max()) < tol * max(1.0, float(ea.abs().max())), steps
        assert float(xb[..., :3].sum(1).abs().max()) < 1e-4


@pytest.mark.gpu
def test_graph_captured_dpm_chain_equals_eager_chain():
    """BASELINE configs[4] (QM9 conditional, DPM-Solver++ singlestep order 2): the chain with one outer step captured
    into a CUDA graph and replayed agrees bit for bit with the eager fused chain."""
    import time
    from jodo_b200 import configs, synth
    from jodo_b200.model import MODELS
    cfg = configs.NAMED['qm9_cond']()
    model = MODELS[cfg.model.name](cfg).cuda().eval()
    b = synth.make_batch(cfg, 48, seed=21)
    d = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()}
    ctx = torch.randn(48, 1, generator=torch.Generator().manual_seed(5)).cuda()
    res, dt = [], []
    for graph in (False, True):
        torch.manual_seed(1234)
        sol = S.DPMSolverSinglestep(S.CosineVP(), 20, order=2)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        x, e = sol.sampling(model, d['xh'], d['node_mask'], d['edge_mask'], d['edge_x'], ctx, graph=graph)
        torch.cuda.synchronize()
        dt.append(time.perf_counter() - t0)
        res.append((x, e))
        assert sol.n_evals == 20
    print(f'dpm chain: eager {dt[0] * 1e3:.1f} ms, graphed {dt[1] * 1e3:.1f} ms for 20 evaluations of 48 molecules')
    assert torch.isfinite(res[0][0]).all()
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])


@pytest.mark.gpu
def test_host_pipelined_steps_equal_in_line_copies():
    """sampler.HostPipelinedSteps (state in pinned host buffers, two half-batches, the copies of one half under the kernels of
    the other) against the same two captured steps replayed one after the other with the copies in line: the same replays in
    the same order (same torch.randn stream), so the host state agrees bit for bit after several steps; and the halves are
    the dealing of shard_molecules."""
    from jodo_b200 import configs, synth
    from jodo_b200.model import MODELS
    cfg = configs.NAMED['qm9_uncond']()
    model = MODELS[cfg.model.name](cfg).cuda().eval()
    B = 40
    b = synth.make_batch(cfg, B, seed=33)
    d = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()}
    N = d['node_mask'].shape[1]
    grid = torch.linspace(0.9946, 1e-3, 1000)[::100]
    parts = S.split_for_pipeline(b['n_nodes'], 2)
    assert sorted(torch.cat(parts).tolist()) == list(range(B)) and abs(len(parts[0]) - len(parts[1])) <= 1
    pin = lambda t: t.detach().cpu().contiguous().pin_memory()

    def build():
        torch.manual_seed(77)
        smp = S.AncestralSampler(S.CosineVP(), grid)
        steps, hosts = [], []
        for idx in parts:
            idx = idx.cuda()
            nm = d['node_mask'][idx].contiguous()
            em = d['edge_mask'].reshape(B, N, N)[idx].reshape(-1, 1).contiguous()
            x, ex = d['xh'][idx].contiguous(), d['edge_x'][idx].contiguous()
            cx = cex = None
            for i in range(2):                                 # first call, then the self-conditioned path (plan, workspaces)
                x, ex, _, _, cx, cex = smp.step(model, i, x, ex, nm, em, cx, cex)
            steps.append(S.GraphedAncestralStep(smp, model, x, ex, cx, cex, nm, em))
            hosts.append(dict(x=pin(x), ex=pin(ex), cx=pin(cx), cex=pin(cex)))
        return steps, hosts

    # in line: copy in, replay, copy out, part after part
    steps, hosts = build()
    for i in range(2, 6):
        for gs, h in zip(steps, hosts):
            for k, attr in S.HostPipelinedSteps.NAMES:
                getattr(gs, attr).copy_(h[k], non_blocking=True)
            gs.run(i)
            for k, attr in S.HostPipelinedSteps.NAMES:
                h[k].copy_(getattr(gs, attr), non_blocking=True)
            torch.cuda.synchronize()
    want = [{k: v.clone() for k, v in h.items()} for h in hosts]
    # pipelined
    steps, hosts = build()
    pipe = S.HostPipelinedSteps(steps, hosts)
    for i in range(2, 6):
        pipe.run(i)
    pipe.synchronize()
    for h, w_ in zip(hosts, want):
        for k in h:
            assert torch.isfinite(h[k]).all()
            assert torch.equal(h[k], w_[k]), k
    with pytest.raises(ValueError):
        S.HostPipelinedSteps(steps, [{k: v.clone() for k, v in h.items()} for h in hosts])      # unpinned host buffers


def test_split_for_pipeline_is_a_balanced_partition():
    """Host logic of the pipelined host-state API: the parts are a partition of the batch, equal in size up to one molecule
    and balanced in cost (sum of n (n - 1), the dealing of shard_molecules)."""
    g = torch.Generator().manual_seed(3)
    n = torch.randint(3, 30, (501,), generator=g)
    for parts in (2, 3):
        idx = S.split_for_pipeline(n, parts)
        assert sorted(torch.cat(idx).tolist()) == list(range(len(n)))
        sizes = [len(i) for i in idx]
        assert max(sizes) - min(sizes) <= 1
        cost = [float((n[i] * (n[i] - 1)).sum()) for i in idx]
        assert (max(cost) - min(cost)) / max(cost) < 0.02
