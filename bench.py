#!/usr/bin/env python
"""Benchmark of the DGT denoiser hot path: denoiser steps/sec (molecules x steps / s).

    python bench.py --gpus N --steps K --warmup W [--workload qm9|geom] [--impl reference]

A "step" is one reverse-SDE step of the reference's ancestral sampler on one batch of synthetic
molecules: one denoiser call (DGT forward, self-conditioned) + posterior-mean update + fresh noise
(reference sampling.py:535-589).  N=1 workload = BASELINE.json configs[1]: QM9 uncond architecture,
batch 2500, N<=29, first K steps of the 1000-step grid, random-init weights, synthetic inputs drawn the
way sampling_fn draws them.  N>1: every rank runs its own 2500 molecules (weak scaling), no collective
in the loop, one NCCL all_gather of the final samples after the timed region.

One JSON line on stdout (rank 0).  `value` = whole-job throughput with state resident in HBM;
`e2e` = same step driven through the public API with HOST (pinned) buffers, H2D of the step's inputs
and D2H of its results inside the timed region -- through sampler.HostPipelinedSteps (two half-batches of
independent molecules, the copies of one half under the kernels of the other; `e2e.in_line_value` = one
batch with the copies in line); `roofline` = the dominant kernel against the measured
peak; `cpu_baseline` = the CPU oracle port timed on this box's host cores on a bounded sample.
--impl reference times the reference's own CPU implementation on rank 0 only: the unmodified reference
modules from oracle/_ref/reference (staged by build(), git-ignored, travels with gpurun) -- kind "reference";
only if those are absent, the oracle port -- kind "port".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = 'denoiser steps/sec (molecules x steps / s)'
UNIT = 'mol-steps/s'
WORKLOADS = {
    # name -> (config factory name, per-GPU batch, max n, description)
    'qm9': ('qm9_uncond', 2500, None, 'QM9 uncond 1000-step ancestral sampling, batch 2500, N<=29'),
    'geom': ('geom_l8', 512, 80, 'GEOM-Drugs uncond medium (n_layers=8, nf=256), batch 512, N<=80'),
    'geom_l10': ('geom_l10', 512, 80, 'GEOM-Drugs uncond medium with the reference default n_layers=10 (nf=256), batch 512, N<=80'),
    'geom_large': ('geom_large', 512, 80, 'GEOM-Drugs uncond large (nf=384, n_layers=10), batch 512 per GPU, N<=80 '
                   '(BASELINE configs[3]: batch 4096 over 8 GPUs)'),
    'qm9_cond': ('qm9_cond', 2500, None, 'QM9 conditional single-property, 50-step DPM-Solver++ (singlestep, order 2), '
                 'batch 2500; a step = one model evaluation'),
}
CPU_SAMPLE = {'qm9': 64, 'geom': 16, 'geom_l10': 16, 'geom_large': 8, 'qm9_cond': 64}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--no-graph', action='store_true', help='time the eager step instead of the CUDA-graph replay')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='qm9', choices=sorted(WORKLOADS))
    ap.add_argument('--batch', type=int, default=None, help='per-GPU batch override (debug)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-pipeline', action='store_true', help='e2e with the host copies in line (one batch) instead of sampler.HostPipelinedSteps')
    ap.add_argument('--noise', default='torch', choices=['torch', 'philox'],
                    help="ancestral noise: the reference's torch.randn stream (default) or Philox drawn inside the update kernels")
    ap.add_argument('--no-extras', action='store_true', help='skip the strong-scaling point (N > 1) and the other workloads (N = 1)')
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=float(d['hbm_gbs']), tf_burst=float(d['bf16_tflops']),
                    tf_sustained=float(d.get('bf16_tflops_sustained', d['bf16_tflops'])), src='measured')
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src='fallback')


# ---- clocks ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '100', '-i', str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 6 or not f[0].isdigit():
                continue
            sm.append(int(f[0]))
            mx.append(int(f[1]))
            for nme, v in zip(names, f[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(nme)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ---- workload ----------------------------------------------------------------------------------------
def cpu_step_rate(cfg, wl, n_steps, threads):
    """CPU arm: the oracle port (fp32, all host threads) driving the same ancestral step on a bounded
    sample of the workload.  Returns (mol-steps/s, sample description)."""
    from jodo_b200 import sampler as S, synth
    from jodo_b200.params import param_spec, synth_state_dict
    from oracle.dgt_dense import dgt_forward
    torch.set_num_threads(threads)
    bs = CPU_SAMPLE[wl]
    _, _, max_n, _ = WORKLOADS[wl]
    b = synth.make_batch(cfg, bs, seed=42, max_n=max_n)
    sd = synth_state_dict(param_spec(cfg), seed=int(cfg.seed))

    def model(t, xh, node_mask, edge_mask, **kw):
        return dgt_forward(sd, cfg, t, xh, node_mask, edge_mask, **kw)

    if wl == 'qm9_cond':
        ctx = torch.randn(bs, 1, generator=torch.Generator().manual_seed(2))
        sol = S.DPMSolverSinglestep(S.CosineVP(), 50, order=2, generator=torch.Generator().manual_seed(1))
        grid = sol.outer_grid('cpu')
        x, ex = b['xh'], b['edge_x']
        n_outer = max(1, n_steps // 2)
        with torch.no_grad():
            x, ex = sol.outer_step(model, 0, grid, x, b['node_mask'], b['edge_mask'], ex, ctx)                  # warm-up
            t0 = time.perf_counter()
            for i in range(n_outer):
                x, ex = sol.outer_step(model, 1 + i, grid, x, b['node_mask'], b['edge_mask'], ex, ctx)
            dt = time.perf_counter() - t0
        return bs * 2 * n_outer / dt, (f'{bs} molecules (same histogram, seed 42) x {2 * n_outer} model evaluations of the '
                                       f'DPM-Solver chain, oracle port fp32')
    smp = S.AncestralSampler(S.CosineVP(), torch.linspace(0.9946, 1e-3, 1000), generator=torch.Generator().manual_seed(1))
    x, ex, cx, cex = b['xh'], b['edge_x'], None, None
    with torch.no_grad():
        x, ex, _, _, cx, cex = smp.step(model, 0, x, ex, b['node_mask'], b['edge_mask'], cx, cex)      # warm-up
        t0 = time.perf_counter()
        for i in range(n_steps):
            x, ex, _, _, cx, cex = smp.step(model, 1 + i, x, ex, b['node_mask'], b['edge_mask'], cx, cex)
        dt = time.perf_counter() - t0
    return bs * n_steps / dt, f'{bs} molecules (same histogram, seed 42) x {n_steps} ancestral steps, oracle port fp32'


REF_FILES = {'qm9': ('vpsde_qm9_uncond_jodo', {}), 'geom': ('vpsde_geom_uncond_jodo', {'n_layers': 8}),
             'geom_l10': ('vpsde_geom_uncond_jodo', {}), 'geom_large': ('vpsde_geom_uncond_jodo', {'nf': 384}),
             'qm9_cond': ('vpsde_qm9_cond_jodo', {})}


def reference_step_rate(cfg, wl, n_steps, threads):
    """CPU arm, kind "reference": the UNMODIFIED reference modules (models/mol_gnn.py, models/layers.py, sampling.py,
    mix_dpm_solver.py, diffusion/noise_schedule.py from /root/reference or its byte-identical staged copy under
    oracle/_ref/reference) through the PyG/scatter shim, fp32, all host threads, driving the reference's own sampler
    loop on a bounded sample of the workload.  Returns (mol-steps/s, sample description) or None when the reference
    sources are not present."""
    from oracle import ref_loader
    if not ref_loader.available():
        return None
    from jodo_b200 import synth
    from jodo_b200.params import param_spec, synth_state_dict
    torch.set_num_threads(threads)
    ref = ref_loader.load()
    fname, over = REF_FILES[wl]
    rcfg = ref_loader.load_config(fname)
    for k, v in over.items():
        rcfg.model[k] = v
    rcfg.device = torch.device('cpu')
    bs = CPU_SAMPLE[wl]
    _, _, max_n, _ = WORKLOADS[wl]
    b = synth.make_batch(cfg, bs, seed=42, max_n=max_n)
    model = ref.model_utils.create_model(rcfg)                # registry + DataParallel wrap (models/utils.py:24-28)
    sd = synth_state_dict(param_spec(cfg), seed=int(cfg.seed))
    model.load_state_dict({'module.' + k: v for k, v in sd.items()}, strict=True)
    if torch.cuda.is_available():
        # On a CPU-only host DataParallel calls the module directly; on a GPU box its constructor moves the module to
        # cuda:0 and scatters the inputs there.  The CPU arm calls the (unmodified) module itself, back on the CPU.
        model = model.module.to('cpu')
    model.eval()
    ns = ref.noise_schedule.NoiseScheduleVP(rcfg.sde.schedule, continuous_beta_0=rcfg.sde.continuous_beta_0,
                                            continuous_beta_1=rcfg.sde.continuous_beta_1)
    torch.manual_seed(1)
    src = 'staged copy oracle/_ref/reference' if ref_loader.is_staged_copy() else ref_loader.REF_ROOT
    with torch.no_grad():
        if wl == 'qm9_cond':
            ctx = torch.randn(bs, 1, generator=torch.Generator().manual_seed(2))
            n_eval = max(2, n_steps - n_steps % 2)
            rcfg.sampling.method = 'fast'
            rcfg.sampling.dpm_solver_method, rcfg.sampling.dpm_solver_order = 'singlestep_fixed', 2
            rcfg.sampling.steps = 2
            ref.mix_dpm_solver.DPM_Solver_hybrid(ns, rcfg).sampling(model, b['xh'], b['node_mask'], b['edge_mask'], b['edge_x'], ctx)
            rcfg.sampling.steps = n_eval
            sol = ref.mix_dpm_solver.DPM_Solver_hybrid(ns, rcfg)
            t0 = time.perf_counter()
            sol.sampling(model, b['xh'], b['node_mask'], b['edge_mask'], b['edge_x'], ctx)
            dt = time.perf_counter() - t0
            return bs * n_eval / dt, (f'{bs} molecules (same histogram, seed 42) x {n_eval} model evaluations of the reference '
                                      f'DPM_Solver_hybrid.sampling, unmodified reference fp32 ({src})')
        grid = torch.linspace(ns.T, 1e-3, 1000)
        mk = lambda ts: ref.sampling.AncestralSampler(ns, ts, rcfg.model.pred_data, rcfg.pred_edge, rcfg.model.self_cond,
                                                      ref.utils.get_self_cond_fn(rcfg))
        mk(grid[:2]).sampling(model, b['xh'], b['node_mask'], b['edge_mask'], b['edge_x'], None)          # warm-up
        smp = mk(grid[:n_steps])
        t0 = time.perf_counter()
        smp.sampling(model, b['xh'], b['node_mask'], b['edge_mask'], b['edge_x'], None)
        dt = time.perf_counter() - t0
    return bs * n_steps / dt, (f'{bs} molecules (same histogram, seed 42) x {n_steps} steps of the reference '
                               f'AncestralSampler.sampling, unmodified reference fp32 ({src})')


def cpu_arm(cfg, wl, n_steps, threads):
    """(rate, sample, kind): the unmodified reference when its sources travelled, else the oracle port."""
    r = reference_step_rate(cfg, wl, n_steps, threads)
    if r is not None:
        return r[0], r[1], 'reference'
    rate, sample = cpu_step_rate(cfg, wl, n_steps, threads)
    return rate, sample, 'port'


def run_reference(args):
    from jodo_b200 import configs
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cfg_name, batch, max_n, desc = WORKLOADS[args.workload]
    cfg = configs.NAMED[cfg_name]()
    threads = os.cpu_count() or 1
    n = max(1, args.steps)
    t0 = time.perf_counter()
    rate, sample, kind = cpu_arm(cfg, args.workload, n, threads)
    wall = time.perf_counter() - t0
    out = {
        'impl': 'reference', 'metric': METRIC, 'value': rate, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * CPU_SAMPLE[args.workload] / rate, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': desc, 'per_gpu_batch': batch},
        'cpu_baseline': {'value': rate, 'unit': UNIT, 'cores': threads, 'kind': kind, 'sample': sample},
        'e2e': {'value': rate, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'wall_s': wall,
    }
    print(json.dumps(out), flush=True)


class Workload:
    """One workload on this rank: model, synthetic state, the (graph-replayed) step function."""

    def __init__(self, wl, dev, world, rank, per_gpu_batch=None, global_batch=None, no_graph=False, noise='torch'):
        from jodo_b200 import configs, roofline, sampler as S, synth
        from jodo_b200.model import create_model
        self.S = S
        self.wl = wl
        cfg_name, batch, max_n, self.desc = WORKLOADS[wl]
        self.cfg_name = cfg_name
        batch = per_gpu_batch or batch
        self.cfg = cfg = configs.NAMED[cfg_name]()
        self.dpm = wl == 'qm9_cond'
        # ONE global batch (seed 42) dealt over the ranks by sampler.shard_molecules, so that the per-rank sum of n (n - 1)
        # -- the cost driver -- is balanced; each rank then draws its own noise for its molecules
        total = global_batch if global_batch is not None else batch * world
        gen = torch.Generator().manual_seed(42)
        n_all = synth.sample_n_nodes(cfg.data.info_name, total, gen, max_n)
        self.idx = S.shard_molecules(n_all, world, rank) if world > 1 else torch.arange(total)
        self.total, self.batch = total, len(self.idx)
        b = synth.make_batch(cfg, self.batch, seed=42 + rank, n_nodes=n_all[self.idx])
        self.model = create_model(cfg, dev)
        self.d = self.model.dims
        self.n_nodes = b['n_nodes']
        self.N = int(self.n_nodes.max())
        self.tot = roofline.batch_totals(self.n_nodes, self.d)
        self.node_mask, self.edge_mask = b['node_mask'].to(dev), b['edge_mask'].to(dev)
        self.state = dict(x=b['xh'].to(dev), ex=b['edge_x'].to(dev), cx=None, cex=None)
        torch.cuda.manual_seed(1234 + rank)                   # the samplers draw from the default CUDA generator
        self.grid = torch.linspace(0.9946, 1e-3, 1000)
        self.smp = S.AncestralSampler(S.CosineVP(), self.grid, noise=noise, seed=1234 + rank)
        if self.dpm:
            self.sol = S.DPMSolverSinglestep(S.CosineVP(), 50, order=2)
            self.ogrid = self.sol.outer_grid(dev)
            self.context = torch.randn(self.batch, 1, generator=torch.Generator().manual_seed(7 + rank)).to(dev)
        self.graphed, self.graph_note, self.no_graph = None, 'eager', no_graph

    def step(self, i):
        """Eager step i (ancestral: one reverse step; DPM: even i = one outer step of two evaluations, odd i = nothing)."""
        st = self.state
        if self.dpm:
            if i % 2:
                return
            x, ex = self.sol.outer_step(self.model, i // 2, self.ogrid, st['x'], self.node_mask, self.edge_mask, st['ex'], self.context)
            st.update(x=x, ex=ex, cx=self.sol.cond_x, cex=self.sol.cond_edge_x, xm=x, em=ex)
            return
        x, ex, xm, em, cx, cex = self.smp.step(self.model, i, st['x'], st['ex'], self.node_mask, self.edge_mask, st['cx'], st['cex'])
        st.update(x=x, ex=ex, cx=cx, cex=cex, xm=xm, em=em)

    def warm(self, W):
        """W eager warm-up steps, then capture the step into a CUDA graph (same kernels, same random stream; removes the
        host launch gaps between the ~125 kernels of a step).  Falls back to the eager step if the capture fails."""
        S = self.S
        for i in range(W):
            self.step(i)
        if self.no_graph or W < 2:
            return
        st = self.state
        try:
            if self.dpm:
                self.graphed = S.GraphedDPMStep(self.sol, self.model, st['x'], st['ex'], self.node_mask, self.edge_mask,
                                                self.context, self.ogrid)
                self.graphed.run(W // 2 - 1)                   # one untimed replay
            else:
                self.graphed = S.GraphedAncestralStep(self.smp, self.model, st['x'], st['ex'], st['cx'], st['cex'],
                                                      self.node_mask, self.edge_mask)
                self.graphed.run(W - 1)
            self.graph_note = 'cuda graph replay'
        except Exception as exc:                               # noqa: BLE001
            self.graphed, self.graph_note = None, 'eager (graph capture failed: %s)' % str(exc)[:120]

    def run(self, i):
        if self.graphed is None:
            return self.step(i)
        if self.dpm:
            if i % 2 == 0:
                self.graphed.run(i // 2)
        else:
            self.graphed.run(i)

    def sync_state(self):
        g = self.graphed
        if g is None:
            return
        if self.dpm:
            self.state.update(x=g.x, ex=g.edge_x, cx=g.cond_x, cex=g.cond_edge_x, xm=g.x, em=g.edge_x)
        else:
            self.state.update(x=g.x, ex=g.edge_x, cx=g.cond_x, cex=g.cond_edge_x, xm=g.x_mean, em=g.edge_mean)

    def launches_per_step(self):
        return None if self.graphed is None else self.graphed.launches_per_step / (2 if self.dpm else 1)


def timed_steps(w, W, K, barrier, dev, prof=False):
    """K steps after W, device-timed; returns this rank's milliseconds."""
    from jodo_b200 import _lib
    barrier()
    l0 = _lib.LAUNCHES
    if prof:
        torch.cuda.cudart().cudaProfilerStart()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(W, W + K):
        w.run(i)
    e1.record()
    barrier()
    if prof:
        torch.cuda.cudart().cudaProfilerStop()
    launches = _lib.LAUNCHES - l0
    if w.graphed is not None:
        launches = int(w.launches_per_step() * K)            # replays do not pass through the ctypes binding
    w.sync_state()
    return e0.elapsed_time(e1), launches


def run_b200(args):
    import torch.distributed as dist
    from jodo_b200 import _lib, roofline, sampler as S
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the DGT hot path has no CPU fallback; use --impl reference)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        # NCCL prints its version banner on stdout while the communicator comes up (the image sets NCCL_DEBUG=VERSION);
        # stdout must carry one JSON line, so file descriptor 1 points at stderr during the initialisation
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def rank_stats(ms_local):
        """(max, min) over ranks of a per-rank device time."""
        t = torch.tensor([ms_local, -ms_local], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), -float(t[1])

    K, W = args.steps, args.warmup
    w = Workload(args.workload, dev, world, rank, per_gpu_batch=args.batch, no_graph=args.no_graph, noise=args.noise)
    dpm, batch, d, tot, N, cfg, total_mols = w.dpm, w.batch, w.d, w.tot, w.N, w.cfg, w.total
    if dpm:
        # a step = one model evaluation; the solver advances in outer steps of two evaluations (order 2)
        if K % 2 or W % 2:
            K, W = K + (K % 2), W + (W % 2)
        if W + K > 50:
            raise SystemExit('qm9_cond: warmup + steps must not exceed the 50 model evaluations of the chain')
    elif W + K > len(w.grid):
        raise SystemExit('warmup + steps must not exceed the 1000-step grid')
    w.warm(W)
    state = w.state

    # ---- device-resident run ------------------------------------------------------------------------
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms_local, launches = timed_steps(w, W, K, barrier, dev, prof=os.environ.get('JODO_CUDA_PROFILER') == '1')
    clk = clocks.stop() if rank == 0 else None
    ms, ms_min = rank_stats(ms_local)
    graphed, graph_note = w.graphed, w.graph_note
    ok = bool(torch.isfinite(state['xm']).all()) and bool(torch.isfinite(state['em']).all())
    model, node_mask, edge_mask = w.model, w.node_mask, w.edge_mask

    # ---- end to end: host buffers, H2D + D2H every step ------------------------------------------------
    e2e = None
    if not args.no_e2e:
        pin = lambda t: t.detach().cpu().contiguous().pin_memory()
        host = dict(x=pin(state['x']), ex=pin(state['ex']), cx=pin(state['cx']), cex=pin(state['cex']))
        h2d = sum(v.numel() * 4 for v in host.values())
        outh = {k: torch.empty_like(v).pin_memory() for k, v in host.items()}
        d2h = sum(v.numel() * 4 for v in outh.values())
        names = dict(x='x', ex='edge_x', cx='cond_x', cex='cond_edge_x')

        def e2e_step(i):
            if graphed is not None:                      # host buffers -> the graph's static inputs -> replay -> host
                for k, attr in names.items():
                    getattr(graphed, attr).copy_(host[k], non_blocking=True)
                graphed.run(min(i // 2, 23) if dpm else i)
                for k, attr in names.items():
                    outh[k].copy_(getattr(graphed, attr), non_blocking=True)
            else:
                dv = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
                if dpm:
                    w.sol.cond_x, w.sol.cond_edge_x = dv['cx'], dv['cex']
                    x, ex = w.sol.outer_step(model, (i // 2) % 24, w.ogrid, dv['x'], node_mask, edge_mask, dv['ex'], w.context)
                    cx, cex = w.sol.cond_x, w.sol.cond_edge_x
                else:
                    x, ex, _, _, cx, cex = w.smp.step(model, i, dv['x'], dv['ex'], node_mask, edge_mask, dv['cx'], dv['cex'])
                for k, v in dict(x=x, ex=ex, cx=cx, cex=cex).items():
                    outh[k].copy_(v, non_blocking=True)
            torch.cuda.current_stream().synchronize()           # the caller reads the result on the host
            for k in host:
                host[k], outh[k] = outh[k], host[k]

        ke = max(3, min(K, 20))
        for i in range(2):
            e2e_step(W + K - 1)
        barrier()
        t0 = time.perf_counter()
        for i in range(ke):
            e2e_step(W + K - 1)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.barrier()
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        evals = 2 if dpm else 1                         # model evaluations per e2e_step
        e2e = {'value': w.total * ke * evals / float(dt), 'unit': UNIT, 'h2d_bytes_per_step': h2d // evals,
               'd2h_bytes_per_step': d2h // evals, 'steps': ke * evals, 'copies': 'in line'}
        # The same state and the same bytes through the public pipelined API (sampler.HostPipelinedSteps): the batch as two
        # halves of independent molecules, each with its own captured step and pinned host state; the copies of one half
        # run under the kernels of the other.  Every step still moves every input host -> device and every result back.
        if graphed is not None and not args.no_pipeline and batch >= 2:
            S = w.S
            N_, ch_ = node_mask.shape[1], state['ex'].shape[-1]
            em3 = edge_mask.reshape(batch, N_, N_)
            steps_p, hosts_p = [], []
            for idx in S.split_for_pipeline(w.n_nodes, 2):
                idx = idx.to(dev)
                nm = node_mask[idx].contiguous()
                em = em3[idx].reshape(-1, 1).contiguous()
                sp = {k: state[k][idx].contiguous() for k in ('x', 'ex', 'cx', 'cex')}
                # the model runs the self-conditioned path once on these masks (plan, workspaces) before the capture
                if dpm:
                    sol_p = S.DPMSolverSinglestep(S.CosineVP(), 50, order=2)
                    ctx_p = w.context[idx].contiguous()
                    sol_p.cond_x, sol_p.cond_edge_x = sp['cx'], sp['cex']
                    sol_p.outer_step(model, 1, w.ogrid, sp['x'], nm, em, sp['ex'], ctx_p)
                    steps_p.append(S.GraphedDPMStep(sol_p, model, sp['x'], sp['ex'], nm, em, ctx_p, w.ogrid))
                else:
                    w.smp.step(model, W + K - 1, sp['x'], sp['ex'], nm, em, sp['cx'], sp['cex'])
                    steps_p.append(S.GraphedAncestralStep(w.smp, model, sp['x'], sp['ex'], sp['cx'], sp['cex'], nm, em))
                hosts_p.append({k: pin(v) for k, v in sp.items()})
            pipe = S.HostPipelinedSteps(steps_p, hosts_p)
            si = min((W + K - 1) // 2, 23) if dpm else W + K - 1          # DPM: an outer step of two evaluations
            for i in range(2):
                pipe.run(si)
            pipe.synchronize()
            barrier()
            t0 = time.perf_counter()
            for i in range(ke):
                pipe.run(si)
            pipe.synchronize()
            torch.cuda.synchronize()
            dtp = torch.tensor([time.perf_counter() - t0], device=dev)
            if world > 1:
                dist.barrier()
                dist.all_reduce(dtp, op=dist.ReduceOp.MAX)
            ok = ok and all(bool(torch.isfinite(h[k]).all()) for h in hosts_p for k in h)
            e2e.update(in_line_value=e2e['value'], value=w.total * ke * evals / float(dtp), h2d_bytes_per_step=pipe.bytes_per_step // evals,
                       d2h_bytes_per_step=pipe.bytes_per_step // evals,
                       copies='pipelined: two half-batches, the H2D / D2H of one half under the kernels of the other '
                              '(sampler.HostPipelinedSteps); in_line_value = one batch with the copies in line')
            del pipe, steps_p, hosts_p

    # ---- per-kernel device times (CUDA events around every C-ABI call, outside the timed region) -------
    _lib.TRACE = []
    reps = 3
    for i in range(reps):
        w.step(W + K - (2 if dpm else 1))
    torch.cuda.synchronize()
    per = {}
    for name, a, z in _lib.TRACE:
        t, c = per.get(name, (0.0, 0))
        per[name] = (t + a.elapsed_time(z), c + 1)
    _lib.TRACE = None
    total_traced = sum(t for t, _ in per.values())
    top = max(per, key=lambda k: per[k][0])
    pk = peaks()
    kf = roofline.per_edge_kernel_flops(d)
    kb = roofline.per_edge_kernel_bytes(d)
    kh = {}                                             # HBM-bound row kernels of the wide path
    if getattr(model, 'wide', False):
        kf, kb, kh = roofline.wide_kernel_flops(d), {}, roofline.wide_kernel_bytes(d)
    kernels = {}
    # kernels that own the symmetric edge state run once per UNORDERED pair: half the rows of the directed-edge count
    rows_of = lambda name: tot['edges'] * (0.5 if name in roofline.PAIR_KERNELS else 1.0)
    for name, (t, c) in sorted(per.items(), key=lambda kv: -kv[1][0]):
        avg_ms = t / c
        ent = {'launches_per_step': c // reps // (2 if dpm else 1), 'avg_ms': round(avg_ms, 4), 'share': round(t / total_traced, 4)}
        if name in kf:
            ent['tflops'] = round(kf[name] * rows_of(name) / (avg_ms * 1e-3) / 1e12, 2)
        if name in kb or name in kh:
            ent['gbs'] = round((kb.get(name) or kh[name]) * rows_of(name) / (avg_ms * 1e-3) / 1e9, 1)
        if name in roofline.PAIR_KERNELS:
            ent['rows'] = 'pairs'
        kernels[name] = ent
    traffic = None
    tp = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f).get(args.workload, {}).get(top)
    roof = None
    if top in kf:
        ach = kf[top] * rows_of(top) / (per[top][0] / per[top][1] * 1e-3) / 1e12
        roof = {'kernel': top, 'bound': 'tensor', 'achieved': ach, 'peak': pk['tf_sustained'], 'unit': 'TFLOP/s',
                'frac': ach / pk['tf_sustained'], 'traffic': traffic,
                'peak_source': pk['src'] + ' bf16 dense sustained (kernels run kind::f16, same nominal rate)',
                'share_of_step': per[top][0] / total_traced}
    elif top in kh:
        ach = kh[top] * rows_of(top) / (per[top][0] / per[top][1] * 1e-3) / 1e9
        roof = {'kernel': top, 'bound': 'hbm', 'achieved': ach, 'peak': pk['hbm'], 'unit': 'GB/s', 'frac': ach / pk['hbm'],
                'traffic': traffic, 'peak_source': pk['src'] + ' HBM copy bandwidth', 'share_of_step': per[top][0] / total_traced}
    whole = {'tflops': tot['flops'] * K / (ms_local * 1e-3) / 1e12, 'hbm_gbs_alg': tot['bytes'] * K / (ms_local * 1e-3) / 1e9}
    whole['tensor_frac'] = whole['tflops'] / pk['tf_sustained']
    whole['hbm_frac'] = whole['hbm_gbs_alg'] / pk['hbm']

    # ---- the one collective: gather the final samples (timed on the device, max over ranks) -----------
    gathered, gather_ms = None, None
    if world > 1:
        Ng = torch.tensor([N], device=dev)
        dist.all_reduce(Ng, op=dist.ReduceOp.MAX)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        gx, ge = S.gather_samples(state['xm'], state['em'], w.idx, w.total, int(Ng))
        g1.record()
        torch.cuda.synchronize()
        gather_ms = rank_stats(g0.elapsed_time(g1))[0]
        gathered = [list(gx.shape), list(ge.shape)]
        del gx, ge

    # ---- strong-scaling point: the single-GPU batch dealt over all ranks ------------------------------
    strong = None
    if world > 1 and not args.no_extras:
        total1 = WORKLOADS[args.workload][1]
        del w, graphed, model, state
        torch.cuda.empty_cache()
        ws_ = Workload(args.workload, dev, world, rank, global_batch=total1, no_graph=args.no_graph)
        ws_.warm(W)
        t_loc, _ = timed_steps(ws_, W, K, barrier, dev)
        t_max, t_min = rank_stats(t_loc)
        strong = {'global_batch': total1, 'per_rank_batch': ws_.batch, 'ms_per_step': t_max / K, 'value': total1 * K / (t_max * 1e-3),
                  'rank_ms_per_step_min': t_min / K, 'rank_ms_per_step_max': t_max / K, 'step_launch': ws_.graph_note}
        del ws_
        torch.cuda.empty_cache()

    # ---- the other BASELINE configs, short runs on one GPU (configs[2..4]) -----------------------------
    others = None
    if world == 1 and rank == 0 and not args.no_extras and args.workload == 'qm9':
        others = {}
        try:
            del w, graphed, model, state
        except NameError:
            pass
        for name, (k2, w2) in (('geom', (10, 4)), ('geom_large', (6, 4)), ('qm9_cond', (10, 4))):
            torch.cuda.empty_cache()
            try:
                wo = Workload(name, dev, 1, 0)
                wo.warm(w2)
                t, nl = timed_steps(wo, w2, k2, barrier, dev)
                others[name] = {'workload': wo.desc, 'ms_per_step': t / k2, 'value': wo.total * k2 / (t * 1e-3), 'steps': k2, 'warmup': w2,
                                'step_launch': wo.graph_note, 'gpu_launches': nl,
                                'tensor_frac': wo.tot['flops'] * k2 / (t * 1e-3) / 1e12 / pk['tf_sustained']}
                del wo
            except Exception as exc:                       # noqa: BLE001
                others[name] = {'error': str(exc)[:200]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        rate, sample, kind = cpu_arm(cfg, args.workload, 24 if args.workload == 'qm9' else 40, threads)
        cpu = {'value': rate, 'unit': UNIT, 'cores': threads, 'kind': kind, 'sample': sample}

    if rank == 0:
        out = {
            'metric': METRIC, 'value': total_mols * K / (ms * 1e-3), 'unit': UNIT, 'n_gpus': world, 'steps': K,
            'warmup': W, 'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f16 operands (tf32 mantissa) / f32 accumulate + f32 elementwise', 'data': 'synthetic',
            'config': {'workload': WORKLOADS[args.workload][3], 'global_batch': total_mols, 'per_gpu_batch': batch, 'N': N, 'atoms': tot['atoms'], 'edges': tot['edges'],
                       'arch': WORKLOADS[args.workload][0], 'weights': 'random init (seed 42)', 'l2': 'inputs larger than L2 '
                       '(edge state %.0f MB per step)' % (tot['bytes'] / 1e6),
                       'parallelism': f'dp{world} (independent molecules, one global batch dealt by size: sampler.shard_molecules)',
                       'step_launch': graph_note,
                       'noise': 'torch.randn in the reference call order' if args.noise == 'torch' or dpm else 'Philox4x32-10 inside the update kernels'},
            'e2e': e2e, 'gpu_launches': launches, 'clocks': clk, 'roofline': roof, 'whole_step': whole,
            'rank_ms_per_step_min': ms_min / K, 'rank_ms_per_step_max': ms / K, 'gather_ms': gather_ms, 'strong': strong,
            'workloads': others, 'kernels': kernels, 'cpu_baseline': cpu, 'finite': ok, 'gathered': gathered,
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
