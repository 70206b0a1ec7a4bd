"""Host side of the reverse-SDE loop around the denoiser: the reference's ancestral sampler, its
VP cosine schedule, and the multi-GPU sharding of a batch of independent molecules.

This mirrors the caller of the hot path, so that the bench and the tests can
drive the denoiser exactly the way ``sampling_fn`` does without the reference tree being present:

* ``CosineVP``            - NoiseScheduleVP('cosine'), reference diffusion/noise_schedule.py:43-52,76-92
* ``AncestralSampler``    - reference sampling.py:518-596 (pred_edge=True, self_cond=True, model_pred_data=True)
* ``node_noise/edge_noise`` - reference models/utils.py:67-99 (same torch.randn call order, so on one
  device with one seed the stream is the one the reference would draw)
* ``DPMSolverSinglestep`` - reference mix_dpm_solver.py:16-59,93-150,285-335 (DPM-Solver++ 'singlestep_fixed',
  order 1 or 2, atom / bond features by the solver update, positions by the ancestral update; the driver of
  conditional QM9 sampling, BASELINE config 5)
* ``shard_molecules`` / ``gather_samples`` - one process per GPU, molecules dealt so that sum n(n-1) is
  balanced, no collective inside the loop, one all_gather of the final samples
"""
from __future__ import annotations

import math

import torch


class CosineVP:
    """Continuous-time VP schedule, cosine variant (reference diffusion/noise_schedule.py)."""

    def __init__(self):
        self.cosine_s = 0.008
        self.cosine_log_alpha_0 = math.log(math.cos(self.cosine_s / (1. + self.cosine_s) * math.pi / 2.))
        self.T = 0.9946
        self.total_N = 1000

    def marginal_log_mean_coeff(self, t):
        return torch.log(torch.cos((t + self.cosine_s) / (1. + self.cosine_s) * math.pi / 2.)) - self.cosine_log_alpha_0

    def marginal_prob(self, t):
        lm = self.marginal_log_mean_coeff(t)
        return torch.exp(lm), torch.sqrt(1. - torch.exp(2. * lm))

    def marginal_std(self, t):
        return torch.sqrt(1. - torch.exp(2. * self.marginal_log_mean_coeff(t)))

    def marginal_lambda(self, t):
        """half log-SNR (reference diffusion/noise_schedule.py:93-99)"""
        lm = self.marginal_log_mean_coeff(t)
        return lm - 0.5 * torch.log(1. - torch.exp(2. * lm))

    def inverse_lambda(self, lamb):
        """reference diffusion/noise_schedule.py:113-117 (cosine branch)"""
        log_alpha = -0.5 * torch.logaddexp(-2. * lamb, torch.zeros((1,)).to(lamb))
        return torch.arccos(torch.exp(log_alpha + self.cosine_log_alpha_0)) * 2. * (1. + self.cosine_s) / math.pi - self.cosine_s

    def get_noise_level(self, t):
        """reference diffusion/noise_schedule.py:119-122"""
        alpha_t = torch.exp(self.marginal_log_mean_coeff(t))
        sigma_t = self.marginal_std(t)
        return torch.log(alpha_t ** 2 / sigma_t ** 2)


def remove_mean_with_mask(x, node_mask):
    """reference models/utils.py:38-45"""
    n = node_mask.sum(1, keepdim=True)
    return x - (x.sum(1, keepdim=True) / n) * node_mask


def node_noise(B, N, feat, node_mask, generator=None):
    """sample_combined_position_feature_noise, reference models/utils.py:83-90"""
    dev = node_mask.device
    zx = torch.randn((B, N, 3), device=dev, generator=generator) * node_mask
    zx = remove_mean_with_mask(zx, node_mask)
    zh = torch.randn((B, N, feat), device=dev, generator=generator) * node_mask
    return torch.cat([zx, zh], dim=2)


def edge_noise(B, N, ch, edge_mask, generator=None):
    """sample_symmetric_edge_feature_noise, reference models/utils.py:93-99"""
    z = torch.randn((B, ch, N, N), device=edge_mask.device, generator=generator)
    z = torch.tril(z, -1)
    z = z + z.transpose(-1, -2)
    return z.permute(0, 2, 3, 1) * edge_mask.reshape(B, N, N, 1)


def ancestral_coefficients(schedule, time_steps, s_array=None):
    """Per-step scalars of the ancestral update, evaluated in fp32 with the reference's operation
    order (sampling.py:536-549).  Returns a CPU tensor [steps, 4]: (c_x, c_pred, sigma, noise_level).
    s_array: the "next" time of every step; default = the grid shifted by one with 0 appended
    (sampling.py:523)."""
    t = time_steps.float().cpu()
    s = torch.cat([t[1:], torch.zeros(1)]) if s_array is None else s_array.float().cpu()
    alpha_t, sigma_t = schedule.marginal_prob(t)
    alpha_s, sigma_s = schedule.marginal_prob(s)
    alpha_ts = alpha_t / alpha_s
    sigma2_ts = sigma_t ** 2 - alpha_ts ** 2 * sigma_s ** 2
    sigma = torch.sqrt(sigma2_ts) * sigma_s / sigma_t
    c_x = alpha_ts * sigma_s ** 2 / sigma_t ** 2
    c_pred = alpha_s * sigma2_ts / sigma_t ** 2
    nl = torch.log(alpha_t ** 2 / sigma_t ** 2)
    return torch.stack([c_x, c_pred, sigma, nl], dim=1)


class AncestralSampler:
    """Ancestral sampling for joint 2D & 3D generation (reference sampling.py:518-596) with
    self-conditioning ('ori' hand-off, reference utils.py:134-136)."""

    def __init__(self, schedule, time_steps, generator=None, noise_fn=None, s_array=None, fused=True, noise='torch', seed=0):
        """noise='torch': the reference's torch.randn draws in the reference's call order (models/utils.py:67-99);
        noise='philox': the draws are generated inside the fused update kernels (Philox4x32-10 keyed by `seed`, counter =
        element / step) -- a different random stream with its own parity chain (tests/test_philox.py), no generator
        launches and no raw-draw buffers."""
        assert noise in ('torch', 'philox')
        self.noise, self.seed = noise, int(seed)
        self.fused = fused                # CUDA tensors: fused update kernel (same random stream as the torch ops)
        self.schedule = schedule
        self.t_array = time_steps
        self.coef = ancestral_coefficients(schedule, time_steps, s_array)
        self.generator = generator
        self.noise_fn = noise_fn          # optional (step, kind, shape...) -> tensor, for replayed noise

    @torch.no_grad()
    def step(self, model, i, x, edge_x, node_mask, edge_mask, cond_x, cond_edge_x, context=None):
        """One reverse step: denoiser call + posterior-mean update + fresh noise.
        Returns (x, edge_x, x_mean, edge_x_mean, pred_x, pred_edge)."""
        bs, N = x.shape[0], x.shape[1]
        c_x, c_pred, sigma, nl = (float(v) for v in self.coef[i])
        vec_t = torch.full((bs,), float(self.t_array[i]), device=x.device)
        noise_level = torch.full((bs,), nl, device=x.device)
        pred, edge_pred = model(vec_t, x, node_mask, edge_mask, edge_x=edge_x, noise_level=noise_level,
                                cond_x=cond_x, cond_edge_x=cond_edge_x, context=context)
        if x.is_cuda and self.noise_fn is None and self.fused:
            return self._fused_update(x, edge_x, pred, edge_pred, node_mask, edge_mask, c_x, c_pred, sigma, step=i)
        if self.noise == 'philox':
            raise ValueError("noise='philox' draws inside the fused CUDA update: it needs CUDA tensors, fused=True and no noise_fn")
        x_mean = c_x * x + c_pred * pred
        if self.noise_fn is not None:
            zn, ze = self.noise_fn(i, 'node'), self.noise_fn(i, 'edge')
        else:
            zn = self._node_noise(bs, N, x.shape[2], node_mask)
            ze = None
        x_new = x_mean + sigma * zn
        edge_mean = c_x * edge_x + c_pred * edge_pred
        if ze is None:
            ze = edge_noise(bs, N, edge_x.shape[-1], edge_mask, self.generator)
        edge_new = edge_mean + sigma * ze
        return x_new, edge_new, x_mean, edge_mean, pred, edge_pred

    def _node_noise(self, bs, N, F_, node_mask):
        return node_noise(bs, N, F_ - 3, node_mask, self.generator)

    def _fused_update(self, x, edge_x, pred, edge_pred, node_mask, edge_mask, c_x, c_pred, sigma, coef_dev=None, step=0):
        """Posterior mean + noise in two launches of libjodo_b200 (jodo_ancestral_update) instead of ~30 torch
        kernels; the raw normal draws are the same torch.randn calls, in the same order, as node_noise / edge_noise."""
        import ctypes
        from . import _lib
        bs, N, F_ = x.shape
        ch = edge_x.shape[-1]
        dev = x.device
        c32 = lambda t: t.contiguous().float()
        x, edge_x, pred, edge_pred = c32(x), c32(edge_x), c32(pred), c32(edge_pred)
        nm, em = c32(node_mask), c32(edge_mask)
        x_new, x_mean = torch.empty_like(x), torch.empty_like(x)
        e_new, e_mean = torch.empty_like(edge_x), torch.empty_like(edge_x)
        f = ctypes.c_float
        if self.noise == 'philox':
            _lib.call('jodo_ancestral_update_philox', _lib.ptr(x), _lib.ptr(pred), _lib.ptr(nm), _lib.ptr(edge_x),
                      _lib.ptr(edge_pred), _lib.ptr(em), ctypes.c_int(bs), ctypes.c_int(N), ctypes.c_int(F_), ctypes.c_int(ch),
                      f(c_x), f(c_pred), f(sigma), _lib.ptr(coef_dev), ctypes.c_ulonglong(self.seed), ctypes.c_uint(step),
                      _lib.ptr(x_new), _lib.ptr(x_mean), _lib.ptr(e_new), _lib.ptr(e_mean), _lib.stream_ptr())
            return x_new, e_new, x_mean, e_mean, pred, edge_pred
        raw_pos = torch.randn((bs, N, 3), device=dev, generator=self.generator)
        raw_feat = torch.randn((bs, N, F_ - 3), device=dev, generator=self.generator)
        raw_edge = torch.randn((bs, ch, N, N), device=dev, generator=self.generator)
        _lib.call('jodo_ancestral_update', _lib.ptr(x), _lib.ptr(pred), _lib.ptr(raw_pos), _lib.ptr(raw_feat), _lib.ptr(nm),
                  _lib.ptr(edge_x), _lib.ptr(edge_pred), _lib.ptr(raw_edge), _lib.ptr(em), ctypes.c_int(bs), ctypes.c_int(N),
                  ctypes.c_int(F_), ctypes.c_int(ch), f(c_x), f(c_pred), f(sigma), _lib.ptr(coef_dev), _lib.ptr(x_new), _lib.ptr(x_mean),
                  _lib.ptr(e_new), _lib.ptr(e_mean), _lib.stream_ptr())
        return x_new, e_new, x_mean, e_mean, pred, edge_pred

    @torch.no_grad()
    def sampling(self, model, z_T, node_mask, edge_mask, edge_z_T, context=None, graph=False):
        """The whole reverse chain (reference sampling.py:530-596).  graph=True captures one self-conditioned reverse
        step (denoiser call + noise draws + fused update + hand-off copies) into a CUDA graph after two eager steps and
        replays it for the rest of the chain: per step the host issues one 16-byte coefficient copy and one graph
        launch instead of ~160 kernel launches, which is what bounds small batches.  Same kernels, same torch.randn
        stream as the eager chain."""
        if graph:
            return self._sampling_graphed(model, z_T, node_mask, edge_mask, edge_z_T, context)
        x, edge_x = z_T, edge_z_T
        cond_x = cond_edge_x = None
        x_mean = edge_mean = None
        for i in range(len(self.t_array)):
            x, edge_x, x_mean, edge_mean, cond_x, cond_edge_x = self.step(
                model, i, x, edge_x, node_mask, edge_mask, cond_x, cond_edge_x, context)
        return x_mean, edge_mean

    def _sampling_graphed(self, model, z_T, node_mask, edge_mask, edge_z_T, context):
        n = len(self.t_array)
        c32 = lambda t: t.contiguous().float()
        x, edge_x = c32(z_T), c32(edge_z_T)
        cond_x = cond_edge_x = None
        x_mean = edge_mean = None
        eager = min(2, n)                       # step 0 takes the first-call path (no self-conditioning); step 1 warms
        for i in range(eager):                  # up every workspace of the self-conditioned path
            x, edge_x, x_mean, edge_mean, cond_x, cond_edge_x = self.step(
                model, i, x, edge_x, node_mask, edge_mask, cond_x, cond_edge_x, context)
        if n <= eager:
            return x_mean, edge_mean
        gs = GraphedAncestralStep(self, model, x, edge_x, cond_x, cond_edge_x, node_mask, edge_mask, context)
        for i in range(eager, n):
            gs.run(i)
        return gs.x_mean.clone(), gs.edge_mean.clone()


class GraphedAncestralStep:
    """One self-conditioned reverse step of an AncestralSampler (denoiser call + noise draws + fused update + hand-off
    copies) captured into a CUDA graph.  ``run(i)`` writes the coefficients of step i into device memory and replays
    the graph; the state lives in the static buffers x, edge_x, cond_x, cond_edge_x (inputs of the next step) and
    x_mean, edge_mean.  The model must have run the self-conditioned path once with these masks before (workspaces and
    the plan are created outside the capture)."""

    def __init__(self, sampler, model, x, edge_x, cond_x, cond_edge_x, node_mask, edge_mask, context=None):
        if not (x.is_cuda and sampler.fused and sampler.noise_fn is None and (sampler.generator is None or sampler.noise == 'philox')):
            raise ValueError('the graph-captured step needs CUDA tensors, the fused update and the default CUDA generator '
                             "(or noise='philox')")
        if cond_x is None:
            raise ValueError('capture the self-conditioned step: run the first reverse step eagerly')
        from . import _lib
        c32 = lambda t: t.contiguous().float()
        dev, bs = x.device, x.shape[0]
        n = len(sampler.t_array)
        self.table = torch.tensor([[float(v) for v in sampler.coef[i]] + [float(i)] for i in range(n)], device=dev, dtype=torch.float32)
        self.coef = torch.zeros(5, device=dev, dtype=torch.float32)       # {c_x, c_pred, sigma, noise level, step} of the step
        self.x, self.edge_x = c32(x).clone(), c32(edge_x).clone()
        self.cond_x, self.cond_edge_x = c32(cond_x).clone(), c32(cond_edge_x).clone()
        vec_t = torch.zeros(bs, device=dev)                               # ignored by the model (reference mol_gnn.py:534)
        ctx = None if context is None else c32(context).clone()
        self._static = (vec_t, ctx)     # read by the captured kernels on every replay: must outlive this constructor

        def one_step():
            nl = self.coef[3:4].expand(bs)
            pred, edge_pred = model(vec_t, self.x, node_mask, edge_mask, edge_x=self.edge_x, noise_level=nl,
                                    cond_x=self.cond_x, cond_edge_x=self.cond_edge_x, context=ctx)
            out = sampler._fused_update(self.x, self.edge_x, pred, edge_pred, node_mask, edge_mask, 0.0, 0.0, 0.0,
                                        coef_dev=self.coef)
            self.x.copy_(out[0]); self.edge_x.copy_(out[1]); self.cond_x.copy_(pred); self.cond_edge_x.copy_(edge_pred)
            return out[2], out[3]

        self.graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        l0 = _lib.LAUNCHES
        with torch.cuda.stream(side):           # capture needs a non-default stream; nothing runs during capture
            with torch.cuda.graph(self.graph, stream=side):
                self.x_mean, self.edge_mean = one_step()
        torch.cuda.current_stream().wait_stream(side)
        self.launches_per_step = _lib.LAUNCHES - l0      # kernels of ours inside one replay
        # the graph holds raw pointers into the model's per-mask workspace and packed weight images: keep both alive and
        # refuse to replay once the model has dropped either (plan-cache eviction, weight reload / EMA swap)
        self._model, self._masks = model, (node_mask, edge_mask)
        self._token = model.graph_token(node_mask, edge_mask) if hasattr(model, 'graph_token') else None

    def run(self, i):
        if self._token is not None:
            hit, packed = self._model.graph_token(*self._masks)
            if hit is not self._token[0] or packed is not self._token[1]:
                raise RuntimeError('GraphedAncestralStep: the model re-packed its weights or evicted the workspace this '
                                   'graph was captured on; capture a new step')
        self.coef.copy_(self.table[i])
        self.graph.replay()


class HostPipelinedSteps:
    """Reverse steps of a batch whose state lives in pinned HOST buffers, with the host <-> device copies hidden behind the
    compute.  The caller hands over the batch as several PARTS of independent molecules (``split_for_pipeline``: two halves
    with balanced cost), each with its own graph-captured step (``GraphedAncestralStep`` on the part's masks) and its own
    pinned host tensors ``{'x', 'ex', 'cx', 'cex'}``.  ``run(i)`` advances every part by reverse step ``i``: for each part
    host -> device copy of its inputs, graph replay, device -> host copy of its results, on three streams ordered by
    events --

        h2d stream      in(A,i)   in(B,i)              in(A,i+1)   in(B,i+1)
        compute stream       [ step A,i ][ step B,i ][ step A,i+1 ][ step B,i+1 ]
        d2h stream                      out(A,i)     out(B,i)       out(A,i+1)

    -- so the copies of one part run under the kernels of the other and the compute stream never waits for PCIe (a step of
    the whole batch with the copies in line costs compute + 2 x 39 MB / 55 GB/s at QM9 B = 2500).  Nothing is reordered
    inside a part: its step i + 1 reads the host buffers its step i wrote (the device-side dependency in(p,i+1) after
    out(p,i) is an event, not a host sync).  ``synchronize()`` makes every result visible to the host.  Results are
    bit-identical to replaying the same parts one after the other with the copies in line (tests/test_sampler.py)."""

    NAMES = (('x', 'x'), ('ex', 'edge_x'), ('cx', 'cond_x'), ('cex', 'cond_edge_x'))

    def __init__(self, steps, hosts):
        if len(steps) != len(hosts) or not steps:
            raise ValueError('one host state per captured step')
        for h in hosts:
            for k, _ in self.NAMES:
                if not h[k].is_pinned():
                    raise ValueError('host buffers must be pinned (the copies are asynchronous)')
        self.steps, self.hosts = list(steps), list(hosts)
        self.h2d, self.d2h = torch.cuda.Stream(), torch.cuda.Stream()
        mk = lambda: [torch.cuda.Event() for _ in steps]
        self.ev_in, self.ev_done, self.ev_out = mk(), mk(), mk()
        self._started = False
        self.bytes_per_step = sum(h[k].numel() * h[k].element_size() for h in hosts for k, _ in self.NAMES)

    def run(self, i):
        cur = torch.cuda.current_stream()
        if not self._started:
            self.h2d.wait_stream(cur)                    # whatever produced the device state so far is complete
        for p, (gs, host) in enumerate(zip(self.steps, self.hosts)):
            with torch.cuda.stream(self.h2d):
                if self._started:
                    self.h2d.wait_event(self.ev_out[p])  # the part's previous results have left its device buffers
                for k, attr in self.NAMES:
                    getattr(gs, attr).copy_(host[k], non_blocking=True)
                self.ev_in[p].record(self.h2d)
            cur.wait_event(self.ev_in[p])
            gs.run(i)
            self.ev_done[p].record(cur)
            with torch.cuda.stream(self.d2h):
                self.d2h.wait_event(self.ev_done[p])
                for k, attr in self.NAMES:
                    host[k].copy_(getattr(gs, attr), non_blocking=True)
                self.ev_out[p].record(self.d2h)
        self._started = True

    def synchronize(self):
        for e in self.ev_out:
            e.synchronize()


def split_for_pipeline(n_nodes, parts=2):
    """Index tensors of `parts` sub-batches with balanced cost (the dealing of shard_molecules)."""
    return [shard_molecules(n_nodes, parts, r) for r in range(parts)]


class AncestralSampler2D(AncestralSampler):
    """Ancestral sampling without 3-D positions (reference sampling.py:599-661, AncestralSampler_2D): the same
    posterior-mean update with masked Gaussian node noise (sample_gaussian_with_mask, models/utils.py:77-80: no CoM
    projection) and symmetric edge noise; drives ``DGT_concat_2D``."""

    def __init__(self, schedule, time_steps, generator=None, noise_fn=None, s_array=None):
        super().__init__(schedule, time_steps, generator=generator, noise_fn=noise_fn, s_array=s_array, fused=False)

    def _node_noise(self, bs, N, F_, node_mask):
        return torch.randn((bs, N, F_), device=node_mask.device, generator=self.generator) * node_mask


def position_noise(B, N, node_mask, generator=None):
    """sample_center_gravity_zero_gaussian_with_mask, reference models/utils.py:67-74"""
    z = torch.randn((B, N, 3), device=node_mask.device, generator=generator) * node_mask
    return remove_mean_with_mask(z, node_mask)


class DPMSolverSinglestep:
    """DPM-Solver++ 'singlestep_fixed' of order 1 or 2 for joint 2D & 3D generation (reference mix_dpm_solver.py):
    features and bonds follow the exponential-integrator update, positions the ancestral update (:44-59), the model
    is self-conditioned on its previous prediction (:285-291).  One model evaluation per order per outer step."""

    def __init__(self, schedule, steps, order=2, generator=None, noise_fn=None):
        assert order in (1, 2), 'orders 1 and 2 (BASELINE config 5 uses dpm_solver_order = 2)'
        self.ns, self.steps, self.order = schedule, steps, order
        self.generator, self.noise_fn = generator, noise_fn
        self.cond_x = self.cond_edge_x = None
        self.n_noise = 0
        self.n_evals = 0

    # -- pieces ---------------------------------------------------------------------------------------
    def _model(self, model, x, node_mask, edge_mask, edge_x, context, t):
        bs = x.shape[0]
        vec_t = torch.ones(bs, device=x.device) * t
        nl = torch.ones(bs, device=x.device) * self.ns.get_noise_level(t)
        pred, edge_pred = model(vec_t, x, node_mask, edge_mask, edge_x=edge_x, noise_level=nl, cond_x=self.cond_x,
                                cond_edge_x=self.cond_edge_x, context=context)
        self.cond_x, self.cond_edge_x = pred, edge_pred
        self.n_evals += 1
        return pred, edge_pred

    def _position_update(self, pos, pos_pred, node_mask, t_start, t_end, last_step=False):
        """mix_dpm_solver.py:44-59"""
        alpha_t, sigma_t = self.ns.marginal_prob(t_start)
        alpha_s, sigma_s = self.ns.marginal_prob(t_end)
        alpha_ts = alpha_t / alpha_s
        sigma2_ts = sigma_t ** 2 - alpha_ts ** 2 * sigma_s ** 2
        sigma = torch.sqrt(sigma2_ts) * sigma_s / sigma_t
        out = (alpha_ts * sigma_s ** 2 / sigma_t ** 2) * pos + (alpha_s * sigma2_ts / sigma_t ** 2) * pos_pred
        if not last_step:
            if self.noise_fn is not None:
                z = self.noise_fn(self.n_noise)
            else:
                z = position_noise(pos.shape[0], pos.shape[1], node_mask, self.generator)
            self.n_noise += 1
            out = out + sigma * z
        return out

    def _first_update(self, model, x, node_mask, edge_mask, edge_x, context, t_start, t_end, last_step):
        """mix_dpm_solver.py:61-91"""
        ns = self.ns
        h = ns.marginal_lambda(t_end) - ns.marginal_lambda(t_start)
        sigma_start, sigma_end = ns.marginal_std(t_start), ns.marginal_std(t_end)
        alpha_end = torch.exp(ns.marginal_log_mean_coeff(t_end))
        phi_1 = torch.expm1(-h)
        pred, edge_pred = self._model(model, x, node_mask, edge_mask, edge_x, context, t_start)
        atom_end = sigma_end / sigma_start * x[..., 3:] - alpha_end * phi_1 * pred[..., 3:]
        edge_end = sigma_end / sigma_start * edge_x - alpha_end * phi_1 * edge_pred
        pos_end = self._position_update(x[..., :3], pred[..., :3], node_mask, t_start, t_end, last_step)
        return torch.cat([pos_end, atom_end], dim=-1), edge_end

    def _second_update(self, model, x, node_mask, edge_mask, edge_x, context, t_start, t_end, last_step, r1):
        """mix_dpm_solver.py:94-150"""
        ns = self.ns
        lambda_start, lambda_end = ns.marginal_lambda(t_start), ns.marginal_lambda(t_end)
        h = lambda_end - lambda_start
        s1 = ns.inverse_lambda(lambda_start + r1 * h)
        sigma_start, sigma_s1, sigma_end = ns.marginal_std(t_start), ns.marginal_std(s1), ns.marginal_std(t_end)
        alpha_s1, alpha_end = torch.exp(ns.marginal_log_mean_coeff(s1)), torch.exp(ns.marginal_log_mean_coeff(t_end))
        phi_11, phi_1 = torch.expm1(-r1 * h), torch.expm1(-h)
        pos_start, atom_start = x[..., :3], x[..., 3:]
        pred, edge_pred = self._model(model, x, node_mask, edge_mask, edge_x, context, t_start)
        atom_s1 = (sigma_s1 / sigma_start) * atom_start - (alpha_s1 * phi_11) * pred[..., 3:]
        edge_s1 = (sigma_s1 / sigma_start) * edge_x - (alpha_s1 * phi_11) * edge_pred
        pos_s1 = self._position_update(pos_start, pred[..., :3], node_mask, t_start, s1)
        pred1, edge_pred1 = self._model(model, torch.cat([pos_s1, atom_s1], dim=-1), node_mask, edge_mask, edge_s1, context, s1)
        atom_end = ((sigma_end / sigma_start) * atom_start - (alpha_end * phi_1) * pred[..., 3:]
                    - (0.5 / r1) * (alpha_end * phi_1) * (pred1[..., 3:] - pred[..., 3:]))
        edge_end = ((sigma_end / sigma_start) * edge_x - (alpha_end * phi_1) * edge_pred
                    - (0.5 / r1) * (alpha_end * phi_1) * (edge_pred1 - edge_pred))
        pos_end = self._position_update(pos_s1, pred1[..., :3], node_mask, s1, t_end, last_step)
        return torch.cat([pos_end, atom_end], dim=-1), edge_end

    # -- fused update (jodo_dpm_update) ------------------------------------------------------------------
    def step_coefficients(self, step, grid):
        """Device tensors (cA[6], cB[6], nl[2]) of outer step `step` for the order-2 singlestep: {a, b, c, cx, cp, sigma} of
        the intermediate update t_start -> s1 and of the final update -> t_end, and the noise levels of the two model
        evaluations.  Every scalar is produced by the torch expressions of `_second_update` / `_position_update` on the
        grid's device, so the fused path sees the values the torch path computes."""
        ns = self.ns
        t_start, t_end = grid[step], grid[step + 1]
        last = step == len(grid) - 2
        inner = torch.linspace(t_start.item(), t_end.item(), self.order + 1).to(grid.device)
        lam = ns.marginal_lambda(inner)
        r1 = (lam[1] - lam[0]) / (lam[-1] - lam[0])
        lambda_start, lambda_end = ns.marginal_lambda(t_start), ns.marginal_lambda(t_end)
        h = lambda_end - lambda_start
        s1 = ns.inverse_lambda(lambda_start + r1 * h)
        sigma_start, sigma_s1, sigma_end = ns.marginal_std(t_start), ns.marginal_std(s1), ns.marginal_std(t_end)
        alpha_s1, alpha_end = torch.exp(ns.marginal_log_mean_coeff(s1)), torch.exp(ns.marginal_log_mean_coeff(t_end))
        phi_11, phi_1 = torch.expm1(-r1 * h), torch.expm1(-h)

        def pos_coef(ta, tb, no_noise):
            alpha_t, sigma_t = ns.marginal_prob(ta)
            alpha_s, sigma_s = ns.marginal_prob(tb)
            alpha_ts = alpha_t / alpha_s
            sigma2_ts = sigma_t ** 2 - alpha_ts ** 2 * sigma_s ** 2
            sigma = torch.sqrt(sigma2_ts) * sigma_s / sigma_t
            return (alpha_ts * sigma_s ** 2 / sigma_t ** 2, alpha_s * sigma2_ts / sigma_t ** 2,
                    torch.zeros_like(sigma) if no_noise else sigma)

        z = torch.zeros((), device=grid.device)
        f = lambda v: torch.as_tensor(v, device=grid.device, dtype=torch.float32).reshape(())
        cA = torch.stack([f(sigma_s1 / sigma_start), f(alpha_s1 * phi_11), z, *map(f, pos_coef(t_start, s1, False))])
        cB = torch.stack([f(sigma_end / sigma_start), f(alpha_end * phi_1), f((0.5 / r1) * (alpha_end * phi_1)),
                          *map(f, pos_coef(s1, t_end, last))])
        nl = torch.stack([f(ns.get_noise_level(t_start)), f(ns.get_noise_level(s1))])
        return cA, cB, nl

    def _fused_outer(self, model, x, node_mask, edge_mask, edge_x, context, cA, cB, nl, last=False, draw_last=False):
        """One order-2 outer step with both updates fused into jodo_dpm_update (two launches each instead of ~40 torch
        kernels); the raw normal draws are the torch.randn calls of `position_noise`, in the same order."""
        import ctypes
        from . import _lib
        bs, N, F_ = x.shape
        ch = edge_x.shape[-1]
        dev = x.device
        c32 = lambda t: t.contiguous().float()
        x, edge_x, nm = c32(x), c32(edge_x), c32(node_mask)
        vec_t = torch.zeros(bs, device=dev)                                # ignored by the model (reference mol_gnn.py:534)

        def evaluate(xx, ee, k):
            pred, edge_pred = model(vec_t, xx, node_mask, edge_mask, edge_x=ee, noise_level=nl[k:k + 1].expand(bs),
                                    cond_x=self.cond_x, cond_edge_x=self.cond_edge_x, context=context)
            self.cond_x, self.cond_edge_x = pred, edge_pred
            self.n_evals += 1
            return c32(pred), c32(edge_pred)

        def update(pos_in, p0, p1, e0, e1, raw, coef):
            xo, eo = torch.empty_like(x), torch.empty_like(edge_x)
            _lib.call('jodo_dpm_update', _lib.ptr(x), _lib.ptr(pos_in), ctypes.c_int(F_), _lib.ptr(p0), _lib.ptr(p1), _lib.ptr(raw),
                      _lib.ptr(nm), _lib.ptr(edge_x), _lib.ptr(e0), _lib.ptr(e1), ctypes.c_int(bs), ctypes.c_int(N), ctypes.c_int(F_),
                      ctypes.c_int(ch), _lib.ptr(coef), _lib.ptr(xo), _lib.ptr(eo), _lib.stream_ptr())
            return xo, eo

        pred, edge_pred = evaluate(x, edge_x, 0)
        raw1 = torch.randn((bs, N, 3), device=dev, generator=self.generator)
        self.n_noise += 1
        x_s1, e_s1 = update(x, pred, None, edge_pred, None, raw1, cA)
        pred1, edge_pred1 = evaluate(x_s1, e_s1, 1)
        if last and not draw_last:
            raw2 = raw1                                                   # sigma = 0: the reference draws nothing on the last step
        else:
            raw2 = torch.randn((bs, N, 3), device=dev, generator=self.generator)
            self.n_noise += 1
        return update(x_s1, pred, pred1, edge_pred, edge_pred1, raw2, cB)

    # -- driver ---------------------------------------------------------------------------------------
    def outer_grid(self, device):
        K = self.steps // self.order
        return torch.linspace(self.ns.T, 1. / self.ns.total_N, K + 1).to(device)

    def outer_step(self, model, step, grid, x, node_mask, edge_mask, edge_x, context=None, fused=None):
        """One outer step (= `order` model evaluations), mix_dpm_solver.py:322-334.  fused (default: CUDA tensors, order 2,
        default generator path): both updates through jodo_dpm_update."""
        t_start, t_end = grid[step], grid[step + 1]
        last = step == len(grid) - 2
        if fused is None:
            fused = x.is_cuda and self.order == 2 and self.noise_fn is None
        if fused:
            cA, cB, nl = self.step_coefficients(step, grid)
            return self._fused_outer(model, x, node_mask, edge_mask, edge_x, context, cA, cB, nl, last)
        if self.order == 1:
            return self._first_update(model, x, node_mask, edge_mask, edge_x, context, t_start, t_end, last)
        inner = torch.linspace(t_start.item(), t_end.item(), self.order + 1).to(x.device)
        lam = self.ns.marginal_lambda(inner)
        r1 = (lam[1] - lam[0]) / (lam[-1] - lam[0])
        return self._second_update(model, x, node_mask, edge_mask, edge_x, context, t_start, t_end, last, r1)

    @torch.no_grad()
    def sampling(self, model, x, node_mask, edge_mask, edge_x, context=None, graph=False, fused=None):
        """The whole chain (mix_dpm_solver.py:304-376).  graph=True (order 2, CUDA): outer step 0 runs eagerly (its first
        evaluation has no self-conditioning), then one outer step is captured into a CUDA graph and replayed with per-step
        coefficients -- same kernels, same torch.randn stream as the eager fused chain, except that the replayed last step
        also draws (and multiplies by sigma = 0) the noise the eager chain skips."""
        self.cond_x = self.cond_edge_x = None
        self.n_noise = self.n_evals = 0
        grid = self.outer_grid(x.device)
        n = len(grid) - 1
        if graph and n > 1:
            x, edge_x = self.outer_step(model, 0, grid, x, node_mask, edge_mask, edge_x, context, fused=True)
            gs = GraphedDPMStep(self, model, x, edge_x, node_mask, edge_mask, context, grid)
            for step in range(1, n):
                gs.run(step)
            return gs.x.clone(), gs.edge_x.clone()
        for step in range(n):
            x, edge_x = self.outer_step(model, step, grid, x, node_mask, edge_mask, edge_x, context, fused=fused)
        return x, edge_x


class GraphedDPMStep:
    """One order-2 outer step of a DPMSolverSinglestep (two denoiser calls + two noise draws + two fused updates + hand-off
    copies) captured into a CUDA graph.  ``run(step)`` writes that step's 14 coefficients into device memory and replays.
    State: x, edge_x (in and out), cond_x / cond_edge_x (the solver's self-conditioning).  The solver must have run one
    outer step with this model and these masks before (self-conditioned path, workspaces, plan)."""

    def __init__(self, solver, model, x, edge_x, node_mask, edge_mask, context, grid):
        if not x.is_cuda or solver.order != 2 or solver.noise_fn is not None or solver.generator is not None:
            raise ValueError('the graph-captured outer step needs CUDA tensors, order 2 and the default CUDA generator')
        if solver.cond_x is None:
            raise ValueError('capture after the first outer step: the first evaluation of a chain is not self-conditioned')
        from . import _lib
        c32 = lambda t: t.contiguous().float()
        n = len(grid) - 1
        rows = [torch.cat(solver.step_coefficients(k, grid)) for k in range(n)]
        self.table = torch.stack(rows)                                     # [steps, 14] = cA | cB | nl
        self.coef = torch.zeros(14, device=x.device, dtype=torch.float32)
        self.x, self.edge_x = c32(x).clone(), c32(edge_x).clone()
        self.cond_x, self.cond_edge_x = c32(solver.cond_x).clone(), c32(solver.cond_edge_x).clone()
        ctx = None if context is None else c32(context).clone()
        self._static = (ctx, node_mask, edge_mask)                         # read by the captured kernels on every replay
        self.solver = solver

        def one_step():
            solver.cond_x, solver.cond_edge_x = self.cond_x, self.cond_edge_x
            xo, eo = solver._fused_outer(model, self.x, node_mask, edge_mask, self.edge_x, ctx, self.coef[0:6], self.coef[6:12],
                                         self.coef[12:14], last=True, draw_last=True)
            self.x.copy_(xo); self.edge_x.copy_(eo)
            self.cond_x.copy_(solver.cond_x); self.cond_edge_x.copy_(solver.cond_edge_x)

        self.graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        l0, e0, z0 = _lib.LAUNCHES, solver.n_evals, solver.n_noise
        with torch.cuda.stream(side):
            with torch.cuda.graph(self.graph, stream=side):
                one_step()
        torch.cuda.current_stream().wait_stream(side)
        self.launches_per_step = _lib.LAUNCHES - l0
        solver.n_evals, solver.n_noise = e0, z0                            # the capture ran nothing
        solver.cond_x, solver.cond_edge_x = self.cond_x, self.cond_edge_x
        self._model, self._masks = model, (node_mask, edge_mask)
        self._token = model.graph_token(node_mask, edge_mask) if hasattr(model, 'graph_token') else None

    def run(self, step):
        if self._token is not None:
            hit, packed = self._model.graph_token(*self._masks)
            if hit is not self._token[0] or packed is not self._token[1]:
                raise RuntimeError('GraphedDPMStep: the model re-packed its weights or evicted the workspace this graph was '
                                   'captured on; capture a new step')
        self.coef.copy_(self.table[step])
        self.graph.replay()
        self.solver.n_evals += 2
        self.solver.n_noise += 2


# ---- multi-GPU: shard independent molecules, gather final samples --------------------------------
def shard_molecules(n_nodes, world_size, rank):
    """Indices of the molecules rank `rank` owns: sort by size (descending) and deal in a snake order
    so that the per-rank sum of n(n-1) (the cost driver) is balanced."""
    n = torch.as_tensor(n_nodes, dtype=torch.int64)
    order = torch.argsort(n, descending=True, stable=True)
    pos = torch.arange(len(n))
    rnd, k = pos // world_size, pos % world_size
    owner = torch.where(rnd % 2 == 0, k, world_size - 1 - k)
    return torch.sort(order[owner == rank]).values


def gather_samples(x, edge_x, idx, total, N, group=None):
    """all_gather the per-rank final samples into global-order tensors [total, N, F] / [total, N, N, ch].
    Every rank pads its shard to the global N and to the largest shard size; one collective per tensor."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = x.device
    cnt = torch.tensor([x.shape[0]], device=dev, dtype=torch.int64)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt, group=group)
    m = int(max(int(c) for c in cnts))
    F_, ch = x.shape[2], edge_x.shape[-1]
    px = torch.zeros(m, N, F_, device=dev, dtype=x.dtype)
    pe = torch.zeros(m, N, N, ch, device=dev, dtype=edge_x.dtype)
    pi = torch.full((m,), -1, device=dev, dtype=torch.int64)
    b, n = x.shape[0], x.shape[1]
    px[:b, :n] = x
    pe[:b, :n, :n] = edge_x
    pi[:b] = idx.to(dev)
    gx = [torch.empty_like(px) for _ in range(world)]
    ge = [torch.empty_like(pe) for _ in range(world)]
    gi = [torch.empty_like(pi) for _ in range(world)]
    dist.all_gather(gx, px, group=group)
    dist.all_gather(ge, pe, group=group)
    dist.all_gather(gi, pi, group=group)
    out_x = torch.zeros(total, N, F_, device=dev, dtype=x.dtype)
    out_e = torch.zeros(total, N, N, ch, device=dev, dtype=edge_x.dtype)
    for r in range(world):
        k = gi[r] >= 0
        out_x[gi[r][k]] = gx[r][k]
        out_e[gi[r][k]] = ge[r][k]
    return out_x, out_e
