"""Host side of the reverse-SDE loop around the denoiser: the reference's ancestral sampler, its
VP cosine schedule, and the multi-GPU sharding of a batch of independent molecules.

This mirrors the caller of the hot path, so that the bench and the tests can
drive the denoiser exactly the way ``sampling_fn`` does without the reference tree being present:

* ``CosineVP``            - NoiseScheduleVP('cosine'), reference diffusion/noise_schedule.py:43-52,76-92
* ``AncestralSampler``    - reference sampling.py:518-596 (pred_edge=True, self_cond=True, model_pred_data=True)
* ``node_noise/edge_noise`` - reference models/utils.py:67-99 (same torch.randn call order, so on one
  device with one seed the stream is the one the reference would draw)
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

    def __init__(self, schedule, time_steps, generator=None, noise_fn=None, s_array=None):
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
        x_mean = c_x * x + c_pred * pred
        if self.noise_fn is not None:
            zn, ze = self.noise_fn(i, 'node'), self.noise_fn(i, 'edge')
        else:
            zn = node_noise(bs, N, x.shape[2] - 3, node_mask, self.generator)
            ze = None
        x_new = x_mean + sigma * zn
        edge_mean = c_x * edge_x + c_pred * edge_pred
        if ze is None:
            ze = edge_noise(bs, N, edge_x.shape[-1], edge_mask, self.generator)
        edge_new = edge_mean + sigma * ze
        return x_new, edge_new, x_mean, edge_mean, pred, edge_pred

    @torch.no_grad()
    def sampling(self, model, z_T, node_mask, edge_mask, edge_z_T, context=None):
        x, edge_x = z_T, edge_z_T
        cond_x = cond_edge_x = None
        x_mean = edge_mean = None
        for i in range(len(self.t_array)):
            x, edge_x, x_mean, edge_mean, cond_x, cond_edge_x = self.step(
                model, i, x, edge_x, node_mask, edge_mask, cond_x, cond_edge_x, context)
        return x_mean, edge_mean


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
