"""ORACLE (test infrastructure, not product code): CPU restatement of the reference's ancestral
reverse-SDE step around the denoiser, with the noise passed in explicitly so that a recorded chain
can be replayed.

Follows reference sampling.py:530-596 (AncestralSampler.sampling with model_pred_data=True,
pred_edge=True, self_cond=True, cond_process_fn = 'ori' identity, reference utils.py:134-136) and the
cosine VP schedule of reference diffusion/noise_schedule.py:43-52,76-92.

Parity status: PINNED on tests/golden/qm9_ancestral_chain.pt (5 steps of the unmodified reference
sampler on the real 1000-step grid with every noise draw recorded; generator oracle/make_golden.py,
test tests/test_sampler.py).
"""
from __future__ import annotations

import math

import torch

COSINE_S = 0.008
COSINE_LOG_ALPHA_0 = math.log(math.cos(COSINE_S / (1. + COSINE_S) * math.pi / 2.))
T_END = 0.9946          # NoiseScheduleVP('cosine').T (diffusion/noise_schedule.py:50-52)


def marginal_prob(t):
    """alpha_t, sigma_t of the cosine schedule (diffusion/noise_schedule.py:76-79,89-92)."""
    log_alpha = torch.log(torch.cos((t + COSINE_S) / (1. + COSINE_S) * math.pi / 2.)) - COSINE_LOG_ALPHA_0
    return torch.exp(log_alpha), torch.sqrt(1. - torch.exp(2. * log_alpha))


def ancestral_step(model, t, s, x, edge_x, node_mask, edge_mask, cond_x, cond_edge_x, z_node, z_edge, context=None):
    """One iteration of the loop at sampling.py:535-589.  z_node / z_edge are the draws of
    sample_combined_position_feature_noise / sample_symmetric_edge_feature_noise for this step.
    Returns (x, edge_x, x_mean, edge_x_mean, pred, edge_pred)."""
    bs = x.shape[0]
    alpha_t, sigma_t = marginal_prob(t)
    alpha_s, sigma_s = marginal_prob(s)
    alpha_ts = alpha_t / alpha_s
    sigma2_ts = sigma_t ** 2 - alpha_ts ** 2 * sigma_s ** 2
    sigma = torch.sqrt(sigma2_ts) * sigma_s / sigma_t
    vec_t = torch.ones(bs) * t
    noise_level = torch.ones(bs) * torch.log(alpha_t ** 2 / sigma_t ** 2)
    pred, edge_pred = model(vec_t, x, node_mask, edge_mask, edge_x=edge_x, noise_level=noise_level, cond_x=cond_x,
                            cond_edge_x=cond_edge_x, context=context)
    c_x = alpha_ts * sigma_s ** 2 / sigma_t ** 2
    c_p = alpha_s * sigma2_ts / sigma_t ** 2
    x_mean = c_x * x + c_p * pred
    x_new = x_mean + sigma * z_node
    e_mean = c_x * edge_x + c_p * edge_pred
    e_new = e_mean + sigma * z_edge
    return x_new, e_new, x_mean, e_mean, pred, edge_pred


def replay_chain(model, t_array, s_array, z_T, edge_z_T, node_mask, edge_mask, noise_node, noise_edge, context=None):
    x, ex = z_T, edge_z_T
    cx = cex = None
    xm = em = None
    for i in range(len(t_array)):
        x, ex, xm, em, cx, cex = ancestral_step(model, t_array[i], s_array[i], x, ex, node_mask, edge_mask, cx, cex,
                                                noise_node[i], noise_edge[i], context)
    return xm, em


# ---- DPM-Solver++ 'singlestep_fixed', order 2 (reference mix_dpm_solver.py:44-59, 93-150, 285-335) -------------------
# Pinned on tests/golden/qm9_cond_dpm_chain.pt (3 outer steps of the unmodified reference solver on the conditional
# QM9 model with every position-noise draw recorded).

def _log_alpha(t):
    return torch.log(torch.cos((t + COSINE_S) / (1. + COSINE_S) * math.pi / 2.)) - COSINE_LOG_ALPHA_0


def _std(t):
    return torch.sqrt(1. - torch.exp(2. * _log_alpha(t)))


def _lambda(t):
    la = _log_alpha(t)
    return la - 0.5 * torch.log(1. - torch.exp(2. * la))


def _inverse_lambda(lamb):
    log_alpha = -0.5 * torch.logaddexp(-2. * lamb, torch.zeros((1,)).to(lamb))
    return torch.arccos(torch.exp(log_alpha + COSINE_LOG_ALPHA_0)) * 2. * (1. + COSINE_S) / math.pi - COSINE_S


def _noise_level(t):
    return torch.log(torch.exp(_log_alpha(t)) ** 2 / _std(t) ** 2)


def dpm_singlestep2_chain(model, x, edge_x, node_mask, edge_mask, context, steps, noise_pos, total_N=1000):
    """model(vec_t, x, node_mask, edge_mask, edge_x=, noise_level=, cond_x=, cond_edge_x=, context=) -> (pred, edge_pred).
    noise_pos: the draws of sample_center_gravity_zero_gaussian_with_mask in call order."""
    K = steps // 2
    grid = torch.linspace(T_END, 1. / total_N, K + 1)
    cond = [None, None]
    draws = iter(noise_pos)
    bs = x.shape[0]

    def call(xx, ee, t):
        p, e = model(torch.ones(bs) * t, xx, node_mask, edge_mask, edge_x=ee, noise_level=torch.ones(bs) * _noise_level(t),
                     cond_x=cond[0], cond_edge_x=cond[1], context=context)
        cond[0], cond[1] = p, e
        return p, e

    def pos_update(pos, pos_pred, t_start, t_end, last=False):
        a_t, s_t = marginal_prob(t_start)
        a_s, s_s = marginal_prob(t_end)
        a_ts = a_t / a_s
        s2_ts = s_t ** 2 - a_ts ** 2 * s_s ** 2
        out = (a_ts * s_s ** 2 / s_t ** 2) * pos + (a_s * s2_ts / s_t ** 2) * pos_pred
        if not last:
            out = out + torch.sqrt(s2_ts) * s_s / s_t * next(draws)
        return out

    for step in range(K):
        t0, t1 = grid[step], grid[step + 1]
        last = step == K - 1
        inner = torch.linspace(t0.item(), t1.item(), 3)
        lam = _lambda(inner)
        r1 = (lam[1] - lam[0]) / (lam[-1] - lam[0])
        l0, l1 = _lambda(t0), _lambda(t1)
        h = l1 - l0
        s1 = _inverse_lambda(l0 + r1 * h)
        sg0, sg_s1, sg1 = _std(t0), _std(s1), _std(t1)
        al_s1, al1 = torch.exp(_log_alpha(s1)), torch.exp(_log_alpha(t1))
        phi_11, phi_1 = torch.expm1(-r1 * h), torch.expm1(-h)
        p0, e0 = call(x, edge_x, t0)
        atom_s1 = (sg_s1 / sg0) * x[..., 3:] - (al_s1 * phi_11) * p0[..., 3:]
        edge_s1 = (sg_s1 / sg0) * edge_x - (al_s1 * phi_11) * e0
        pos_s1 = pos_update(x[..., :3], p0[..., :3], t0, s1)
        p1, e1 = call(torch.cat([pos_s1, atom_s1], -1), edge_s1, s1)
        atom1 = (sg1 / sg0) * x[..., 3:] - (al1 * phi_1) * p0[..., 3:] - (0.5 / r1) * (al1 * phi_1) * (p1[..., 3:] - p0[..., 3:])
        edge1 = (sg1 / sg0) * edge_x - (al1 * phi_1) * e0 - (0.5 / r1) * (al1 * phi_1) * (e1 - e0)
        pos1 = pos_update(pos_s1, p1[..., :3], s1, t1, last)
        x, edge_x = torch.cat([pos1, atom1], -1), edge1
    return x, edge_x
