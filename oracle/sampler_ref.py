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
