"""Synthetic DGT inputs of the shapes the reference sampler feeds the denoiser.

Mirrors what ``sampling_fn`` builds before the hot loop (reference sampling.py:179-209):
molecule sizes from the dataset histogram (models/node_distribution.py:27), prefix node masks,
edge mask = outer product minus diagonal (sampling.py:194-201), CoM-free Gaussian positions +
Gaussian features (models/utils.py:67-90) and symmetric edge noise (:93-99).
Everything is generated on the CPU with an explicit generator, so the same seed gives the same
batch here and on the GPU box.
"""
from __future__ import annotations

import torch

from .datasets_info import HISTOGRAMS


def sample_n_nodes(info_name, batch, gen, max_n=None):
    hist = HISTOGRAMS[info_name]
    keys = sorted(hist)
    if max_n is not None:
        keys = [k for k in keys if k <= max_n]
    prob = torch.tensor([hist[k] for k in keys], dtype=torch.float64)
    idx = torch.multinomial(prob / prob.sum(), batch, replacement=True, generator=gen)
    return torch.tensor(keys, dtype=torch.int64)[idx]


def make_masks(n_nodes, N=None):
    """node_mask [B,N,1], edge_mask [B*N*N,1] exactly as sampling.py:194-201."""
    B = len(n_nodes)
    N = int(max(n_nodes)) if N is None else N
    node_mask = (torch.arange(N)[None, :] < torch.as_tensor(n_nodes)[:, None]).float()
    edge_mask = node_mask[:, None, :] * node_mask[:, :, None]
    edge_mask = edge_mask * (~torch.eye(N, dtype=torch.bool))[None]
    return node_mask.unsqueeze(2), edge_mask.reshape(B * N * N, 1)


def remove_mean_with_mask(x, node_mask):
    """models/utils.py:38-45."""
    n = node_mask.sum(1, keepdim=True)
    return x - (x.sum(1, keepdim=True) / n) * node_mask


def node_noise(B, N, feat, node_mask, gen, device=None, dtype=torch.float32):
    """sample_combined_position_feature_noise (models/utils.py:83-90)."""
    dev = device if device is not None else node_mask.device
    zx = torch.randn((B, N, 3), generator=gen, device=dev, dtype=dtype) * node_mask
    zx = remove_mean_with_mask(zx, node_mask)
    zh = torch.randn((B, N, feat), generator=gen, device=dev, dtype=dtype) * node_mask
    return torch.cat([zx, zh], dim=2)


def edge_noise(B, N, ch, edge_mask, gen, device=None, dtype=torch.float32):
    """sample_symmetric_edge_feature_noise (models/utils.py:93-99)."""
    dev = device if device is not None else edge_mask.device
    z = torch.randn((B, ch, N, N), generator=gen, device=dev, dtype=dtype)
    z = torch.tril(z, -1)
    z = z + z.transpose(-1, -2)
    return z.permute(0, 2, 3, 1) * edge_mask.reshape(B, N, N, 1)


def make_batch(config, batch, seed=42, max_n=None, n_nodes=None, self_cond=False, context=False,
               noise_level=None, dtype=torch.float32):
    """One denoiser call's worth of inputs (CPU tensors).

    self_cond=True also returns cond_x / cond_edge_x of the form the sampler hands back
    (a previous prediction: masked, CoM-free positions, symmetric edges)."""
    gen = torch.Generator().manual_seed(seed)
    inn = int(config.data.atom_types) + int(config.model.include_fc_charge)
    ch = int(config.model.edge_ch)
    if n_nodes is None:
        n_nodes = sample_n_nodes(config.data.info_name, batch, gen, max_n)
    n_nodes = torch.as_tensor(n_nodes, dtype=torch.int64)
    B = len(n_nodes)
    N = int(n_nodes.max())
    node_mask, edge_mask = make_masks(n_nodes, N)
    node_mask, edge_mask = node_mask.to(dtype), edge_mask.to(dtype)
    out = dict(n_nodes=n_nodes, node_mask=node_mask, edge_mask=edge_mask)
    two_d = bool(getattr(config, 'only_2D', False))          # 2-D models: xh = atom features only (no coordinates)
    feat_noise = lambda: torch.randn((B, N, inn), generator=gen, dtype=dtype) * node_mask
    out['xh'] = feat_noise() if two_d else node_noise(B, N, inn, node_mask, gen, dtype=dtype)
    out['edge_x'] = edge_noise(B, N, ch, edge_mask, gen, dtype=dtype)
    if noise_level is None:
        nl = torch.empty(B, dtype=dtype).uniform_(-6.0, 6.0, generator=gen)
    else:
        nl = torch.full((B,), float(noise_level), dtype=dtype)
    out['noise_level'] = nl
    out['t'] = torch.rand(B, generator=gen, dtype=dtype)
    if self_cond:
        if two_d:
            cx = feat_noise()
        else:
            cx = node_noise(B, N, inn, node_mask, gen, dtype=dtype)
            cx[..., :3] = cx[..., :3] * 1.5
        out['cond_x'] = cx
        out['cond_edge_x'] = edge_noise(B, N, ch, edge_mask, gen, dtype=dtype) * 0.7
    else:
        out['cond_x'] = None
        out['cond_edge_x'] = None
    cond_ch = int(getattr(config.model, 'cond_ch', 1))
    out['context'] = torch.randn(B, cond_ch, generator=gen, dtype=dtype) if context else None
    return out
