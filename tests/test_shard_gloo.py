"""Multi-GPU host logic on CPU (gloo, world_size 2): molecules are independent, each rank owns a cost-balanced shard,
there is no collective inside the denoising loop and one gather of the final samples (the reference's
single-process consumer is sampling.py:211-213)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from jodo_b200 import configs, sampler as S, synth


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_denoiser(t, xh, node_mask, edge_mask, **kw):
    """A deterministic per-molecule function (no cross-molecule coupling), standing in for the CUDA denoiser."""
    B, N = xh.shape[:2]
    pos = S.remove_mean_with_mask(torch.tanh(xh[..., :3]) * node_mask, node_mask)
    pred = torch.cat([pos, torch.sin(xh[..., 3:]) * node_mask], dim=2)
    e = torch.cos(kw['edge_x']) * edge_mask.reshape(B, N, N, 1)
    return pred, 0.5 * (e + e.transpose(1, 2))


def _worker(rank, world, port, n_nodes, steps, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        cfg = configs.NAMED['qm9_uncond']()
        full = synth.make_batch(cfg, len(n_nodes), seed=5, n_nodes=n_nodes)
        idx = S.shard_molecules(n_nodes, world, rank)
        n_loc = n_nodes[idx]
        N_loc = int(n_loc.max())
        nm, em = synth.make_masks(n_loc, N_loc)
        x = full['xh'][idx][:, :N_loc]
        ex = full['edge_x'][idx][:, :N_loc, :N_loc]
        # noise replay keyed by the GLOBAL molecule index, so that the result does not depend on the sharding
        N_glob = int(n_nodes.max())

        def noise(i, kind):
            g = torch.Generator().manual_seed(1000 + i)
            zn = S.node_noise(len(n_nodes), N_glob, x.shape[2] - 3, full['node_mask'], g)
            ze = S.edge_noise(len(n_nodes), N_glob, ex.shape[-1], full['edge_mask'], g)
            return zn[idx][:, :N_loc] if kind == 'node' else ze[idx][:, :N_loc, :N_loc]

        smp = S.AncestralSampler(S.CosineVP(), torch.linspace(0.9946, 1e-3, 1000)[:steps], noise_fn=noise)
        xm, em_ = smp.sampling(_fake_denoiser, x, nm, em, ex)
        gx, ge = S.gather_samples(xm, em_, idx, len(n_nodes), N_glob)
        if rank == 0:
            torch.save(dict(x=gx, e=ge), os.path.join(out_dir, 'gathered.pt'))
    finally:
        dist.destroy_process_group()


def test_shards_are_balanced_and_disjoint():
    g = torch.Generator().manual_seed(0)
    n = synth.sample_n_nodes('qm9_with_h', 2500, g)
    for world in (2, 4, 8):
        shards = [S.shard_molecules(n, world, r) for r in range(world)]
        allidx = torch.cat(shards)
        assert sorted(allidx.tolist()) == list(range(len(n)))
        cost = torch.tensor([float((n[s] * (n[s] - 1)).sum()) for s in shards])
        assert float(cost.max() / cost.mean()) < 1.01            # snake deal over sizes: < 1 % imbalance in n(n-1)
        assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1


@pytest.mark.timeout(300)
def test_two_rank_sampling_matches_single_process(tmp_path):
    n_nodes = torch.tensor([5, 9, 3, 12, 7, 4, 12, 6, 8])
    steps = 3
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_nodes, steps, str(tmp_path)), nprocs=2, join=True)
    got = torch.load(os.path.join(tmp_path, 'gathered.pt'))
    # single-process run over the whole batch with the same replayed noise
    cfg = configs.NAMED['qm9_uncond']()
    full = synth.make_batch(cfg, len(n_nodes), seed=5, n_nodes=n_nodes)
    N = int(n_nodes.max())

    def noise(i, kind):
        g = torch.Generator().manual_seed(1000 + i)
        zn = S.node_noise(len(n_nodes), N, full['xh'].shape[2] - 3, full['node_mask'], g)
        ze = S.edge_noise(len(n_nodes), N, full['edge_x'].shape[-1], full['edge_mask'], g)
        return zn if kind == 'node' else ze

    smp = S.AncestralSampler(S.CosineVP(), torch.linspace(0.9946, 1e-3, 1000)[:steps], noise_fn=noise)
    xm, em = smp.sampling(_fake_denoiser, full['xh'], full['node_mask'], full['edge_mask'], full['edge_x'])
    assert got['x'].shape == xm.shape and got['e'].shape == em.shape
    assert float((got['x'] - xm).abs().max()) < 1e-6            # per-molecule math: padding width must not matter
    assert float((got['e'] - em).abs().max()) < 1e-6
