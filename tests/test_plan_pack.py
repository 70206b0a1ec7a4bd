"""Host-side logic of the hot path that runs without a GPU: the varlen plan (replacement of the reference's
per-forward dense_to_sparse, models/mol_gnn.py:512-514), the operand-image packers and the algorithmic roofline model."""
import numpy as np
import pytest
import torch

from jodo_b200 import configs, roofline
from jodo_b200.pack import (Packed, image_to_matrix, image_to_matrix_h, matrix_to_image, pack_model, pad2, weight_image,
                            weight_image_h)
from jodo_b200.params import dims_from_config, param_spec, synth_state_dict
from jodo_b200.plan import TILE, Plan


def _mask(n_list, N=None):
    n = torch.tensor(n_list)
    N = int(n.max()) if N is None else N
    return (torch.arange(N)[None] < n[:, None]).float().unsqueeze(-1)


@pytest.mark.parametrize('n_list', [[3], [1, 1, 5], [29] * 7, [2, 9, 1, 17, 29, 3, 3, 3, 12], [80, 44, 63], [129]])
def test_plan_invariants(n_list):
    m = _mask(n_list)
    p = Plan(m)
    n = np.array(n_list)
    assert p.B == len(n_list) and p.Nn == int(n.sum()) and p.n_edges == int((n * (n - 1)).sum())
    row_g, row_j = p.row_g.numpy(), p.row_j.numpy()
    meta = p.row_meta.numpy().view(np.uint32)
    valid = row_g >= 0
    assert valid.sum() == p.n_edges and (row_j[~valid] == -1).all()
    mol = p.node_mol.numpy()
    # every ordered pair (g, j), g != j, of the same molecule appears exactly once
    pairs = set(zip(row_g[valid].tolist(), row_j[valid].tolist()))
    assert len(pairs) == p.n_edges
    assert all(mol[g] == mol[j] and g != j for g, j in pairs)
    # groups: contiguous, never split across tiles, partners in ascending atom order, metadata consistent
    R = np.nonzero(valid)[0]
    gs, gl, gi = meta[R] & 255, (meta[R] >> 8) & 255, (meta[R] >> 16) & 255
    tile = R // TILE
    for g in np.unique(row_g[valid]):
        rows = R[row_g[R] == g]
        assert (np.diff(rows) == 1).all() and len(np.unique(rows // TILE)) == 1
        assert (np.diff(row_j[rows]) > 0).all()
        k = row_g[R] == g
        assert (gs[k] == rows[0] % TILE).all() and (gl[k] == len(rows)).all() and len(np.unique(gi[k])) == 1
        assert len(rows) == n[mol[g]] - 1
    ng = p.tile_ngroups.numpy()
    assert len(ng) == p.n_tiles
    for t in range(p.n_tiles):
        k = tile == t
        assert ng[t] == (len(np.unique(row_g[R][k])) if k.any() else 0)
    # dense <-> rows round trip in both orientations
    dense = torch.randn(p.B, p.N, p.N, 3) * (m * m.transpose(1, 2)).unsqueeze(-1) * (1 - torch.eye(p.N))[None, :, :, None]
    for gf in (True, False):
        assert torch.equal(p.rows_to_dense(p.dense_to_rows(dense, gf), gf), dense)


@pytest.mark.parametrize('n_list', [[130], [181, 40, 2, 1], [256, 3], [5, 9, 1]])
def test_loose_plan_lifts_the_tile_limit(n_list):
    """Molecules with more than 129 atoms (GEOM-Drugs goes to 181): groups back to back, straddling tiles -- the layout
    of the wide path.  Same enumeration, same group tables."""
    m = _mask(n_list)
    p = Plan(m, loose=True if max(n_list) <= 129 else None)
    assert p.loose
    n = np.array(n_list)
    assert p.n_edges == int((n * (n - 1)).sum()) and p.n_tiles == max(1, -(-p.n_edges // TILE))
    row_g, row_j = p.row_g.numpy(), p.row_j.numpy()
    valid = row_g >= 0
    assert valid[:p.n_edges].all() and not valid[p.n_edges:].any()          # no holes
    r0, gl = p.grp_row0.numpy(), p.grp_len.numpy()
    mol = p.node_mol.numpy()
    assert p.max_group == int(n.max()) - 1
    for g in range(p.Nn):
        assert gl[g] == n[mol[g]] - 1
        rows = np.arange(r0[g], r0[g] + gl[g])
        assert (row_g[rows] == g).all() and (np.diff(row_j[rows]) > 0).all()
        assert (p.row_mol.numpy()[rows] == mol[g]).all()
    dense = torch.randn(p.B, p.N, p.N, 2) * (m * m.transpose(1, 2)).unsqueeze(-1) * (1 - torch.eye(p.N))[None, :, :, None]
    assert torch.equal(p.rows_to_dense(p.dense_to_rows(dense)), dense)


def test_tight_plan_group_tables():
    p = Plan(_mask([2, 9, 1, 17, 29, 3]))
    assert not p.loose
    row_g, r0, gl = p.row_g.numpy(), p.grp_row0.numpy(), p.grp_len.numpy()
    for g in range(p.Nn):
        rows = np.arange(r0[g], r0[g] + gl[g])
        assert (row_g[rows] == g).all()


def test_plan_rejects_oversized_and_empty_molecules():
    with pytest.raises(ValueError):
        Plan(_mask([130]), loose=False)
    with pytest.raises(ValueError):
        Plan(_mask([257]))
    with pytest.raises(ValueError):
        Plan(_mask([0, 4], N=4))


def test_plan_matches_reference_edge_enumeration():
    """The reference enumerates edges in (b, row, col) order over the dense adjacency (dense_to_sparse); grouping
    the same list by its first index must give the plan's groups."""
    m = _mask([4, 2, 6])
    p = Plan(m)
    nm = m[..., 0]
    adj = nm[:, :, None] * nm[:, None, :] * (1 - torch.eye(p.N))[None]
    b, r, c = adj.nonzero(as_tuple=True)
    dense_id = (b * p.N + r).numpy(), (b * p.N + c).numpy()
    nd = p.node_dense.numpy()
    rg, rj = p.row_g.numpy(), p.row_j.numpy()
    v = rg >= 0
    ours = sorted(zip(nd[rg[v]].tolist(), nd[rj[v]].tolist()))
    assert ours == sorted(zip(dense_id[0].tolist(), dense_id[1].tolist()))


@pytest.mark.parametrize('n,k,nt', [(64, 64, 64), (256, 128, 128), (768, 256, 256), (16, 64, 16)])
def test_weight_image_h_round_trip(n, k, nt):
    w = torch.randn(n, k)
    img = weight_image_h(w, nt).view(torch.float16)
    tiles = img.reshape(n // nt, -1)
    back = torch.cat([image_to_matrix_h(tiles[t], nt, k) for t in range(n // nt)])
    assert torch.equal(back, w.half().float())


def test_fp32_image_round_trip_and_tf32_rounding():
    m = torch.randn(128, 96)
    assert torch.equal(image_to_matrix(matrix_to_image(m), 128, 96), m)
    w = torch.tensor([[1.0 + 2 ** -11, 1.0 + 2 ** -12, -3.0, 65519.0] * 8] * 8)
    r = image_to_matrix(weight_image(w, 8), 8, 32)
    assert float(r[0, 0]) == 1.0 + 2 ** -10 and float(r[0, 1]) == 1.0 and float(r[0, 2]) == -3.0


def test_split_heads_layout():
    """lin_query / lin_key / lin_edge0 rows in the packed q | k | v image: heads 0..6 at rows [0, 126), heads 7..13 at rows
    [128, 254) of every 256-row block, zeros elsewhere (csrc/attn.cu reads one 128-column half per warp group)."""
    cfg = configs.NAMED['qm9_uncond']()
    d = dims_from_config(cfg)
    sd = synth_state_dict(param_spec(cfg), seed=1)
    pk = pack_model(sd, d, 'cpu')
    img = pk['b0.qkv.img'].view(torch.float16)
    tiles = [image_to_matrix_h(img[t * 256 * 256:(t + 1) * 256 * 256], 256, 256) for t in range(3)]
    wq = sd['e_block_0.attn_mpnn.lin_query.weight'].half().float()
    wk = sd['e_block_0.attn_mpnn.lin_key.weight'].half().float()
    for s_, w in ((tiles[0], wq), (tiles[1], wk)):
        assert torch.equal(s_[:126], w[:126]) and torch.equal(s_[128:254], w[126:])
        assert float(s_[126:128].abs().sum()) == 0 and float(s_[254:].abs().sum()) == 0
    assert torch.equal(tiles[2], sd['e_block_0.attn_mpnn.lin_value.weight'].half().float())


def test_recorded_pieces_equal_the_reference_layout_functions():
    """The three destination formats of jodo_pack_weights (as emulated on the CPU) against the stand-alone layout
    functions: a zero-padded matrix assembled from offset pieces, scaled, in N tiles."""
    g = torch.Generator().manual_seed(0)
    a, b = torch.randn(40, 70, generator=g), torch.randn(24, 50, generator=g) * 300.0
    pk = Packed('cpu')
    pk.image_h('h', 128, 192, 64, [(a, 3, 5), (b, 64, 128, 0.5)])
    pk.image_tf32('t', 64, 96, 64, [(a[:, :60], 8, 32)])
    pk.mat('m', 4, 80, [(a[0], 0, 0), (a[1, :10], 2, 7, 2.0, 1.0)], host=True)
    pk.finish()
    ref = torch.zeros(128, 192)
    ref[3:43, 5:75] = a
    ref[64:88, 128:178] = 0.5 * b
    assert torch.equal(pk['h'].view(torch.int32), weight_image_h(ref, 64).view(torch.int32))
    ref = torch.zeros(64, 96)
    ref[8:48, 32:92] = a[:, :60]
    assert torch.equal(pk['t'], weight_image(ref, 64))
    ref = torch.zeros(4, 80)
    ref[0, :70] = a[0]
    ref[2, 7:17] = 2.0 * a[1, :10] + 1.0
    assert torch.equal(pk['m'].reshape(4, 80), ref) and list(pk.host['m']) == ref.reshape(-1).tolist()


@pytest.mark.parametrize('name', ['qm9_uncond', 'qm9_cond', 'geom_l8', 'geom_l10'])
def test_pack_model_cpu(name):
    """pack_model runs on CPU tensors (no kernels involved): every reference parameter must be consumed into an
    image / table of the expected size, and the per-column constant tables passed by value have the ABI's lengths."""
    cfg = configs.NAMED[name]()
    d = dims_from_config(cfg)
    sd = synth_state_dict(param_spec(cfg), seed=1)
    pk = pack_model(sd, d, 'cpu')
    assert pk.meta['ld_tab'] % 256 == 0 and pk.meta['keh'] == 192
    for l in range(d.L):
        p = f'b{l}.'
        assert len(pk.host[p + 'b0h']) == 256 and len(pk.host[p + 'gbf4']) == 256
        assert pk[p + 'w2.img'].numel() * 4 == 16 * 256 * 2
        assert len(pk.host[p + 'ff3.b']) == 256 and len(pk.host[p + 'emb.b']) == 64
        assert pk[p + 'wc0h.img'].numel() * 4 == 256 * 256 * 2
        assert pk[p + 'ff3.img'].numel() * 4 == 64 * d.r * 64 * 2
    assert torch.isfinite(pk.buf).all()
    # SiLU half-scaling folded into the images is exact in fp16
    w = sd['e_block_0.equi_update.coord_mlp.0.weight']
    img = pk['b0.wc0h.img'].view(torch.float16)
    assert torch.equal(image_to_matrix_h(img, 256, 256), (0.5 * w).half().float())


def test_roofline_model_matches_survey_numbers():
    """SURVEY.md 8d quotes 315 636 FLOP per edge per layer and 1 241 088 per atom per layer for the QM9 architecture."""
    d = dims_from_config(configs.NAMED['qm9_uncond']())
    k = roofline.per_edge_kernel_flops(d)
    assert k['jodo_attn'] + k['jodo_edge_update'] + k['jodo_equi'] == 315636
    f18, f29 = roofline.flops_alg(18, d), roofline.flops_alg(29, d)
    assert abs(f18 / 1.022e9 - 1) < 0.01 and abs(f29 / 2.448e9 - 1) < 0.01
    assert roofline.bytes_alg(18, d) < roofline.bytes_alg(29, d)
