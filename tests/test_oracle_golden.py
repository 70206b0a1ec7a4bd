"""The oracle (oracle/dgt_dense.py) is pinned against outputs of the unmodified reference
(fixtures written by oracle/make_golden.py)."""
import json
import os

import pytest
import torch

from helpers import FORWARD_CASES, GOLDEN, golden_weights, load_golden, oracle_forward
from jodo_b200 import configs
from jodo_b200.params import param_spec


@pytest.mark.parametrize('name', FORWARD_CASES)
def test_oracle_matches_reference_fp64(name):
    g, cfg = load_golden(name)
    sd = golden_weights(g, cfg)
    x, e = oracle_forward(sd, cfg, g['inputs'], torch.float64)
    rx, re = g['ref_fp64']
    assert x.shape == rx.shape and e.shape == re.shape
    # fp64 vs fp64: only summation order differs (dense grid vs scatter)
    assert float((x - rx).abs().max()) < 1e-10 * max(1.0, float(rx.abs().max()))
    assert float((e - re).abs().max()) < 1e-10 * max(1.0, float(re.abs().max()))


@pytest.mark.parametrize('name', ['qm9_first', 'qm9_selfcond', 'geom_l8'])
def test_oracle_fp32_close_to_reference_fp32(name):
    g, cfg = load_golden(name)
    sd = golden_weights(g, cfg)
    x, e = oracle_forward(sd, cfg, g['inputs'], torch.float32)
    rx, re = g['ref_fp32']
    assert float((x - rx).abs().max()) < 2e-5 * max(1.0, float(rx.abs().max()))
    assert float((e - re).abs().max()) < 2e-5 * max(1.0, float(re.abs().max()))


@pytest.mark.parametrize('name', ['qm9_selfcond', 'geom_l8'])
def test_oracle_block_intermediates(name):
    g, cfg = load_golden(name)
    sd = golden_weights(g, cfg)
    col = []
    oracle_forward(sd, cfg, g['inputs'], torch.float64, collect=col)
    m = g['inputs']['node_mask'].double()
    assert len(col) == len(g['blocks_fp64'])
    for i, ((h, _, pos), (rh, rpos)) in enumerate(zip(col, g['blocks_fp64'])):
        # the reference hook captures pos BEFORE the per-block CoM removal (mol_gnn.py:563-566)
        from oracle.dgt_dense import remove_mean_with_mask
        rp = remove_mean_with_mask(rpos.double(), m)
        assert float((h.float() - rh).abs().max()) < 1e-5 * max(1.0, float(rh.abs().max())), i
        assert float((pos - rp).abs().max()) < 1e-9, i


def test_masked_entries_exactly_zero_and_symmetric():
    g, cfg = load_golden('qm9_selfcond')
    rx, re = g['ref_fp32']
    inp = g['inputs']
    B, N = rx.shape[:2]
    em = inp['edge_mask'].reshape(B, N, N, 1)
    assert float((rx * (1 - inp['node_mask'])).abs().max()) == 0.0
    assert float((re * (1 - em)).abs().max()) == 0.0
    assert float((re - re.permute(0, 2, 1, 3)).abs().max()) == 0.0
    x, e = oracle_forward(golden_weights(g, cfg), cfg, inp, torch.float64)
    assert float((x * (1 - inp['node_mask'].double())).abs().max()) == 0.0
    assert float((e * (1 - em.double())).abs().max()) == 0.0


@pytest.mark.parametrize('cfg_name', ['qm9_uncond', 'qm9_cond', 'geom_l8', 'geom_l10', 'geom_large', 'moses_2d', 'qm9_cond_multi'])
def test_param_tree_matches_reference(cfg_name):
    with open(os.path.join(GOLDEN, f'param_tree_{cfg_name}.json')) as f:
        ref = [(k, tuple(s)) for k, s in json.load(f)]
    assert param_spec(configs.NAMED[cfg_name]()) == ref
