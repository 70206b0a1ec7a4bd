"""Post-processing of the final sampler state: jodo_b200/postprocess.py against the outputs of
the reference's post_process / mol_process (sampling.py:12-97) recorded in tests/golden/postprocess.pt.
Integer outputs (atom types, charges, bond orders) must be bit-exact."""
import os

import pytest
import torch

from helpers import GOLDEN
from jodo_b200 import configs
from jodo_b200.postprocess import mol_process, mol_process_2d, post_process, post_process_2d


@pytest.mark.parametrize('cfg_name', ['qm9_uncond', 'geom_l8'])
def test_post_process_matches_reference(cfg_name):
    g = torch.load(os.path.join(GOLDEN, 'postprocess.pt'), weights_only=False)[cfg_name]
    cfg = configs.NAMED[cfg_name]()
    pos, one_hot, fc, edge = post_process(cfg, g['xh'].clone(), g['node_mask'], g['edge_x'].clone(), g['edge_mask'])
    assert torch.equal(one_hot, g['one_hot'])
    assert torch.equal(fc, g['fc'])
    assert torch.equal(edge, g['edge'])
    assert torch.equal(pos, g['pos'])
    assert set(edge.unique().tolist()) <= {0., 1., 2., 3., 4.}
    mols = mol_process(one_hot, pos, fc, g['n_nodes'], edge)
    assert len(mols) == len(g['mols'])
    for (p, a, e, c), (rp, ra, re_, rc) in zip(mols, g['mols']):
        assert torch.equal(p, rp) and torch.equal(a, ra) and torch.equal(e, re_) and torch.equal(c, rc)


def test_post_process_2d_matches_reference():
    """post_process_2D / mol_process_2D of the reference (sampling.py:35-50, 100-144) on a MOSES-shaped final state
    (7 atom types, no charges, 3 compressed edge channels incl. aromatic)."""
    g = torch.load(os.path.join(GOLDEN, 'postprocess.pt'), weights_only=False)['moses_2d']
    cfg = configs.NAMED['moses_2d']()
    one_hot, fc, edge = post_process_2d(cfg, g['xh'].clone(), g['node_mask'], g['edge_x'].clone(), g['edge_mask'])
    assert torch.equal(one_hot, g['one_hot']) and torch.equal(fc, g['fc']) and torch.equal(edge, g['edge'])
    assert 4. in edge.unique().tolist()                        # aromatic bonds occur in the fixture
    mols = mol_process_2d(one_hot, fc, g['n_nodes'], edge)
    assert len(mols) == len(g['mols'])
    for (p, a, e, c), (rp, ra, re_, rc) in zip(mols, g['mols']):
        assert p is None and rp is None
        assert torch.equal(a, ra) and torch.equal(e, re_) and torch.equal(c, rc)


def test_mol_process_without_edges_and_charges():
    one_hot = torch.eye(5)[torch.tensor([[0, 1, 2], [3, 4, 0]])]
    x = torch.randn(2, 3, 3)
    mols = mol_process(one_hot, x, torch.zeros(0), [2, 3])
    assert [m[1].tolist() for m in mols] == [[0, 1], [3, 4, 0]]
    assert mols[0][0].shape == (2, 3) and mols[1][0].shape == (3, 3)


@pytest.mark.gpu
def test_post_process_on_device_matches_host():
    g = torch.load(os.path.join(GOLDEN, 'postprocess.pt'), weights_only=False)['geom_l8']
    cfg = configs.NAMED['geom_l8']()
    d = lambda t: t.cuda()
    pos, one_hot, fc, edge = post_process(cfg, d(g['xh']), d(g['node_mask']), d(g['edge_x']), d(g['edge_mask']))
    assert torch.equal(one_hot.cpu(), g['one_hot']) and torch.equal(fc.cpu(), g['fc']) and torch.equal(edge.cpu(), g['edge'])
    mols = mol_process(one_hot, pos, fc, g['n_nodes'], edge)          # one D2H per tensor
    for (p, a, e, c), (rp, ra, re_, rc) in zip(mols, g['mols']):
        assert torch.equal(a, ra) and torch.equal(e, re_) and torch.equal(c, rc)
        assert float((p - rp).abs().max()) < 1e-6
