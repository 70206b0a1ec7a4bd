"""Property classifier (reference cond_gen/model.py:26-220, cond_gen/utils.py:18-40).

CPU: the oracle restatement (oracle/egnn_dense.py) is pinned on the fixture generated from the UNMODIFIED reference
class (oracle/make_golden_egnn.py) and, when the reference sources are present (/root/reference or the staged copy),
on the live class with other sizes; the product module mirrors the reference's parameter tree and edge-list builder.
GPU: the CUDA forward (through the C ABI) against the reference outputs of the fixture and against the oracle on a
larger ragged batch."""
import os

import pytest
import torch

from jodo_b200 import _lib
from jodo_b200.classifier import EGNN, egnn_param_spec, egnn_synth_state_dict, get_adj_matrix_fn
from oracle import ref_loader
from oracle.egnn_dense import adj_matrix, egnn_forward

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'egnn_qm9.pt')
CASES = ['egnn_qm9', 'egnn_small_attr']
TOL = 5e-3          # of the largest |pred|: fp16 tensor-core operands (the mantissa of tf32), fp32 accumulation; the scalar
                    # is a sum over atoms and H hidden units with cancellation (measured 1e-3 .. 2.8e-3)


def _case(name):
    c = torch.load(GOLD)[name]
    a = c['args']
    spec = egnn_param_spec(5, a['nf'], a['n_layers'], a['attention'], a['node_attr'])
    sd = egnn_synth_state_dict(spec, **c['weights'])
    return c, a, spec, sd


def _batch(n_list, in_nf=5, seed=3):
    g = torch.Generator().manual_seed(seed)
    B, N = len(n_list), max(n_list)
    nm = torch.zeros(B, N)
    for i, n in enumerate(n_list):
        nm[i, :n] = 1
    em = nm.unsqueeze(1) * nm.unsqueeze(2) * (~torch.eye(N, dtype=torch.bool)).unsqueeze(0)
    x = torch.randn(B, N, 3, generator=g) * 1.5 * nm[..., None]
    h0 = torch.nn.functional.one_hot(torch.randint(0, in_nf, (B, N), generator=g), in_nf).float() * nm[..., None]
    return dict(h0=h0.reshape(B * N, in_nf), x=x.reshape(B * N, 3), node_mask=nm.reshape(B * N, 1),
                edge_mask=em.reshape(B * N * N, 1), n_nodes=N)


@pytest.mark.parametrize('name', CASES)
def test_oracle_pinned_on_reference_fixture(name):
    c, a, spec, sd = _case(name)
    assert [k for k, _ in spec] == c['param_names']
    i = c['inputs']
    y = egnn_forward({k: v.double() for k, v in sd.items()}, i['h0'], i['x'], i['node_mask'], i['edge_mask'], i['n_nodes'],
                     a['n_layers'], a['attention'], a['node_attr'])
    assert float((y - c['ref_fp64']).abs().max()) < 1e-10
    y32 = egnn_forward(sd, i['h0'], i['x'], i['node_mask'], i['edge_mask'], i['n_nodes'], a['n_layers'], a['attention'], a['node_attr'])
    assert float((y32 - c['ref_fp32']).abs().max()) < 1e-4 * float(c['ref_fp32'].abs().max())


@pytest.mark.skipif(not ref_loader.available(), reason='reference sources not present')
@pytest.mark.parametrize('nf,L,att,na', [(64, 3, True, True), (128, 2, False, False)])
def test_oracle_against_live_reference(nf, L, att, na):
    cg = ref_loader.load_cond_gen()
    spec = egnn_param_spec(5, nf, L, att, na)
    sd = egnn_synth_state_dict(spec, seed=5, gain=1.3)
    ref = cg.model.EGNN(in_node_nf=5, in_edge_nf=0, hidden_nf=nf, device='cpu', n_layers=L, attention=att, node_attr=na).double().eval()
    ref.load_state_dict({k: v.double() for k, v in sd.items()}, strict=True)
    b = _batch([7, 2, 11, 5])
    edges = cg.utils.get_adj_matrix_fn()(b['n_nodes'], 4, 'cpu')
    with torch.no_grad():
        want = ref(edges=edges, edge_attr=None, **{k: (v.double() if torch.is_tensor(v) else v) for k, v in b.items()})
    got = egnn_forward({k: v.double() for k, v in sd.items()}, b['h0'], b['x'], b['node_mask'], b['edge_mask'], b['n_nodes'], L, att, na)
    assert float((got - want).abs().max()) < 1e-10
    # the edge-list builder: same lists as the reference's triple loop
    mine = get_adj_matrix_fn()(b['n_nodes'], 4, 'cpu')
    assert torch.equal(mine[0], edges[0]) and torch.equal(mine[1], edges[1])


def test_adj_matrix_matches_reference_fixture():
    want = torch.load(GOLD)['adj_4_3']
    for got in (get_adj_matrix_fn()(4, 3, 'cpu'), adj_matrix(4, 3)):
        assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1]) and got[0].dtype == torch.int64
    fn = get_adj_matrix_fn()
    assert fn(4, 3, 'cpu') is fn(4, 3, 'cpu')                   # cached (the reference rebuilds the lists on every call)


@pytest.mark.parametrize('name', CASES)
def test_parameter_tree_mirrors_reference(name):
    c, a, spec, sd = _case(name)
    m = EGNN(in_node_nf=5, in_edge_nf=0, hidden_nf=a['nf'], device='cpu', n_layers=a['n_layers'], attention=a['attention'],
             node_attr=a['node_attr'])
    assert [k for k, _ in m.named_parameters()] == c['param_names']           # names AND registration order
    assert list(m.state_dict().keys()) == c['param_names']
    m.load_state_dict(sd, strict=True)
    for k, shape in spec:
        assert tuple(m.state_dict()[k].shape) == tuple(shape)
    # the packed operand images (torch emulation of jodo_pack_weights on the CPU): the hoisted edge_mlp.0 split
    pk = m._weights()
    H = a['nf']
    w0 = sd['gcl_0.edge_mlp.0.weight']
    assert torch.equal(pk['l0.wr'][:H], w0[:, 2 * H]) and pk.meta['l0.pq'] == dict(N=2 * H, K=H, NT=128 if H % 128 == 0 else 64)
    from jodo_b200.pack import image_to_matrix_h
    nt = pk.meta['l0.pq']['NT']
    img = pk['l0.pq.img'].view(torch.float16).reshape(2 * H // nt, -1)
    pq = torch.cat([image_to_matrix_h(img[t], nt, H) for t in range(2 * H // nt)])
    assert torch.equal(pq, torch.cat([w0[:, :H], w0[:, H:2 * H]]).half().float())


def test_no_cpu_fallback_and_unsupported_sizes():
    m = EGNN(in_node_nf=5, in_edge_nf=0, hidden_nf=64, device='cpu', n_layers=1, attention=True, node_attr=False).eval()
    b = _batch([3, 2])
    with pytest.raises(_lib.JodoError):
        m(edges=None, edge_attr=None, **b)
    with pytest.raises(NotImplementedError):
        EGNN(in_node_nf=5, in_edge_nf=0, hidden_nf=100, device='cpu')
    with pytest.raises(NotImplementedError):
        EGNN(in_node_nf=5, in_edge_nf=2, hidden_nf=64, device='cpu')


# ---------------------------------------------------------------------------------------------------- GPU
def _cuda_model(a, sd):
    m = EGNN(in_node_nf=5, in_edge_nf=0, hidden_nf=a['nf'], device='cuda', n_layers=a['n_layers'], attention=a['attention'],
             node_attr=a['node_attr']).eval()
    m.load_state_dict(sd, strict=True)
    return m


@pytest.mark.gpu
@pytest.mark.parametrize('name', CASES)
def test_cuda_classifier_against_reference_fixture(name):
    c, a, spec, sd = _case(name)
    m = _cuda_model(a, sd)
    i = c['inputs']
    dev = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in i.items()}
    y = m(edges=get_adj_matrix_fn()(i['n_nodes'], len(c['n_per_mol']), 'cuda'), edge_attr=None, **dev)
    torch.cuda.synchronize()
    err = float((y.double().cpu() - c['ref_fp64']).abs().max() / c['ref_fp64'].abs().max())
    print(f'{name}: rel err vs reference fp64 {err:.2e}')
    assert y.shape == c['ref_fp64'].shape and err < TOL


@pytest.mark.gpu
def test_cuda_classifier_ragged_batch_against_oracle():
    """256 molecules from 1 to 29 atoms (QM9 sizes) incl. single atoms (no edges): CUDA vs the fp64 oracle."""
    c, a, spec, sd = _case('egnn_qm9')
    m = _cuda_model(a, sd)
    g = torch.Generator().manual_seed(11)
    n_list = [int(v) for v in torch.randint(1, 30, (256,), generator=g)]
    b = _batch(n_list, seed=12)
    want = egnn_forward({k: v.double() for k, v in sd.items()}, b['h0'], b['x'], b['node_mask'], b['edge_mask'], b['n_nodes'],
                        a['n_layers'], a['attention'], a['node_attr'])
    dev = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()}
    y = m(edges=None, edge_attr=None, **dev)
    y2 = m(edges=None, edge_attr=None, **dev)                    # second call: cached plan, same result
    torch.cuda.synchronize()
    err = float((y.double().cpu() - want).abs().max() / want.abs().max())
    print(f'ragged 256: rel err vs oracle {err:.2e}')
    assert err < TOL and torch.equal(y, y2)
    # an in-place weight update is picked up (parameter versions key the packed images)
    with torch.no_grad():
        m.get_parameter('graph_dec.2.bias').add_(1.0)
    y3 = m(edges=None, edge_attr=None, **dev)
    assert float((y3 - y - 1.0).abs().max()) < 1e-4
