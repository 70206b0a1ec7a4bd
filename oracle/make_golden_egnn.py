"""ORACLE tooling: generate tests/golden/egnn_qm9.pt from the UNMODIFIED reference property classifier
(cond_gen/model.py EGNN; imports only torch, runs as is).  Run in the build container:

    python -m oracle.make_golden_egnn

The fixture holds the (seed, gain) of the synthetic weights, the reference's parameter names, a ragged batch built the way sampling.py:330-337 builds masks, the reference output in fp32 and fp64,
and the edge list of cond_gen/utils.get_adj_matrix for a small case."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from jodo_b200.classifier import egnn_param_spec, egnn_synth_state_dict  # noqa: E402


def make_inputs(B, N, n_nodes, in_nf, gen):
    node_mask = torch.zeros(B, N)
    for i in range(B):
        node_mask[i, :n_nodes[i]] = 1
    edge_mask = node_mask.unsqueeze(1) * node_mask.unsqueeze(2)
    edge_mask *= ~torch.eye(N, dtype=torch.bool).unsqueeze(0)
    x = torch.randn(B, N, 3, generator=gen) * 1.5 * node_mask[..., None]
    x = x - (x.sum(1, keepdim=True) / node_mask.sum(1).view(B, 1, 1)) * node_mask[..., None]
    types = torch.randint(0, in_nf, (B, N), generator=gen)
    one_hot = torch.nn.functional.one_hot(types, in_nf).float() * node_mask[..., None]
    return dict(h0=one_hot.reshape(B * N, in_nf), x=x.reshape(B * N, 3), node_mask=node_mask.reshape(B * N, 1),
                edge_mask=edge_mask.reshape(B * N * N, 1), n_nodes=N)


def main():
    cg = ref_loader.load_cond_gen()
    out = {}
    for name, (nf, L, att, na) in {'egnn_qm9': (128, 7, True, False), 'egnn_small_attr': (64, 2, False, True)}.items():
        model = cg.model.EGNN(in_node_nf=5, in_edge_nf=0, hidden_nf=nf, device='cpu', n_layers=L, coords_weight=1.0,
                              attention=att, node_attr=na).eval()
        # weights: the seeded synthetic init of jodo_b200.classifier (gain 1.5: gates and activations leave their linear
        # range), regenerated from (seed, gain) by the tests instead of being stored
        spec = egnn_param_spec(5, nf, L, att, na)
        assert [k for k, _ in model.named_parameters()] == [k for k, _ in spec]
        model.load_state_dict(egnn_synth_state_dict(spec, seed=42, gain=1.5), strict=True)
        gen = torch.Generator().manual_seed(7)
        n_nodes = [9, 3, 14, 1, 12, 14]
        inp = make_inputs(len(n_nodes), max(n_nodes), n_nodes, 5, gen)
        edges = cg.utils.get_adj_matrix_fn()(inp['n_nodes'], len(n_nodes), 'cpu')
        with torch.no_grad():
            y32 = model(edges=edges, edge_attr=None, **inp)
            m64 = model.double()
            y64 = m64(edges=edges, edge_attr=None, **{k: (v.double() if torch.is_tensor(v) else v) for k, v in inp.items()})
        out[name] = dict(args=dict(nf=nf, n_layers=L, attention=att, node_attr=na), weights=dict(seed=42, gain=1.5), inputs=inp,
                         n_per_mol=n_nodes, ref_fp32=y32.float(), ref_fp64=y64, param_names=[k for k, _ in model.named_parameters()])
    e = cg.utils.get_adj_matrix_fn()(4, 3, 'cpu')
    out['adj_4_3'] = [e[0].clone(), e[1].clone()]
    path = os.path.join(ROOT, 'tests', 'golden', 'egnn_qm9.pt')
    torch.save(out, path)
    print('wrote', path, os.path.getsize(path), 'bytes')
    for k in ('egnn_qm9', 'egnn_small_attr'):
        print(k, out[k]['ref_fp64'])


if __name__ == '__main__':
    main()
