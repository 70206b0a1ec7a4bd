"""ORACLE (test infrastructure, not product code): CPU restatement of the reference's property classifier --
the EGNN that scores conditional samples (reference cond_gen/model.py:26-220, called from sampling.py:363-367) --
in dense per-molecule form, any float dtype.  Only tests/ import this.

Parity pinned: tests/test_classifier.py compares this restatement with the UNMODIFIED reference class
(cond_gen/model.py imports nothing but torch, so it runs as is, from /root/reference or the staged copy under
oracle/_ref/reference) and with the committed fixture tests/golden/egnn_qm9.pt (generator: oracle/make_golden_egnn.py).

The reference works on the flattened full graph: edges = every (i, j) of every molecule INCLUDING i == j and padding,
rows = i, cols = j, (b, i, j) order (cond_gen/utils.py:18-40), with edge_mask = node_mask x node_mask minus the
diagonal (sampling.py:333-337) multiplied onto the edge features (model.py:207).  Dense form, per molecule:

    h = embedding(h0)                                                     model.py:56
    per layer (E_GCL_mask.forward, model.py:201-216; coordinates are NOT updated, :211):
        radial[i, j] = |x_i - x_j|^2                                      model.py:164-167
        m = SiLU(W2 SiLU(W0 cat[h_i, h_j, radial] + b0) + b2)             edge_mlp, model.py:93-97, 126-131
        m = m * sigmoid(w_a m + b_a)            (attention)               model.py:132-134
        m = m * edge_mask                                                 model.py:207
        agg_i = sum_j m[i, j]                   (unsorted_segment_sum on rows)   model.py:138-139
        h = h + node_mlp(cat[h, agg(, h0)])     (recurrent)               model.py:140-147
    h = node_dec(h) * node_mask ; pred = graph_dec(sum_i h_i)             model.py:64-70
"""
import torch
import torch.nn.functional as F


def _lin(sd, name, x):
    return F.linear(x, sd[name + '.weight'], sd.get(name + '.bias'))


def egnn_forward(sd, h0, x, node_mask, edge_mask, n_nodes, n_layers, attention=True, node_attr=False):
    """sd: the reference EGNN state dict (any float dtype); h0 [B*N, in], x [B*N, 3], node_mask [B*N, 1],
    edge_mask [B*N*N, 1]; returns pred [B]."""
    N = n_nodes
    B = h0.shape[0] // N
    dt = sd['embedding.weight'].dtype
    h0 = h0.to(dt).reshape(B, N, -1)
    x = x.to(dt).reshape(B, N, 3)
    nm = node_mask.to(dt).reshape(B, N, 1)
    em = edge_mask.to(dt).reshape(B, N, N, 1)
    h = _lin(sd, 'embedding', h0)
    H = h.shape[-1]
    diff = x[:, :, None, :] - x[:, None, :, :]                    # coord[row] - coord[col]
    radial = (diff ** 2).sum(-1, keepdim=True)                    # [B, N, N, 1]
    for l in range(n_layers):
        p = f'gcl_{l}.'
        hi = h[:, :, None, :].expand(B, N, N, H)                  # source = h[row]
        hj = h[:, None, :, :].expand(B, N, N, H)                  # target = h[col]
        m = F.silu(_lin(sd, p + 'edge_mlp.0', torch.cat([hi, hj, radial], dim=-1)))
        m = F.silu(_lin(sd, p + 'edge_mlp.2', m))
        if attention:
            m = m * torch.sigmoid(_lin(sd, p + 'att_mlp.0', m))
        m = m * em
        agg = m.sum(dim=2)                                        # over cols j, onto row i
        cat = [h, agg] + ([h0] if node_attr else [])
        out = _lin(sd, p + 'node_mlp.2', F.silu(_lin(sd, p + 'node_mlp.0', torch.cat(cat, dim=-1))))
        h = h + out
    h = _lin(sd, 'node_dec.2', F.silu(_lin(sd, 'node_dec.0', h))) * nm
    g = h.sum(dim=1)
    return _lin(sd, 'graph_dec.2', F.silu(_lin(sd, 'graph_dec.0', g))).squeeze(1)


def adj_matrix(n_nodes, batch_size):
    """The full edge list of cond_gen/utils.py:18-40 (get_adj_matrix): rows[k] = i + b n, cols[k] = j + b n for every
    (b, i, j) in lexicographic order, self-loops included."""
    b = torch.arange(batch_size).view(-1, 1, 1)
    i = torch.arange(n_nodes).view(1, -1, 1)
    j = torch.arange(n_nodes).view(1, 1, -1)
    rows = (i + b * n_nodes).expand(batch_size, n_nodes, n_nodes).reshape(-1)
    cols = (j + b * n_nodes).expand(batch_size, n_nodes, n_nodes).reshape(-1)
    return [rows.long(), cols.long()]
