"""PyG 2.1 utils used on the hot path: softmax (layers.py:178), dense_to_sparse (mol_gnn.py:514)."""
import torch


def dense_to_sparse(adj):
    """PyG 2.1 semantics: batched [B,N,N] adjacency -> edge_index [2,E] with node ids b*N+i,
    enumerated in (b, row, col) lexicographic order, and the selected values."""
    assert adj.dim() in (2, 3)
    if adj.dim() == 2:
        idx = adj.nonzero(as_tuple=True)
        return torch.stack(idx, dim=0), adj[idx]
    idx = adj.nonzero(as_tuple=True)
    n = adj.size(1)
    row = idx[0] * n + idx[1]
    col = idx[0] * n + idx[2]
    return torch.stack([row, col], dim=0), adj[idx]


def softmax(src, index=None, ptr=None, num_nodes=None, dim=0):
    """PyG 2.1 grouped softmax: exp(src - max_g) / (sum_g exp(.) + 1e-16), groups given by index."""
    assert ptr is None and dim == 0
    n = int(index.max()) + 1 if num_nodes is None else num_nodes
    shape = (n,) + tuple(src.shape[1:])
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    src_max = torch.full(shape, float("-inf"), dtype=src.dtype, device=src.device)
    src_max = src_max.scatter_reduce(0, idx, src.detach(), reduce="amax", include_self=True)
    out = (src - src_max.gather(0, idx)).exp()
    out_sum = torch.zeros(shape, dtype=src.dtype, device=src.device).scatter_add_(0, idx, out)
    return out / (out_sum.gather(0, idx) + 1e-16)


def to_dense_batch(*a, **k):  # imported by unrelated reference modules only
    raise NotImplementedError


def to_dense_adj(*a, **k):
    raise NotImplementedError
