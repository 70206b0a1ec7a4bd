"""Stand-in for torch_scatter==2.0.9 (requirements.txt:9): scatter(reduce=add|sum|max|mean).
TEST INFRASTRUCTURE ONLY (see oracle/README.md)."""
import torch


def scatter(src, index, dim=0, out=None, dim_size=None, reduce="sum"):
    assert out is None
    dim = dim % src.dim()
    if dim_size is None:
        dim_size = int(index.max()) + 1
    shape = list(src.shape)
    shape[dim] = dim_size
    view = [1] * src.dim()
    view[dim] = -1
    idx = index.view(view).expand_as(src)
    if reduce in ("add", "sum"):
        return torch.zeros(shape, dtype=src.dtype, device=src.device).scatter_add_(dim, idx, src)
    if reduce == "mean":
        s = torch.zeros(shape, dtype=src.dtype, device=src.device).scatter_add_(dim, idx, src)
        c = torch.zeros(shape, dtype=src.dtype, device=src.device).scatter_add_(dim, idx, torch.ones_like(src))
        return s / c.clamp(min=1)
    if reduce == "max":
        o = torch.zeros(shape, dtype=src.dtype, device=src.device)
        return o.scatter_reduce(dim, idx, src, reduce="amax", include_self=False)
    raise ValueError(reduce)


def scatter_add(src, index, dim=0, out=None, dim_size=None):
    return scatter(src, index, dim, out, dim_size, "sum")


def scatter_mean(src, index, dim=0, out=None, dim_size=None):
    return scatter(src, index, dim, out, dim_size, "mean")
