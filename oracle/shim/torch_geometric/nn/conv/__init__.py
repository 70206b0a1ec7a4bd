"""MessagePassing(aggr='add', node_dim=0) as PyG 2.1 runs it for layers.py:152."""
import inspect
import torch


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr="add", node_dim=0, **kwargs):
        super().__init__()
        assert aggr == "add" and node_dim == 0
        self.aggr = aggr
        self.node_dim = node_dim

    def propagate(self, edge_index, size=None, **kwargs):
        src, dst = edge_index[0], edge_index[1]  # x_j = x[edge_index[0]], x_i = x[edge_index[1]]
        n_nodes = None
        for v in kwargs.values():
            if torch.is_tensor(v) and v.dim() > 0:
                n_nodes = v.size(0)
                break
        args = {}
        for name in inspect.signature(self.message).parameters:
            if name.endswith("_i"):
                if name == "size_i":
                    args[name] = n_nodes
                else:
                    args[name] = kwargs[name[:-2]][dst]
            elif name.endswith("_j"):
                args[name] = kwargs[name[:-2]][src]
            elif name == "index":
                args[name] = dst
            elif name == "ptr":
                args[name] = None
            else:
                args[name] = kwargs[name]
        msg = self.message(**args)
        out = torch.zeros((n_nodes,) + tuple(msg.shape[1:]), dtype=msg.dtype, device=msg.device)
        idx = dst.view(-1, *([1] * (msg.dim() - 1))).expand_as(msg)
        return out.scatter_add_(0, idx, msg)


class GINEConv(torch.nn.Module):  # import-only stub (models/cdgs.py)
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError


class GATConv(GINEConv):
    pass
