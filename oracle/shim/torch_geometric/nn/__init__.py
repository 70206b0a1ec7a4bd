"""Names that /root/reference/models/cdgs.py imports at module import time (never executed here)."""
from torch.nn import Linear  # noqa: F401
from .conv import MessagePassing, GINEConv, GATConv  # noqa: F401


def global_mean_pool(*a, **k):
    raise NotImplementedError


def global_add_pool(*a, **k):
    raise NotImplementedError
