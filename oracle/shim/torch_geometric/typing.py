"""Type aliases used by /root/reference/models/layers.py:3."""
from typing import Optional, Tuple
from torch import Tensor

Adj = Tensor
OptTensor = Optional[Tensor]
PairTensor = Tuple[Tensor, Tensor]
OptPairTensor = Tuple[Tensor, Optional[Tensor]]
Size = Optional[Tuple[int, int]]
