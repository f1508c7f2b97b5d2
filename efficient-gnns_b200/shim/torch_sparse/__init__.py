"""`torch_sparse` surface used by the reference (mag_pyg/gnn.py:13; arxiv_pyg via T.ToSparseTensor)."""
import efficient_gnns_b200  # noqa: F401
from efficient_gnns_b200.sparse import SparseTensor  # noqa: F401
from efficient_gnns_b200.ops import matmul  # noqa: F401
from efficient_gnns_b200 import torch_ops as _torch_ops  # noqa: F401  registers torch.ops.b200gnn.{spmm_sum, spmm_mean, ind2ptr, ...}

__version__ = "0.6.9+b200gnn"
