"""`torch_geometric` surface used by the reference, backed by efficient_gnns_b200."""
import efficient_gnns_b200  # noqa: F401
from . import nn, utils, transforms, data, datasets  # noqa: F401

__version__ = "1.7.0+b200gnn"
