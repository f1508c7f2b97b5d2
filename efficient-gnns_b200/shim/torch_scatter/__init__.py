"""`torch_scatter` surface: scatter(src, index, dim, dim_size=, reduce=) for sum / add / mean."""
import efficient_gnns_b200  # noqa: F401
from efficient_gnns_b200.nn import scatter  # noqa: F401


def scatter_add(src, index, dim=0, out=None, dim_size=None):
    return scatter(src, index, dim, dim_size, "sum")


def scatter_mean(src, index, dim=0, out=None, dim_size=None):
    return scatter(src, index, dim, dim_size, "mean")


__version__ = "2.0.6+b200gnn"
