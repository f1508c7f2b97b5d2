from efficient_gnns_b200.nn import softmax, subgraph, to_undirected  # noqa: F401
from . import hetero  # noqa: F401


def _unused(name):
    def f(*a, **k):
        raise NotImplementedError(f"torch_geometric.utils.{name} is imported but never called by the reference")
    return f


# imported by criterion.py:5 but never used there
to_dense_adj = _unused("to_dense_adj")
negative_sampling = _unused("negative_sampling")
add_self_loops = _unused("add_self_loops")
