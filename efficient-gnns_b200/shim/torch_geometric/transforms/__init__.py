"""T.ToSparseTensor (arxiv_pyg/gnn.py:236-237)."""
from efficient_gnns_b200.sparse import SparseTensor, device_argsort


class ToSparseTensor:
    def __init__(self, remove_edge_index: bool = True, fill_cache: bool = True):
        self.remove_edge_index = remove_edge_index

    def __call__(self, data):
        row, col = data.edge_index
        n = data.num_nodes
        perm = device_argsort(col, row, n, n)       # (col*n + row).argsort(): the device radix sort on CUDA inputs
        data.adj_t = SparseTensor(row=col[perm], col=row[perm], sparse_sizes=(n, n), is_sorted=True)
        if self.remove_edge_index:
            data.edge_index = None
        return data
