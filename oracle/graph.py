"""Integer graph preparation, restated with numpy (bit-exact targets).

Follows SURVEY.md Appendix A.1 / A.7 for the calls the reference makes at
arxiv_pyg/gnn.py:236-249 (ToSparseTensor, to_symmetric, coo, subgraph) and
mag_pyg/gnn.py:151,333 (SparseTensor(row=col,col=row), to_undirected).
"""
from __future__ import annotations

import numpy as np


def to_sparse_adj_t(edge_index: np.ndarray, num_nodes: int):
    """T.ToSparseTensor(): adj_t with row = message target, col = message source, sorted by (row, col).
    Returns (row, col, rowptr) as int64."""
    src, dst = edge_index[0].astype(np.int64), edge_index[1].astype(np.int64)
    order = np.lexsort((src, dst))          # primary dst (=adj_t row), secondary src (=adj_t col)
    row, col = dst[order], src[order]
    return row, col, ind2ptr(row, num_nodes)


def ind2ptr(row_sorted: np.ndarray, n: int) -> np.ndarray:
    ptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(ptr, row_sorted + 1, 1)
    return np.cumsum(ptr)


def coalesce(row: np.ndarray, col: np.ndarray, n_cols: int, val: np.ndarray | None = None):
    """Sort by (row, col) and drop duplicates (values summed)."""
    order = np.lexsort((col, row))
    row, col = row[order], col[order]
    val = None if val is None else val[order]
    key = row * n_cols + col
    keep = np.ones(key.shape[0], dtype=bool)
    keep[1:] = key[1:] != key[:-1]
    if val is not None:
        seg = np.cumsum(keep) - 1
        out = np.zeros(int(seg[-1]) + 1 if seg.size else 0, dtype=val.dtype)
        np.add.at(out, seg, val)
        val = out
    return row[keep], col[keep], val


def to_symmetric(row: np.ndarray, col: np.ndarray, n: int):
    """adj_t.to_symmetric(): union with the transpose, coalesced (arxiv_pyg/gnn.py:240)."""
    r, c, _ = coalesce(np.concatenate([row, col]), np.concatenate([col, row]), n)
    return r, c


def to_undirected(edge_index: np.ndarray, n: int) -> np.ndarray:
    """torch_geometric.utils.to_undirected (mag_pyg/gnn.py:333): both directions, coalesced, sorted by (row, col)."""
    r, c = to_symmetric(edge_index[0].astype(np.int64), edge_index[1].astype(np.int64), n)
    return np.stack([r, c])


def fill_diag(row: np.ndarray, col: np.ndarray, val: np.ndarray, n: int, fill: float = 1.0):
    """fill_diag(adj_t, 1.): diagonal entries replaced / added; result sorted by (row, col)."""
    off = row != col
    d = np.arange(n, dtype=np.int64)
    r = np.concatenate([row[off], d])
    c = np.concatenate([col[off], d])
    v = np.concatenate([val[off], np.full(n, fill, dtype=val.dtype)])
    order = np.lexsort((c, r))
    return r[order], c[order], v[order]


def gcn_norm(row: np.ndarray, col: np.ndarray, n: int):
    """PyG gcn_norm on a value-less SparseTensor: A+I, D^-1/2 (A+I) D^-1/2, fp32."""
    val = np.ones(row.shape[0], dtype=np.float32)
    r, c, v = fill_diag(row, col, val, n, 1.0)
    deg = np.zeros(n, dtype=np.float32)
    np.add.at(deg, r, v)
    with np.errstate(divide="ignore"):
        dis = np.power(deg, np.float32(-0.5)).astype(np.float32)
    dis[np.isinf(dis)] = 0.0
    v = (dis[r] * v * dis[c]).astype(np.float32)
    return r, c, v


def csr2csc(row: np.ndarray, col: np.ndarray):
    """Permutation sorting the non-zeros by (col, row); returns (perm, colptr-ready sorted col)."""
    return np.lexsort((row, col))


def subgraph(subset: np.ndarray, edge_index: np.ndarray, relabel_nodes: bool = True):
    """torch_geometric.utils.subgraph (arxiv_pyg/gnn.py:249): induced subgraph, order preserved,
    new ids = positions in ``subset``."""
    num_nodes = int(edge_index.max()) + 1 if edge_index.size else 0
    num_nodes = max(num_nodes, int(subset.max()) + 1 if subset.size else 0)
    n_mask = np.zeros(num_nodes, dtype=bool)
    n_mask[subset] = True
    mask = n_mask[edge_index[0]] & n_mask[edge_index[1]]
    ei = edge_index[:, mask]
    if relabel_nodes:
        n_idx = np.zeros(num_nodes, dtype=np.int64)
        n_idx[subset] = np.arange(subset.shape[0])
        ei = n_idx[ei]
    return ei, mask
