"""Sparse operators restated in plain PyTorch on the CPU (differentiable through autograd).

SpMM is given in three independent forms that must agree before any is trusted
(SURVEY.md §8c): (i) gather + scatter_add_ (what upstream torch_scatter.scatter_sum
executes), (ii) torch.sparse_csr @ dense, (iii) dense A @ X for tiny graphs.
"""
from __future__ import annotations

from typing import Optional

import torch


def scatter(src: torch.Tensor, index: torch.Tensor, dim_size: int, reduce: str = "sum") -> torch.Tensor:
    """torch_scatter.scatter(src, index, dim=0, dim_size, reduce) for sum/mean/max."""
    shape = (dim_size,) + tuple(src.shape[1:])
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    if reduce in ("sum", "add"):
        return torch.zeros(shape, dtype=src.dtype).scatter_add_(0, idx, src)
    if reduce == "mean":
        s = torch.zeros(shape, dtype=src.dtype).scatter_add_(0, idx, src)
        cnt = torch.zeros(dim_size, dtype=src.dtype).scatter_add_(0, index, torch.ones_like(index, dtype=src.dtype))
        return s / cnt.clamp(min=1).view(-1, *([1] * (src.dim() - 1)))
    if reduce == "max":
        out = torch.full(shape, float("-inf"), dtype=src.dtype).scatter_reduce_(0, idx, src, reduce="amax", include_self=True)
        return torch.where(torch.isinf(out) & (out < 0), torch.zeros_like(out), out)
    raise ValueError(reduce)


def spmm_scatter(row, col, val: Optional[torch.Tensor], x, n_rows: int, reduce: str = "sum"):
    """Form (i): x_j = x[col]; out = scatter(val*x_j, row)."""
    msg = x.index_select(0, col)
    if val is not None:
        msg = msg * val.view(-1, 1)
    return scatter(msg, row, n_rows, reduce)


class _CsrMatmul(torch.autograd.Function):
    """A @ x with the backward restated the way upstream torch_sparse does it: dX = A^T @ dY through a
    PRE-BUILT CSR of A^T (the cached colptr / csr2csc view), instead of autograd re-deriving a transpose per call."""

    @staticmethod
    def forward(ctx, x, A, At):
        ctx.At = At
        return A @ x

    @staticmethod
    def backward(ctx, g):
        return ctx.At @ g.contiguous(), None, None


_CSR_CACHE = {}


def _csr_pair(rowptr, col, v, n_rows, n_cols):
    """(A, A^T) as torch sparse CSR tensors, cached per (rowptr, col, val) identity like upstream's SparseStorage."""
    key = (rowptr.data_ptr(), col.data_ptr(), v.data_ptr(), v.dtype, n_rows, n_cols)
    hit = _CSR_CACHE.get(key)
    if hit is None:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            A = torch.sparse_csr_tensor(rowptr, col, v, size=(n_rows, n_cols))
            row = torch.repeat_interleave(torch.arange(n_rows), rowptr[1:] - rowptr[:-1])
            perm = torch.argsort(col * n_rows + row, stable=True)
            ct = torch.zeros(n_cols + 1, dtype=torch.long)
            torch.cumsum(torch.bincount(col, minlength=n_cols), 0, out=ct[1:])
            At = torch.sparse_csr_tensor(ct, row[perm], v[perm], size=(n_cols, n_rows))
        if len(_CSR_CACHE) > 16:
            _CSR_CACHE.clear()
        hit = _CSR_CACHE[key] = (A, At, rowptr, col, v)     # keep the keyed tensors alive
    return hit[0], hit[1]


def spmm_csr(rowptr, col, val: Optional[torch.Tensor], x, n_rows: int, reduce: str = "sum"):
    """Form (ii): torch.sparse_csr_tensor @ x (MKL); mean divides by max(rowcount,1)."""
    v = torch.ones(col.numel(), dtype=x.dtype) if val is None else val.to(x.dtype)
    if val is not None and v.data_ptr() != val.data_ptr():
        A = torch.sparse_csr_tensor(rowptr, col, v, size=(n_rows, x.shape[0]))
        out = A @ x
    elif val is None:
        A = torch.sparse_csr_tensor(rowptr, col, v, size=(n_rows, x.shape[0]))
        out = A @ x
    else:
        A, At = _csr_pair(rowptr, col, v, n_rows, x.shape[0])
        out = _CsrMatmul.apply(x, A, At)
    if reduce == "mean":
        cnt = (rowptr[1:] - rowptr[:-1]).clamp(min=1).to(x.dtype)
        out = out / cnt.view(-1, 1)
    return out


def spmm_dense(row, col, val: Optional[torch.Tensor], x, n_rows: int, reduce: str = "sum"):
    """Form (iii): dense A @ x, small graphs only."""
    A = torch.zeros(n_rows, x.shape[0], dtype=x.dtype)
    v = torch.ones(col.numel(), dtype=x.dtype) if val is None else val.to(x.dtype)
    A.index_put_((row, col), v, accumulate=True)
    out = A @ x
    if reduce == "mean":
        cnt = torch.zeros(n_rows, dtype=x.dtype).scatter_add_(0, row, torch.ones_like(row, dtype=x.dtype))
        out = out / cnt.clamp(min=1).view(-1, 1)
    return out


def segment_softmax(src: torch.Tensor, index: torch.Tensor, num_nodes: Optional[int] = None) -> torch.Tensor:
    """torch_geometric.utils.softmax(src, index) over an UNSORTED index (arxiv_pyg/criterion.py:103-113):
    N = index.max()+1; out = exp(src - max_seg) / (sum_seg + 1e-16)."""
    N = int(index.max()) + 1 if num_nodes is None else num_nodes
    shape = (N,) + tuple(src.shape[1:])
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    m = torch.full(shape, float("-inf"), dtype=src.dtype).scatter_reduce_(0, idx, src.detach(), reduce="amax", include_self=True)
    e = (src - m.index_select(0, index)).exp()
    s = torch.zeros(shape, dtype=src.dtype).scatter_add_(0, idx, e)
    return e / (s.index_select(0, index) + 1e-16)
