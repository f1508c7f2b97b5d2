"""Operators: thin, checked launches of the C ABI plus their autograd wiring.

Every function here ends in a ``libb200gnn.so`` call on the current CUDA stream;
PyTorch only provides the device buffers and the autograd tape.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import lib
from .sparse import CsrGraph, SparseTensor

_REDUCE = {"sum": lib.REDUCE_SUM, "add": lib.REDUCE_SUM, "mean": lib.REDUCE_MEAN}


def stat_slots(g: CsrGraph) -> int:
    return int(lib.load().b200gnn_spmm_stat_slots(g.n_rows, g.n_hub))


def spmm_csr(g: CsrGraph, x: torch.Tensor, reduce: str = "sum", bias: Optional[torch.Tensor] = None,
             out: Optional[torch.Tensor] = None, stat_partial: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Y = reduce_e val[e]·X[col[e]] (+bias), optional fused column statistics. No autograd."""
    if reduce not in _REDUCE:
        raise ValueError(f"reduce={reduce!r}: the engine implements sum/add/mean (what the reference path uses)")
    if x.dim() != 2:
        raise lib.B200GnnError("spmm: dense operand must be [n_src, K]")
    if x.shape[0] != g.n_cols:
        raise lib.B200GnnError(f"spmm: dense operand has {x.shape[0]} rows, matrix has {g.n_cols} columns")
    K = x.shape[1]
    if out is None:
        out = torch.empty(g.n_rows, K, dtype=torch.float32, device=x.device)
    if g.n_rows == 0 or K == 0:
        return out
    L = lib.load()
    ws = g.hub_workspace(K)
    rc = L.b200gnn_spmm_csr_f32(
        lib.dptr(g.rowptr, torch.int32, "rowptr"), lib.dptr(g.col, torch.int32, "col"),
        lib.dptr(g.val, torch.float32, "val"), lib.dptr(x, torch.float32, "x"), x.stride(0),
        lib.dptr(out, torch.float32, "out"), out.stride(0), g.n_rows, g.n_cols, K, _REDUCE[reduce],
        lib.dptr(bias, torch.float32, "bias"), lib.dptr(stat_partial, torch.float32, "stat_partial"),
        g.hub_threshold, g.seg_len,
        g.hub_rows.data_ptr() if g.n_hub else None, g.hub_segptr.data_ptr() if g.n_hub else None,
        g.n_hub, g.n_seg, None if ws is None else ws.data_ptr(), lib.stream_ptr())
    lib.check(rc, "spmm_csr_f32")
    return out


class _SpMM(torch.autograd.Function):
    """matmul(adj, x, reduce): backward is the same kernel on the cached CSC view
    (upstream torch_sparse spmm backward)."""

    @staticmethod
    def forward(ctx, x, adj: SparseTensor, reduce: str):
        st = adj.storage
        g = st.engine_csr() if st.value() is not None else st.engine_csr_unweighted()
        ctx.adj, ctx.reduce = adj, reduce
        return spmm_csr(g, x.contiguous(), reduce)

    @staticmethod
    def backward(ctx, grad_out):
        st = ctx.adj.storage
        gt = st.engine_csc("mean" if ctx.reduce == "mean" else "value")
        return spmm_csr(gt, grad_out.contiguous(), "sum"), None, None


def matmul(adj: SparseTensor, x: torch.Tensor, reduce: str = "sum") -> torch.Tensor:
    """torch_sparse.matmul(adj, dense, reduce) for reduce in {sum, add, mean}."""
    if reduce not in _REDUCE:
        raise ValueError(f"reduce={reduce!r} not implemented (reference path uses add/mean)")
    v = adj.storage.value()
    if v is not None and v.requires_grad:
        raise NotImplementedError("gradients w.r.t. sparse values are never taken on the reference path")
    if x.dim() == 1:
        return matmul(adj, x.unsqueeze(-1), reduce).squeeze(-1)
    return _SpMM.apply(x, adj, reduce)
