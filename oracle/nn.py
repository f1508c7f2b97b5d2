"""Layer / model forwards restated functionally (weights passed in), CPU PyTorch.

GCNConv / SAGEConv semantics follow SURVEY.md Appendix A.2 / A.3 as used by the reference's
``GCN`` / ``SAGE`` modules (arxiv_pyg/gnn.py:23-85); BatchNorm1d / dropout per A.8.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn.functional as F

from . import ops


def gcn_conv(x, rowptr, col, val, weight, bias, form: str = "csr"):
    """out = Â (x W) + b  — transform first, then aggregate (PyG GCNConv, SparseTensor path). weight: [in,out]."""
    h = x @ weight
    n = rowptr.numel() - 1
    if form == "csr":
        out = ops.spmm_csr(rowptr, col, val, h, n, "sum")
    else:
        row = torch.repeat_interleave(torch.arange(n), rowptr[1:] - rowptr[:-1])
        out = ops.spmm_scatter(row, col, val, h, n, "sum")
    return out if bias is None else out + bias


def sage_conv(x, rowptr, col, w_l, b_l, w_r, form: str = "csr"):
    """out = lin_l(mean_j x_j) + lin_r(x) (PyG SAGEConv; w_l/w_r are nn.Linear weights [out,in])."""
    n = rowptr.numel() - 1
    if form == "csr":
        agg = ops.spmm_csr(rowptr, col, None, x, n, "mean")
    else:
        row = torch.repeat_interleave(torch.arange(n), rowptr[1:] - rowptr[:-1])
        agg = ops.spmm_scatter(row, col, None, x, n, "mean")
    return F.linear(agg, w_l, b_l) + F.linear(x, w_r)


def batch_norm_train(x, gamma, beta, eps: float = 1e-5):
    """BatchNorm1d in training mode: biased batch variance for the normalisation."""
    mean = x.mean(0)
    var = x.var(0, unbiased=False)
    return (x - mean) / torch.sqrt(var + eps) * gamma + beta


def gcn_forward(x, rowptr, col, val, weights: List[torch.Tensor], biases: List[torch.Tensor],
                bn_gamma: List[torch.Tensor], bn_beta: List[torch.Tensor],
                dropout_masks: Optional[List[torch.Tensor]] = None, p: float = 0.5, form: str = "csr"):
    """GCN.forward (arxiv_pyg/gnn.py:45-53) in training mode. ``dropout_masks[i]`` (bool keep-mask) replaces the
    RNG so CUDA and oracle share the mask; None => no dropout. Returns (logits, last hidden = model.out_feat)."""
    hidden = None
    for i in range(len(weights) - 1):
        x = gcn_conv(x, rowptr, col, val, weights[i], biases[i], form)
        x = batch_norm_train(x, bn_gamma[i], bn_beta[i])
        x = torch.relu(x)
        if dropout_masks is not None:
            x = x * dropout_masks[i].to(x.dtype) / (1.0 - p)
        hidden = x
    return gcn_conv(x, rowptr, col, val, weights[-1], biases[-1], form), hidden


def sage_forward(x, rowptr, col, params: List[dict], bn_gamma, bn_beta, dropout_masks=None, p: float = 0.5):
    """SAGE.forward (arxiv_pyg/gnn.py:77-85); params[i] = dict(w_l, b_l, w_r)."""
    hidden = None
    for i in range(len(params) - 1):
        q = params[i]
        x = sage_conv(x, rowptr, col, q["w_l"], q["b_l"], q["w_r"])
        x = batch_norm_train(x, bn_gamma[i], bn_beta[i])
        x = torch.relu(x)
        if dropout_masks is not None:
            x = x * dropout_masks[i].to(x.dtype) / (1.0 - p)
        hidden = x
    q = params[-1]
    return sage_conv(x, rowptr, col, q["w_l"], q["b_l"], q["w_r"]), hidden


def gat_aggregate(ft, el, er, row, col, n_dst: int, heads: int, negative_slope: float = 0.2, softmax_eps: float = 0.0):
    """Edge attention + weighted aggregation restated with plain torch (DGL: apply_edges(u_add_v) -> leaky_relu ->
    edge_softmax -> update_all(u_mul_e, sum), arxiv_dgl/models.py:202-217; PyG GATConv with softmax_eps=1e-16).
    row = destination, col = source of every edge; ft [N, H*D], el [N, H], er [n_dst, H] or None."""
    H = heads
    D = ft.shape[1] // H
    e = el.index_select(0, col) + (er.index_select(0, row) if er is not None else 0)
    e = torch.where(e > 0, e, e * negative_slope)
    idx = row.view(-1, 1).expand_as(e)
    m = torch.full((n_dst, H), float("-inf"), dtype=e.dtype).scatter_reduce_(0, idx, e.detach(), "amax", include_self=True)
    ex = (e - m.index_select(0, row)).exp()
    s = torch.zeros(n_dst, H, dtype=e.dtype).scatter_add_(0, idx, ex)
    a = ex / (s.index_select(0, row) + softmax_eps)
    msg = ft.index_select(0, col).view(-1, H, D) * a.unsqueeze(-1)
    out = torch.zeros(n_dst, H, D, dtype=ft.dtype).index_add_(0, row, msg)
    return out.reshape(n_dst, H * D)


def dgl_graph_conv_both(x, row, col, n: int, weight, bias=None):
    """DGL GraphConv(norm='both') as called at arxiv_dgl/models.py:65,80 (DGL 0.5/0.6 semantics restated: scale sources by
    out_degree^-1/2 (degrees clamped to >= 1), multiply by W first iff in > out, sum over in-edges, scale by in_degree^-1/2,
    add bias).  row = destination, col = source of each edge."""
    d_out = torch.bincount(col, minlength=n).clamp(min=1).to(x.dtype).pow(-0.5)
    d_in = torch.bincount(row, minlength=n).clamp(min=1).to(x.dtype).pow(-0.5)
    h = x * d_out[:, None]
    if weight.shape[0] > weight.shape[1]:
        h = h @ weight
        rst = ops.scatter(h[col], row, n, "sum")
    else:
        rst = ops.scatter(h[col], row, n, "sum") @ weight
    rst = rst * d_in[:, None]
    return rst if bias is None else rst + bias


def neighbor_average_features(x, row, col, n: int, R: int):
    """arxiv_dgl/sign.py:175-183: R rounds of update_all(copy_u, mean) — mean over in-edges, zero where there are none."""
    res = [x]
    for _ in range(R):
        res.append(ops.scatter(res[-1][col], row, n, "mean"))
    return res


def projection_gcd(x, rowptr, col, val, lin_w, lin_b, conv_w, conv_b, gamma, beta):
    """ProjectionGCD.forward (arxiv_pyg/gnn.py:95-99): relu(BN(Linear(x) + GCNConv(x, adj_t))); rowptr/col/val is the
    gcn-normalised adjacency (the reference's conv is non-cached, i.e. it re-derives the same normalisation every call)."""
    h = F.linear(x, lin_w, lin_b) + gcn_conv(x, rowptr, col, val, conv_w, conv_b)
    return torch.relu(batch_norm_train(h, gamma, beta))


def dgl_gat_conv(x, row, col, n: int, fc_w, attn_l, attn_r, res_w, heads: int, negative_slope: float = 0.2,
                 symmetric_norm: bool = True):
    """The reference's DGL GATConv.forward (arxiv_dgl/models.py:154-236) without the stochastic teacher-training dropouts:
    ft = fc(x) [* out_deg^-1/2]; el = <ft, attn_l>, er = <fc(x), attn_r> (destination side NOT rescaled); a = edge_softmax(leaky_relu(el[src] + er[dst]));
    rst = sum_e a ft[src] [* in_deg^+1/2 — the exponent the reference uses, :218-223] (+ res_fc(x)).  Returns [n, H, D]."""
    H = heads
    D = fc_w.shape[0] // H
    ft = F.linear(x, fc_w).view(-1, H, D)
    ft_dst = ft                              # :187-188: feat_dst is bound before feat_src is rescaled, so er sees the raw projection
    if symmetric_norm:
        ft = ft * torch.bincount(col, minlength=n).clamp(min=1).to(x.dtype).pow(-0.5).view(-1, 1, 1)
    el = (ft * attn_l).sum(-1)
    er = (ft_dst * attn_r).sum(-1) if attn_r is not None else None
    rst = gat_aggregate(ft.reshape(-1, H * D), el, er, row, col, n, H, negative_slope, 0.0).view(-1, H, D)
    if symmetric_norm:
        rst = rst * torch.bincount(row, minlength=n).clamp(min=1).to(x.dtype).pow(0.5).view(-1, 1, 1)
    if res_w is not None:
        rst = rst + F.linear(x, res_w).view(x.shape[0], -1, D)
    return rst


def rgcn_group_input(x_dict, emb_dict, node_type, local_node_idx, in_channels: int):
    """RGCN.group_input (mag_pyg/gnn.py:111-124): rows of the homogeneous feature matrix come from the typed feature
    matrices (x_dict, keyed by node-type int) or the learned embedding tables (emb_dict, keyed by str(node-type int))."""
    h = torch.zeros(node_type.numel(), in_channels, dtype=next(iter(x_dict.values())).dtype)
    for key, x in list(x_dict.items()) + [(int(k), e) for k, e in emb_dict.items()]:
        idx = (node_type == key).nonzero().view(-1)
        h = h.index_add(0, idx, x[local_node_idx[idx]])
    return h


def rgcn_conv(x, edge_index, edge_type, node_type, rel_w, root_w, root_b):
    """RGCNConv.forward / .message (mag_pyg/gnn.py:54-68): for every relation, mean over incoming edges of that relation of
    rel_lins[r](x_j) (transform per edge, then scatter-mean over ALL nodes), plus the per-node-type root Linear."""
    n = x.shape[0]
    out = torch.zeros(n, rel_w[0].shape[0], dtype=x.dtype)
    for r, W in enumerate(rel_w):
        ei = edge_index[:, edge_type == r]
        out = out + ops.scatter(F.linear(x[ei[0]], W), ei[1], n, "mean")
    for t, (W, b) in enumerate(zip(root_w, root_b)):
        idx = (node_type == t).nonzero().view(-1)
        out = out.index_add(0, idx, F.linear(x[idx], W, b))
    return out


def rgcn_inference_layer(x_dict, edge_index_dict, key2int, rel_w, root_w, root_b, relu: bool):
    """One layer of RGCN.inference (mag_pyg/gnn.py:153-169): out[t] = root_lins[t](x_t) + sum over relations into t of
    rel_lins[r](mean over sources of x_src) — aggregate first (SparseTensor.matmul(reduce='mean')), then transform."""
    out = {t: F.linear(x, root_w[t], root_b[t]) for t, x in x_dict.items()}
    for keys, ei in edge_index_dict.items():
        s, t = key2int[keys[0]], key2int[keys[-1]]
        agg = ops.scatter(x_dict[s][ei[0]], ei[1], x_dict[t].shape[0], "mean")
        out[t] = out[t] + F.linear(agg, rel_w[key2int[keys]])
    return {t: torch.relu(v) if relu else v for t, v in out.items()}
