"""Graph-attention kernels (edge softmax + multi-head aggregation, forward and backward) vs the fp64 oracle; the DGL-style
and PyG-style modules on top of them."""
import numpy as np
import pytest
import torch

import efficient_gnns_b200  # noqa: F401
from conftest import rel_err
from efficient_gnns_b200 import nn as bnn
from efficient_gnns_b200.sparse import SparseTensor
from efficient_gnns_b200.synthetic import skewed_edges
from oracle import graph as og, nn as onn

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def graph(n, e, seed, self_loops=True):
    ei = skewed_edges(n, e, seed).numpy()
    row, col, _ = og.to_sparse_adj_t(ei, n)
    r, c = og.to_symmetric(row, col, n)
    if self_loops:
        r, c, _ = og.fill_diag(r, c, np.ones(r.shape[0], dtype=np.float32), n)
    return torch.from_numpy(r), torch.from_numpy(c)


@pytest.mark.parametrize("H,D,with_er,eps", [(8, 32, True, 0.0), (3, 250, False, 0.0), (4, 16, True, 1e-16), (1, 7, True, 0.0),
                                             (6, 10, True, 0.0), (4, 100, True, 0.0), (4, 200, False, 0.0),
                                             (16, 80, True, 0.0), (3, 101, True, 0.0), (3, 40, True, 0.0)])
def test_gat_aggregate_forward_backward(H, D, with_er, eps):
    n = 4000
    r, c = graph(n, 30_000, H + D)          # skewed: contains hub rows above the split threshold
    g = torch.Generator().manual_seed(D)
    ft = torch.randn(n, H * D, generator=g)
    el, er = torch.randn(n, H, generator=g), (torch.randn(n, H, generator=g) if with_er else None)
    w = torch.randn(n, H * D, generator=g)
    ftr, elr = ft.double().requires_grad_(True), el.double().requires_grad_(True)
    err = er.double().requires_grad_(True) if with_er else None
    ref = onn.gat_aggregate(ftr, elr, err, r, c, n, H, 0.2, eps)
    (ref * w.double()).sum().backward()
    adj = SparseTensor(row=r.cuda(), col=c.cuda(), sparse_sizes=(n, n), is_sorted=True)
    assert adj.storage.engine_csr_unweighted().n_hub > 0
    ftc, elc = ft.cuda().requires_grad_(True), el.cuda().requires_grad_(True)
    erc = er.cuda().requires_grad_(True) if with_er else None
    out = bnn.gat_aggregate(ftc, elc, erc, adj, H, 0.2, eps)
    (out * w.cuda()).sum().backward()
    assert rel_err(out, ref) < 1e-5
    assert rel_err(ftc.grad, ftr.grad) < 2e-5
    assert rel_err(elc.grad, elr.grad) < 2e-5
    if with_er:
        assert rel_err(erc.grad, err.grad) < 2e-5
    out2 = bnn.gat_aggregate(ftc, elc, erc, adj, H, 0.2, eps)
    assert torch.equal(out, out2)


@pytest.mark.parametrize("name", ["attn_dst", "no_attn_dst"])
def test_dgl_gatconv_reproduces_reference_class_fixture(golden_gat, name):
    """nn.DGLGATConv (CUDA kernels) vs the fixture produced by the reference's own GATConv class."""
    G, m = golden_gat, golden_gat["layers"][name]
    n = G["x"].shape[0]
    adj = SparseTensor(row=G["row"].cuda(), col=G["col"].cuda(), sparse_sizes=(n, n), is_sorted=True)
    layer = bnn.DGLGATConv(16, 8, num_heads=3, residual=True, use_symmetric_norm=True,
                           use_attn_dst=(name == "attn_dst")).cuda()
    layer.load_state_dict(m["state"])
    x = G["x"].cuda().requires_grad_(True)
    out = layer(adj, x)
    assert rel_err(out, m["out"]) < 1e-5
    (out * m["w"].cuda()).sum().backward()
    assert rel_err(x.grad, m["d_x"]) < 5e-5
    for k, p in layer.named_parameters():
        assert rel_err(p.grad, m["grads"][k]) < 5e-5, k


def test_dgl_style_gatconv_layer_matches_oracle_composition():
    n, Fin, H, D = 1500, 64, 3, 20
    r, c = graph(n, 9000, 1)
    adj = SparseTensor(row=r.cuda(), col=c.cuda(), sparse_sizes=(n, n), is_sorted=True)
    torch.manual_seed(0)
    layer = bnn.DGLGATConv(Fin, D, num_heads=H, residual=True, use_symmetric_norm=True, activation=None).cuda()
    x = torch.randn(n, Fin)
    out = layer(adj, x.cuda())
    # the layer restated in fp64 (oracle.nn.dgl_gat_conv, itself pinned on the reference class's fixture)
    W, al, ar, Wr = (p.detach().cpu().double() for p in (layer.fc.weight, layer.attn_l, layer.attn_r, layer.res_fc.weight))
    rst = onn.dgl_gat_conv(x.double(), r, c, n, W, al, ar, Wr, H, 0.2, True)
    assert rel_err(out, rst) < 1e-5


def test_pyg_style_gatconv_adds_self_loops_and_concats():
    n, Fin, H, C = 800, 48, 4, 12
    ei = skewed_edges(n, 5000, 2)
    torch.manual_seed(1)
    layer = bnn.GATConv(Fin, C, heads=H).cuda()
    x = torch.randn(n, Fin)
    out = layer(x.cuda(), ei.cuda())
    assert out.shape == (n, H * C)
    W, al, ar, b = (p.detach().cpu().double() for p in (layer.lin_l.weight, layer.att_l, layer.att_r, layer.bias))
    xl = (x.double() @ W.t()).view(n, H, C)
    src, dst = ei[0], ei[1]
    loops = torch.arange(n)
    row, col = torch.cat([dst, loops]), torch.cat([src, loops])
    ref = onn.gat_aggregate(xl.reshape(n, H * C), (xl * al).sum(-1), (xl * ar).sum(-1), row, col, n, H, 0.2, 1e-16) + b
    assert rel_err(out, ref) < 1e-5


def _dense_attention_ref(ft, el, er, row, col, n, H, D, slope, keep, scale):
    """fp64 restatement of arxiv_dgl/models.py:202-217 with given edge keep-mask and attention-dropout scale (per edge)."""
    e = torch.nn.functional.leaky_relu(el[col] + (er[row] if er is not None else 0), slope)        # [nnz, H]
    if keep is not None:
        e = e.masked_fill(~keep.bool().view(-1, 1), float("-inf"))
    m = torch.full((n, H), float("-inf"), dtype=e.dtype).scatter_reduce_(0, row.view(-1, 1).expand_as(e), e.detach(), "amax")
    m = torch.where(torch.isinf(m), torch.zeros_like(m), m)
    ex = (e - m[row]).exp()
    ssum = torch.zeros(n, H, dtype=e.dtype).index_add_(0, row, ex)
    a = ex / ssum[row].clamp_min(1e-300)
    if scale is not None:
        a = a * scale
    msg = ft.view(-1, H, D)[col] * a.unsqueeze(-1)
    return torch.zeros(n, H, D, dtype=e.dtype).index_add_(0, row, msg).reshape(n, H * D)


@pytest.mark.parametrize("H,D", [(3, 40), (8, 32)])
def test_gat_edge_drop_and_attention_dropout_forward_backward(H, D):
    """Teacher-training knobs of the reference's GATConv (arxiv_dgl/models.py:206-214) inside the fused kernels: a given edge
    keep-mask (dropped edges leave the softmax; a destination may lose all its edges) and a given coefficient dropout."""
    n = 3000
    r, c = graph(n, 25_000, 3)
    nnz = r.numel()
    g = torch.Generator().manual_seed(H)
    ft, el, er = torch.randn(n, H * D, generator=g), torch.randn(n, H, generator=g), torch.randn(n, H, generator=g)
    w = torch.randn(n, H * D, generator=g)
    keep = (torch.rand(nnz, generator=g) >= 0.3).to(torch.uint8)
    keep[r == 11] = 0                                              # one destination loses every edge
    scale = (torch.rand(nnz, H, generator=g) >= 0.25).float() / 0.75
    ftr, elr, err = (t.double().requires_grad_(True) for t in (ft, el, er))
    ref = _dense_attention_ref(ftr, elr, err, r, c, n, H, D, 0.2, keep, scale.double())
    (ref * w.double()).sum().backward()
    adj = SparseTensor(row=r.cuda(), col=c.cuda(), sparse_sizes=(n, n), is_sorted=True)
    ftc, elc, erc = (t.cuda().requires_grad_(True) for t in (ft, el, er))
    out = bnn.gat_aggregate(ftc, elc, erc, adj, H, 0.2, 0.0, keep.cuda(), scale.cuda())
    (out * w.cuda()).sum().backward()
    assert torch.count_nonzero(out[11]) == 0
    assert rel_err(out, ref) < 1e-5
    assert rel_err(ftc.grad, ftr.grad) < 2e-5 and rel_err(elc.grad, elr.grad) < 2e-5 and rel_err(erc.grad, err.grad) < 2e-5


def test_dgl_gatconv_training_knobs_run_and_are_inert_in_eval():
    n = 1200
    r, c = graph(n, 8000, 4)
    adj = SparseTensor(row=r.cuda(), col=c.cuda(), sparse_sizes=(n, n), is_sorted=True)
    torch.manual_seed(0)
    layer = bnn.DGLGATConv(32, 16, num_heads=3, attn_drop=0.1, edge_drop=0.2, residual=True, use_symmetric_norm=True).cuda()
    plain = bnn.DGLGATConv(32, 16, num_heads=3, residual=True, use_symmetric_norm=True).cuda()
    plain.load_state_dict(layer.state_dict())
    x = torch.randn(n, 32).cuda().requires_grad_(True)
    layer.train()
    out = layer(adj, x)
    out.sum().backward()
    assert torch.isfinite(out).all() and torch.isfinite(x.grad).all()
    layer.eval()
    assert torch.equal(layer(adj, x), plain(adj, x))


@pytest.mark.parametrize("concat", [True, False])
@pytest.mark.parametrize("H,C", [(4, 12), (4, 121), (6, 121)])      # 121: PPI's class count — a head width that is not a multiple of 4
def test_pyg_gatconv_gradients_and_head_mean(concat, H, C):
    """PyG GATConv as ppi_pyg/gnn.py:27-31 uses it (heads 4/6, last layer concat=False): forward and every gradient
    against the fp64 restatement (self-loops re-added, softmax eps 1e-16)."""
    n, Fin = 700, 50
    ei = skewed_edges(n, 4000, 5)
    torch.manual_seed(2)
    layer = bnn.GATConv(Fin, C, heads=H, concat=concat).cuda()
    x = torch.randn(n, Fin)
    w = torch.randn(n, H * C if concat else C)
    xc = x.cuda().requires_grad_(True)
    out = layer(xc, ei.cuda())
    (out * w.cuda()).sum().backward()
    W, al, ar, b = (p.detach().cpu().double().requires_grad_(True) for p in (layer.lin_l.weight, layer.att_l, layer.att_r, layer.bias))
    xr = x.double().requires_grad_(True)
    xl = (xr @ W.t()).view(n, H, C)
    loops = torch.arange(n)
    row, col = torch.cat([ei[1], loops]), torch.cat([ei[0], loops])
    agg = onn.gat_aggregate(xl.reshape(n, H * C), (xl * al).sum(-1), (xl * ar).sum(-1), row, col, n, H, 0.2, 1e-16).view(n, H, C)
    ref = (agg.reshape(n, H * C) if concat else agg.mean(1)) + b
    (ref * w.double()).sum().backward()
    assert rel_err(out, ref) < 1e-5
    assert rel_err(xc.grad, xr.grad) < 5e-5
    for p_, r_ in ((layer.lin_l.weight, W), (layer.att_l, al), (layer.att_r, ar), (layer.bias, b)):
        assert rel_err(p_.grad, r_.grad) < 5e-5
