"""The PyG-surface mirrors (GCNConv / SAGEConv / MessagePassing / scatter) on the GPU: the module-level path the
reference's unmodified scripts take.  Checked against the fixtures produced by the reference's GCN / SAGE classes."""
import sys
from pathlib import Path

import pytest
import torch
import torch.nn.functional as F

import efficient_gnns_b200  # noqa: F401
from conftest import rel_err
from efficient_gnns_b200 import nn as bnn
from efficient_gnns_b200.sparse import SparseTensor
from oracle import ops as oo

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


class Student(torch.nn.Module):
    """Same module tree as the reference's GCN / SAGE students: convs.{i}, bns.{i}; conv -> BN -> ReLU -> dropout."""

    def __init__(self, conv_cls, dims, dropout=0.0, **kw):
        super().__init__()
        self.convs = torch.nn.ModuleList([conv_cls(dims[i], dims[i + 1], **kw) for i in range(len(dims) - 1)])
        self.bns = torch.nn.ModuleList([torch.nn.BatchNorm1d(d) for d in dims[1:-1]])
        self.dropout = dropout

    def forward(self, x, adj_t):
        for conv, bn in zip(self.convs[:-1], self.bns):
            x = F.dropout(F.relu(bn(conv(x, adj_t))), p=self.dropout, training=self.training)
            self.out_feat = x
        return self.convs[-1](x, adj_t)


@pytest.mark.parametrize("name,cls,kw", [("gcn", bnn.GCNConv, dict(cached=True)), ("sage", bnn.SAGEConv, {})])
def test_module_path_reproduces_reference_fixture(golden_model, name, cls, kw):
    G, m = golden_model, golden_model["models"][name]
    n = G["x"].shape[0]
    adj = SparseTensor(row=G["sym_row"].cuda(), col=G["sym_col"].cuda(), sparse_sizes=(n, n), is_sorted=True)
    model = Student(cls, [16, 32, 32, 8], **kw).cuda()
    model.load_state_dict({k: v for k, v in m["state"].items()})
    model.train()
    out = model(G["x"].cuda(), adj)
    assert rel_err(out, m["logits_train"]) < 1e-5
    assert rel_err(model.out_feat, m["out_feat"]) < 1e-5
    idx = G["train_idx"].cuda()
    loss = F.cross_entropy(out[idx], m["y"].cuda()[idx])
    assert abs(loss.item() - m["loss"].item()) < 1e-5 * m["loss"].item()
    loss.backward()
    for k, p in model.named_parameters():
        ref = m["grads"][k]
        if k.endswith("bias") and k.startswith("convs") and not k.startswith("convs.2") and name == "gcn":
            continue                                   # zero-gradient bias in front of BatchNorm
        if name == "sage" and k.endswith("lin_l.bias") and not k.startswith("convs.2"):
            continue
        assert rel_err(p.grad, ref) < 5e-5, k
    model.load_state_dict({k: v for k, v in m["state"].items()})   # the train-mode pass above advanced the running stats
    model.eval()
    assert rel_err(model(G["x"].cuda(), adj), m["logits_eval"]) < 1e-5


class ProjectionGCD(torch.nn.Module):
    """Same module tree as the reference's ProjectionGCD (arxiv_pyg/gnn.py:88-99) on the mirrored layers."""

    def __init__(self, hidden, proj):
        super().__init__()
        self.lin = bnn.Linear(hidden, proj)
        self.conv = bnn.GCNConv(hidden, proj)          # non-cached: normalises the adjacency on every call
        self.bn = torch.nn.BatchNorm1d(proj)

    def forward(self, x, adj_t):
        return F.relu(self.bn(self.lin(x) + self.conv(x, adj_t)))


def test_projection_gcd_reproduces_reference_fixture(golden_model):
    G, m = golden_model, golden_model["models"]["proj_gcd"]
    n = G["x"].shape[0]
    adj = SparseTensor(row=G["sym_row"].cuda(), col=G["sym_col"].cuda(), sparse_sizes=(n, n), is_sorted=True)
    head = ProjectionGCD(16, 12).cuda()
    head.load_state_dict(m["state"])
    head.train()
    x = G["x"].cuda().requires_grad_(True)
    out = head(x, adj)
    assert rel_err(out, m["out_train"]) < 1e-5
    (out * m["w"].cuda()).sum().backward()
    assert rel_err(x.grad, m["d_x"]) < 5e-5
    for k, p in head.named_parameters():
        if k in ("lin.bias", "conv.bias"):
            continue                                   # zero-gradient biases in front of BatchNorm
        assert rel_err(p.grad, m["grads"][k]) < 5e-5, k


def test_message_passing_mean_matches_oracle_scatter():
    class Rel(bnn.MessagePassing):
        def __init__(self):
            super().__init__(aggr="mean")
            self.lin = bnn.Linear(24, 16, bias=False)

        def forward(self, x, edge_index):
            return self.propagate(edge_index, x=x, edge_type=3)

        def message(self, x_j, edge_type: int):
            assert edge_type == 3
            return self.lin(x_j)

    g = torch.Generator().manual_seed(0)
    n, E = 500, 4000
    ei = torch.randint(0, n, (2, E), generator=g)
    x = torch.randn(n, 24, generator=g)
    layer = Rel().cuda()
    xc = x.cuda().requires_grad_(True)
    out = layer(xc, ei.cuda())
    w = layer.lin.weight.detach().cpu().double()
    xr = x.double().requires_grad_(True)
    ref = oo.scatter(xr.index_select(0, ei[0]) @ w.t(), ei[1], n, "mean")
    assert rel_err(out, ref) < 1e-5
    (out.sum() * 1.0).backward(); ref.sum().backward()
    assert rel_err(xc.grad, xr.grad) < 2e-5


def test_scatter_sum_and_linear_fallbacks():
    g = torch.Generator().manual_seed(1)
    src, idx = torch.randn(3000, 12, generator=g), torch.randint(0, 200, (3000,), generator=g)
    out = bnn.scatter(src.cuda(), idx.cuda(), 0, 250, "sum")
    assert out.shape == (250, 12) and rel_err(out, oo.scatter(src.double(), idx, 250, "sum")) < 1e-5
    x, w = torch.randn(100, 10, generator=g), torch.randn(7, 10, generator=g)      # widths not multiples of 4 -> library GEMM
    assert rel_err(bnn.linear(x.cuda(), w.cuda()), x @ w.t()) < 1e-5
