"""Round-2 additions against the CPU oracle: PPI (binary cross-entropy) criteria, KD rows with 349 classes (ogbn-mag),
GINConv, mean-SpMM of a WEIGHTED matrix (gradient), PyG GATConv gradients / head-mean, empty relations."""
import pytest
import torch

import efficient_gnns_b200  # noqa: F401
from conftest import rel_err
from efficient_gnns_b200 import criterion as bc, criterion_ppi as bp, nn as bnn, ops
from efficient_gnns_b200.sparse import SparseTensor
from efficient_gnns_b200.synthetic import skewed_edges
from oracle import criterion as oc, graph as og, nn as onn, ops as oo

pytestmark = pytest.mark.gpu


def test_kd_rows_349_classes_like_ogbn_mag():
    g = torch.Generator().manual_seed(0)
    n, C = 3000, 349
    z = torch.randn(n, C, generator=g).requires_grad_(True)
    t = torch.randn(n, C, generator=g) * 2
    y = torch.randint(0, C, (n,), generator=g)
    ref = oc.kd_criterion(z.double(), y, t.double(), 0.9, 4.0)
    ref[0].backward()
    zc = z.detach().cuda().requires_grad_(True)
    got = bc.kd_criterion(zc, y.cuda(), t.cuda(), 0.9, 4.0)
    got[0].backward()
    for a, b in zip(got, ref):
        assert abs(float(a) - float(b)) <= 1e-5 * abs(float(b))
    assert rel_err(zc.grad, z.grad) < 1e-5
    # every other criterion routes its classification term through the same kernel
    assert abs(float(bc.cross_entropy(zc, y.cuda())) - float(oc.cross_entropy(z.double(), y))) < 1e-5 * 6


@pytest.mark.parametrize("n,C", [(2400, 121), (7, 3)])
def test_ppi_kd_criterion_bce(n, C):
    g = torch.Generator().manual_seed(1)
    z = torch.randn(n, C, generator=g).requires_grad_(True)
    t = torch.randn(n, C, generator=g) * 3
    y = (torch.rand(n, C, generator=g) < 0.3).float()
    ref = oc.kd_criterion_ppi(z.double(), y.double(), t.double())
    ref[0].backward()
    zc = z.detach().cuda().requires_grad_(True)
    got = bp.kd_criterion(zc, y.cuda(), t.cuda())
    got[0].backward()
    for a, b in zip(got, ref):
        assert abs(float(a) - float(b)) <= 1e-5 * abs(float(b))
    assert rel_err(zc.grad, z.grad) < 1e-5


def test_ppi_aux_criteria_use_bce_classification_term():
    g = torch.Generator().manual_seed(2)
    n, C, F = 500, 121, 64
    z, fs, ft = torch.randn(n, C, generator=g), torch.randn(n, F, generator=g), torch.randn(n, F, generator=g)
    y = (torch.rand(n, C, generator=g) < 0.3).float()
    got = bp.fitnet_criterion(z.cuda(), y.cuda(), fs.cuda(), ft.cuda())
    cls = oc.bce_with_logits(z.double(), y.double())
    aux = oc.fitnet_criterion(z.double(), torch.zeros(n, dtype=torch.long), fs.double(), ft.double())[2]
    assert abs(float(got[1]) - float(cls)) < 1e-5 * float(cls)
    assert abs(float(got[0]) - float(cls + 1000 * aux)) < 1e-5 * float(cls + 1000 * aux)
    ei = skewed_edges(n, 3000, 3)
    got = bp.lpw_criterion(z.cuda(), y.cuda(), fs.cuda(), ft.cuda(), ei.cuda())
    aux = oc.lpw_criterion(z.double(), torch.zeros(n, dtype=torch.long), fs.double(), ft.double(), ei)[2]
    assert abs(float(got[2]) - float(aux)) < 1e-5 * abs(float(aux))
    assert abs(float(got[0]) - float(cls + 100 * aux)) < 1e-5 * float(cls + 100 * aux)


def _sym_graph(n, e, seed):
    row, col, _ = og.to_sparse_adj_t(skewed_edges(n, e, seed).numpy(), n)
    r, c = og.to_symmetric(row, col, n)
    return torch.from_numpy(r), torch.from_numpy(c)


def test_ginconv_sum_aggregation_plus_mlp():
    n, F, Hd = 1500, 32, 48
    r, c = _sym_graph(n, 8000, 4)
    torch.manual_seed(0)
    mlp = torch.nn.Sequential(bnn.Linear(F, Hd), torch.nn.ReLU(), bnn.Linear(Hd, Hd))
    conv = bnn.GINConv(mlp, eps=0.25, train_eps=True).cuda()
    x = torch.randn(n, F)
    adj = SparseTensor(row=r.cuda(), col=c.cuda(), sparse_sizes=(n, n), is_sorted=True)
    xc = x.cuda().requires_grad_(True)
    out = conv(xc, adj)
    out.pow(2).sum().backward()
    xr = x.double().requires_grad_(True)
    W1, b1, W2, b2 = (p.detach().cpu().double() for p in (mlp[0].weight, mlp[0].bias, mlp[2].weight, mlp[2].bias))
    eps = torch.tensor(0.25, dtype=torch.double, requires_grad=True)
    h = oo.spmm_scatter(r, c, None, xr, n, "sum") + (1 + eps) * xr
    ref = torch.relu(h @ W1.t() + b1) @ W2.t() + b2
    ref.pow(2).sum().backward()
    assert rel_err(out, ref) < 1e-5
    assert rel_err(xc.grad, xr.grad) < 2e-5
    assert abs(float(conv.eps.grad) - float(eps.grad)) < 2e-5 * abs(float(eps.grad))
    # the shim exports it under the name the reference would import
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(bnn.__file__).parent / "shim"))
    from torch_geometric.nn import GINConv  # noqa: F401
    assert GINConv is bnn.GINConv


def test_mean_spmm_of_weighted_matrix_gradient():
    """matmul(adj_with_values, x, 'mean') = sum(val * x_j) / rowcount; d x must carry val / rowcount."""
    n, K = 900, 24
    r, c = _sym_graph(n, 5000, 5)
    g = torch.Generator().manual_seed(6)
    v = torch.rand(r.numel(), generator=g) + 0.5
    x = torch.randn(n, K, generator=g)
    adj = SparseTensor(row=r.cuda(), col=c.cuda(), value=v.cuda(), sparse_sizes=(n, n), is_sorted=True)
    xc = x.cuda().requires_grad_(True)
    out = adj.matmul(xc, "mean")
    w = torch.randn(n, K, generator=g)
    (out * w.cuda()).sum().backward()
    xr = x.double().requires_grad_(True)
    ref = oo.spmm_scatter(r, c, v.double(), xr, n, "mean")
    (ref * w.double()).sum().backward()
    assert rel_err(out, ref) < 1e-5
    assert rel_err(xc.grad, xr.grad) < 1e-5


def test_spmm_accepts_extra_trailing_rows_of_the_dense_operand():
    """mag_pyg/gnn.py:151-162: SparseTensor(row=,col=) infers n_cols = max(col)+1, x has all nodes of the source type."""
    row, col = torch.tensor([0, 0, 2, 3]), torch.tensor([1, 4, 0, 4])
    adj = SparseTensor(row=row.cuda(), col=col.cuda())
    assert adj.sizes() == [4, 5]
    x = torch.randn(9, 8)
    out = adj.matmul(x.cuda(), "mean")
    ref = oo.spmm_scatter(row, col, None, x.double(), 4, "mean")
    assert rel_err(out, ref) < 1e-6


@pytest.mark.parametrize("concat", [True, False])
def test_pyg_gatconv_forward_and_gradients(concat):
    n, Fin, H, C = 700, 40, 4, 12
    ei = skewed_edges(n, 4500, 7)
    torch.manual_seed(2)
    layer = bnn.GATConv(Fin, C, heads=H, concat=concat).cuda()
    with torch.no_grad():
        layer.bias.normal_()
    x = torch.randn(n, Fin)
    xc = x.cuda().requires_grad_(True)
    out = layer(xc, ei.cuda())
    assert out.shape == (n, H * C if concat else C)
    wsum = torch.randn(out.shape, generator=torch.Generator().manual_seed(3))
    (out * wsum.cuda()).sum().backward()
    W, al, ar, b = (p.detach().cpu().double().requires_grad_(True) for p in (layer.lin_l.weight, layer.att_l, layer.att_r, layer.bias))
    xr = x.double().requires_grad_(True)
    xl = (xr @ W.t()).view(n, H, C)
    loops = torch.arange(n)
    keep = ei[0] != ei[1]
    row, col = torch.cat([ei[1][keep], loops]), torch.cat([ei[0][keep], loops])
    agg = onn.gat_aggregate(xl.reshape(n, H * C), (xl * al).sum(-1), (xl * ar).sum(-1), row, col, n, H, 0.2, 1e-16)
    ref = (agg if concat else agg.view(n, H, C).mean(1)) + b
    (ref * wsum.double()).sum().backward()
    assert rel_err(out, ref) < 1e-5
    assert rel_err(xc.grad, xr.grad) < 5e-5
    assert rel_err(layer.lin_l.weight.grad, W.grad) < 5e-5
    assert rel_err(layer.att_l.grad, al.grad) < 5e-5 and rel_err(layer.att_r.grad, ar.grad) < 5e-5
    assert rel_err(layer.bias.grad, b.grad) < 1e-5


def test_empty_relation_and_empty_node_mask():
    """RGCNConv's root_lins[i](x[mask]) with an empty mask and a relation without edges (mag_pyg/gnn.py:54-68)."""
    lin = bnn.Linear(16, 8).cuda()
    x = torch.randn(10, 16).cuda()
    assert lin(x[:0]).shape == (0, 8)

    class Conv(bnn.MessagePassing):
        def __init__(self):
            super().__init__(aggr="mean")
            self.lin = bnn.Linear(16, 8)

        def message(self, x_j):
            return self.lin(x_j)
    conv = Conv().cuda()
    out = conv.propagate(torch.zeros(2, 0, dtype=torch.long).cuda(), x=x)
    assert out.shape == (10, 8) and float(out.abs().max()) == 0.0
