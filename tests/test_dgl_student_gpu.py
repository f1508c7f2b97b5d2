"""DGL-side users of the aggregation kernel: GraphConv(norm='both') (the arxiv_dgl GCN student, models.py:46-92) and the
SIGN neighbour-averaging precompute (sign.py:175-183), against the fp64 oracle."""
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


def directed_graph(n, e, seed):
    ei = skewed_edges(n, e, seed).numpy()
    row, col, _ = og.to_sparse_adj_t(ei, n)           # row = destination, col = source, sorted by (row, col)
    return torch.from_numpy(row), torch.from_numpy(col)


@pytest.mark.parametrize("fin,fout", [(128, 40), (64, 256), (96, 96)])
def test_graph_conv_both_forward_backward(fin, fout):
    n = 3000
    r, c = directed_graph(n, 24_000, fin)             # directed: in- and out-degrees differ, some are zero
    g = torch.Generator().manual_seed(fout)
    x, w = torch.randn(n, fin, generator=g), torch.randn(n, fout, generator=g)
    conv = bnn.DGLGraphConv(fin, fout, "both", bias=True).cuda()
    with torch.no_grad():
        conv.bias.copy_(torch.randn(fout, generator=g))
    xr = x.double().requires_grad_(True)
    Wr, br = conv.weight.detach().cpu().double().requires_grad_(True), conv.bias.detach().cpu().double().requires_grad_(True)
    ref = onn.dgl_graph_conv_both(xr, r, c, n, Wr, br)
    (ref * w.double()).sum().backward()
    adj = SparseTensor(row=r.cuda(), col=c.cuda(), sparse_sizes=(n, n), is_sorted=True)
    xc = x.cuda().requires_grad_(True)
    out = conv(adj, xc)
    (out * w.cuda()).sum().backward()
    assert rel_err(out, ref) < 1e-5
    assert rel_err(xc.grad, xr.grad) < 1e-5
    assert rel_err(conv.weight.grad, Wr.grad) < 1e-5
    assert rel_err(conv.bias.grad, br.grad) < 1e-5
    assert torch.equal(conv(adj, xc), out)            # cached normalisation, deterministic kernels


def test_sign_neighbor_average_features():
    n, F, R = 5000, 128, 5                             # R = 5 is the reference default (sign.py --R)
    r, c = directed_graph(n, 40_000, 3)
    x = torch.randn(n, F, generator=torch.Generator().manual_seed(0))
    ref = onn.neighbor_average_features(x.double(), r, c, n, R)
    adj = SparseTensor(row=r.cuda(), col=c.cuda(), sparse_sizes=(n, n), is_sorted=True)
    got = bnn.neighbor_average_features(adj, x.cuda(), R)
    assert len(got) == R + 1 and torch.equal(got[0].cpu(), x)
    for hop in range(1, R + 1):
        assert rel_err(got[hop], ref[hop]) < 1e-5
    deg = torch.bincount(r, minlength=n)
    assert (deg == 0).any() and torch.all(got[1].cpu()[deg == 0] == 0)   # no in-edges => zeros, as DGL's mean reducer
