"""Fused GraphSAGE student step (engine_sage.SAGEStudentTrainer) vs the fp64 restatement of SAGE.forward (arxiv_pyg/gnn.py:77-85)
+ kd_criterion + backward + Adam; dropout masks injected; the kd + beta*G-CRD form of BASELINE configs[2]; graph replay."""
import numpy as np
import pytest
import torch

import efficient_gnns_b200  # noqa: F401
from conftest import rel_err
from efficient_gnns_b200 import criterion as C, ops
from efficient_gnns_b200.engine_sage import SAGEStudentTrainer
from efficient_gnns_b200.sparse import SparseTensor
from efficient_gnns_b200.synthetic import skewed_edges
from oracle import criterion as oc, graph as og, nn as onn

pytestmark = pytest.mark.gpu


def build(n=3000, e=20_000, dims=(32, 64, 64, 8), p=0.5, seed=0):
    ei = skewed_edges(n, e, seed)
    row, col, _ = og.to_sparse_adj_t(ei.numpy(), n)
    r, c = og.to_symmetric(row, col, n)
    adj = SparseTensor(row=torch.from_numpy(r).cuda(), col=torch.from_numpy(c).cuda(), sparse_sizes=(n, n), is_sorted=True)
    tr = SAGEStudentTrainer(adj, list(dims), dropout=p, lr=0.01, seed=seed)
    g = torch.Generator().manual_seed(seed + 9)
    x = torch.randn(n, dims[0], generator=g)
    y = torch.randint(0, dims[-1], (n,), generator=g)
    t = torch.randn(n, dims[-1], generator=g) * 2
    idx = torch.randperm(n, generator=g)[: n // 2].sort().values
    return tr, (r, c), x, y, t, idx


def oracle(tr, rc, x, y, t, idx, masks, aux=None, beta=0.0):
    n = x.shape[0]
    ptr, c = torch.from_numpy(og.ind2ptr(rc[0], n)), torch.from_numpy(rc[1])
    sd = {k: v.cpu().double() for k, v in tr.state_dict().items()}
    L = tr.L
    params = [dict(w_l=sd[f"convs.{i}.lin_l.weight"].clone().requires_grad_(True), b_l=sd[f"convs.{i}.lin_l.bias"].clone().requires_grad_(True),
                   w_r=sd[f"convs.{i}.lin_r.weight"].clone().requires_grad_(True)) for i in range(L)]
    ga = [sd[f"bns.{i}.weight"].clone().requires_grad_(True) for i in range(L - 1)]
    be = [sd[f"bns.{i}.bias"].clone().requires_grad_(True) for i in range(L - 1)]
    logits, hidden = onn.sage_forward(x.double(), ptr, c, params, ga, be, masks, p=tr.p)
    loss, lc, la = oc.kd_criterion(logits[idx], y[idx], t[idx].double(), tr.alpha, tr.kd_T)
    total = loss if aux is None else loss + beta * aux(hidden)
    flat = []
    for i in range(L):
        flat += [params[i]["w_l"], params[i]["b_l"], params[i]["w_r"]]
        if i < L - 1:
            flat += [ga[i], be[i]]
    grads = torch.autograd.grad(total, flat)
    return logits.detach(), hidden.detach(), float(total), grads


def engine_grads(tr):
    got = []
    for l in range(tr.L):
        got += [tr.gWl[l], tr.gbl[l], tr.gWr[l]]
        if l < tr.L - 1:
            got += [tr.ggamma[l], tr.gbeta[l]]
    return got


@pytest.mark.parametrize("p", [0.0, 0.5])
def test_sage_step_matches_oracle(p):
    tr, rc, x, y, t, idx = build(p=p)
    n = x.shape[0]
    masks = None
    if p > 0:
        masks = [ops.dropout_mask(n, tr.dims[l + 1], p, tr.seed, tr.dropout_offset(l, 0)).cpu().bool() for l in range(tr.L - 1)]
    ref_logits, ref_hidden, ref_loss, ref_grads = oracle(tr, rc, x, y, t, idx, masks)
    loss = tr.train_step(x.cuda(), y.cuda(), idx.cuda(), t.cuda()).cpu()
    assert rel_err(tr.Y[-1], ref_logits) < 1e-5 and rel_err(tr.A[-1], ref_hidden) < 1e-5
    assert abs(float(loss[0]) - ref_loss) < 1e-5 * abs(ref_loss)
    for i, (a, b) in enumerate(zip(engine_grads(tr), ref_grads)):
        is_hidden_bias = (i % 5 == 1) and i < 5 * (tr.L - 1)       # lin_l.bias in front of BatchNorm: exact gradient 0
        if is_hidden_bias:
            assert a.abs().max().item() < 1e-5 * max(g.abs().max().item() for g in ref_grads)
        else:
            assert rel_err(a, b) < 2e-5, f"grad {i}"
    assert int(tr.step_count.item()) == 1


def test_sage_with_gcrd_auxiliary_loss_and_graph_replay():
    """BASELINE configs[2]: SAGE + G-CRD (nce_criterion on projected out_feat, arxiv_pyg/gnn_kd_and_aux.py:161-171)."""
    tr, rc, x, y, t, idx = build(p=0.0)
    n, H = x.shape[0], tr.dims[-2]
    g = torch.Generator().manual_seed(5)
    t_feat = torch.randn(n, 16, generator=g)
    head = torch.randn(16, H, generator=g) * 0.2
    beta = 0.5
    ic, yc = idx.cuda(), y.cuda()
    hc, tfc = head.cuda(), t_feat.cuda()
    aux_gpu = lambda f: C.nce_criterion(tr.Y[-1][ic], yc[ic], f[ic] @ hc.t(), tfc[ic], 1.0, 0.075, 10 ** 9)[2]
    aux_ref = lambda hid: oc.nce_criterion(None, None, hid[idx] @ head.double().t(), t_feat[idx].double(), 1.0, 0.075, 10 ** 9)[2] \
        if False else _nce_ref(hid[idx] @ head.double().t(), t_feat[idx].double())
    _, _, ref_total, ref_grads = oracle(tr, rc, x, y, t, idx, None, aux_ref, beta)
    loss = tr.train_step(x.cuda(), yc, ic, t.cuda(), aux=aux_gpu, beta=beta).cpu()
    assert abs(float(loss[0]) - ref_total) < 2e-5 * abs(ref_total)
    for i, (a, b) in enumerate(zip(engine_grads(tr), ref_grads)):
        if (i % 5 == 1) and i < 5 * (tr.L - 1):
            continue
        assert rel_err(a, b) < 1e-4, f"grad {i}"
    # CUDA-graph replay of the plain kd step equals the eager step bitwise
    tr1, _, x1, y1, t1, idx1 = build(p=0.5, seed=1)
    tr2, *_ = build(p=0.5, seed=1)
    args = (x1.cuda(), y1.cuda(), idx1.cuda(), t1.cuda())
    eager = [tr2.train_step(*args).clone() for _ in range(2)]
    tr1.capture(*args, warmup=1)
    tr1.reset_parameters(1)
    got = [tr1.replay().clone() for _ in range(2)]
    assert all(torch.equal(a, b) for a, b in zip(got, eager)) and torch.equal(tr1.params, tr2.params)


def _nce_ref(fs, ft, T=0.075):
    fs = torch.nn.functional.normalize(fs, p=2, dim=-1)
    ft = torch.nn.functional.normalize(ft, p=2, dim=-1)
    logits = fs @ ft.t() / T
    return torch.nn.functional.cross_entropy(logits, torch.arange(fs.shape[0]))
