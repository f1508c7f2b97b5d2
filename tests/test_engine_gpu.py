"""Fused GCN student step (engine) vs the CPU oracle: forward logits, KD loss, every gradient, one Adam update.
Dropout: the kernel's Philox mask is materialised through the C ABI and injected into the oracle."""
import numpy as np
import pytest
import torch

import efficient_gnns_b200  # noqa: F401
from conftest import rel_err
from efficient_gnns_b200 import ops
from efficient_gnns_b200.engine import GCNStudentTrainer, gcn_norm
from efficient_gnns_b200.sparse import SparseTensor
from efficient_gnns_b200.synthetic import skewed_edges
from oracle import criterion as oc, graph as og, nn as onn

pytestmark = pytest.mark.gpu


def build(n=3000, e=20_000, dims=(32, 64, 64, 8), p=0.5, seed=0, lr=0.01, agg_first=None):
    ei = skewed_edges(n, e, seed)
    row, col, _ = og.to_sparse_adj_t(ei.numpy(), n)
    r, c = og.to_symmetric(row, col, n)
    adj = SparseTensor(row=torch.from_numpy(r).cuda(), col=torch.from_numpy(c).cuda(), sparse_sizes=(n, n), is_sorted=True)
    tr = GCNStudentTrainer(adj, list(dims), dropout=p, lr=lr, seed=seed, aggregate_first=agg_first)
    g = torch.Generator().manual_seed(seed + 9)
    x = torch.randn(n, dims[0], generator=g)
    y = torch.randint(0, dims[-1], (n,), generator=g)
    t = torch.randn(n, dims[-1], generator=g) * 2
    idx = torch.randperm(n, generator=g)[: n // 2].sort().values
    return tr, (r, c), x, y, t, idx


def oracle_step(tr, rc, x, y, t, idx, masks, kd=True, lr=0.01):
    n = x.shape[0]
    r, c, v = og.gcn_norm(rc[0], rc[1], n)
    ptr, c, v = torch.from_numpy(og.ind2ptr(r, n)), torch.from_numpy(c), torch.from_numpy(v)
    sd = {k: w.cpu() for k, w in tr.state_dict().items()}
    L = tr.L
    W = [sd[f"convs.{i}.weight"].clone().requires_grad_(True) for i in range(L)]
    B = [sd[f"convs.{i}.bias"].clone().requires_grad_(True) for i in range(L)]
    ga = [sd[f"bns.{i}.weight"].clone().requires_grad_(True) for i in range(L - 1)]
    be = [sd[f"bns.{i}.bias"].clone().requires_grad_(True) for i in range(L - 1)]
    logits, hidden = onn.gcn_forward(x, ptr, c, v, W, B, ga, be, masks, p=tr.p)
    if kd:
        loss, lc, la = oc.kd_criterion(logits[idx], y[idx], t[idx], tr.alpha, tr.kd_T)
    else:
        loss = lc = oc.cross_entropy(logits[idx], y[idx]); la = loss * 0
    params = []
    for i in range(L):
        params += [W[i], B[i]]
        if i < L - 1:
            params += [ga[i], be[i]]
    opt = torch.optim.Adam(params, lr=lr)
    opt.zero_grad(); loss.backward()
    grads = [p_.grad.clone() for p_ in params]
    opt.step()
    return logits.detach(), hidden.detach(), (loss.detach(), lc.detach(), la.detach()), grads, [p_.detach() for p_ in params]


def flat(ts):
    return torch.cat([t.reshape(-1) for t in ts])


@pytest.mark.parametrize("agg_first", [False, True])
@pytest.mark.parametrize("p,kd", [(0.0, True), (0.5, True), (0.5, False)])
def test_gcn_step_matches_oracle(p, kd, agg_first):
    tr, rc, x, y, t, idx = build(p=p, agg_first=agg_first)
    assert tr.agg_first == agg_first
    n = x.shape[0]
    masks = None
    if p > 0:
        masks = [ops.dropout_mask(n, tr.dims[l + 1], p, tr.seed, tr.dropout_offset(l, 0)).cpu().bool() for l in range(tr.L - 1)]
        frac = float(masks[0].float().mean())
        assert abs(frac - (1 - p)) < 0.01
    ref_logits, ref_hidden, ref_loss, ref_grads, ref_params = oracle_step(tr, rc, x, y, t, idx, masks, kd)
    xc, yc, tc, ic = x.cuda(), y.cuda(), t.cuda(), idx.cuda()
    loss = tr.train_step(xc, yc, ic, tc if kd else None).cpu()
    assert rel_err(tr.Y[-1], ref_logits) < 1e-5
    assert rel_err(tr.A[-1], ref_hidden) < 1e-5
    assert abs(loss[0] - ref_loss[0]) < 1e-5 * abs(ref_loss[0]) + 1e-7
    assert abs(loss[1] - ref_loss[1]) < 1e-5 * abs(ref_loss[1]) + 1e-7
    if kd:
        assert abs(loss[2] - ref_loss[2]) < 1e-5 * abs(ref_loss[2]) + 1e-7
    # gradients, tensor by tensor (a conv bias in front of BatchNorm has zero gradient up to rounding noise)
    got = []
    for l in range(tr.L):
        got += [tr.gW[l], tr.gb[l]]
        if l < tr.L - 1:
            got += [tr.ggamma[l], tr.gbeta[l]]
    scale = max(g.abs().max().item() for g in ref_grads)
    for i, (a, b) in enumerate(zip(got, ref_grads)):
        is_hidden_bias = (i % 4 == 1) and i < 4 * (tr.L - 1)
        if is_hidden_bias:
            assert a.abs().max().item() < 1e-5 * scale
        else:
            assert rel_err(a, b) < 2e-5, f"grad {i}"
    # Adam update (skip the noise-driven hidden biases)
    new = []
    for l in range(tr.L):
        new += [tr.W[l], tr.b[l]]
        if l < tr.L - 1:
            new += [tr.gamma[l], tr.beta[l]]
    for i, (a, b) in enumerate(zip(new, ref_params)):
        if (i % 4 == 1) and i < 4 * (tr.L - 1):
            continue
        assert (a.cpu() - b).abs().max().item() < 2e-5, f"param {i}"
    assert int(tr.step_count.item()) == 1


def test_graph_capture_replays_identically_and_advances_dropout():
    tr, rc, x, y, t, idx = build(p=0.5, seed=1)
    tr2, *_ = build(p=0.5, seed=1)
    xc, yc, tc, ic = x.cuda(), y.cuda(), t.cuda(), idx.cuda()
    # eager reference run: 3 steps
    eager = [tr2.train_step(xc, yc, ic, tc).clone() for _ in range(3)]
    # captured run: capture() itself performs warm-up steps, so reset state afterwards
    tr.capture(xc, yc, ic, tc, warmup=1)
    tr.reset_parameters(1)
    for l in range(tr.L - 1):
        tr.running_mean[l].zero_(); tr.running_var[l].fill_(1.0)
    got = [tr.replay().clone() for _ in range(3)]
    torch.cuda.synchronize()
    for a, b in zip(got, eager):
        assert torch.equal(a, b)          # bitwise: same kernels, same order, deterministic reductions
    assert not torch.equal(got[0], got[1])
    assert torch.equal(tr.params, tr2.params)
    assert int(tr.step_count.item()) == 3


def test_eval_forward_uses_running_statistics():
    tr, rc, x, y, t, idx = build(p=0.5, seed=2)
    xc = x.cuda()
    tr.train_step(xc, y.cuda(), idx.cuda(), t.cuda())
    logits = tr.forward(xc, training=False).cpu()
    n = x.shape[0]
    r, c, v = og.gcn_norm(rc[0], rc[1], n)
    ptr, c, v = torch.from_numpy(og.ind2ptr(r, n)), torch.from_numpy(c), torch.from_numpy(v)
    sd = {k: w.cpu() for k, w in tr.state_dict().items()}
    h = x
    for l in range(tr.L):
        h = onn.gcn_conv(h, ptr, c, v, sd[f"convs.{l}.weight"], sd[f"convs.{l}.bias"])
        if l < tr.L - 1:
            h = (h - sd[f"bns.{l}.running_mean"]) / torch.sqrt(sd[f"bns.{l}.running_var"] + 1e-5) * sd[f"bns.{l}.weight"] + sd[f"bns.{l}.bias"]
            h = torch.relu(h)
    assert rel_err(logits, h) < 1e-5


def test_sharded_trainer_world1_matches_plain_engine():
    """The node-parallel code path (relabelling, row-shard CSR, all-gather / all-reduce plumbing, two-phase BatchNorm
    backward) on a 1-rank NCCL group must reproduce the plain engine; multi-rank equivalence is tests/dist_equiv.py."""
    import os
    import socket
    import torch.distributed as dist
    from efficient_gnns_b200.dist import ShardedGCNTrainer
    if not dist.is_initialized():
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    tr, rc, x, y, t, idx = build(p=0.0, seed=4)
    n = x.shape[0]
    adj = SparseTensor(row=torch.from_numpy(rc[0]).cuda(), col=torch.from_numpy(rc[1]).cuda(), sparse_sizes=(n, n), is_sorted=True)
    sh = ShardedGCNTrainer(adj, tr.dims, dropout=0.0, seed=4)
    xc, yc, tc, ic = x.cuda(), y.cuda(), t.cuda(), idx.cuda()
    l_ref = tr.train_step(xc, yc, ic, tc).clone()
    l_sh = sh.train_step(*sh.shard_inputs(xc, yc, ic, tc)).clone()
    assert rel_err(sh.gather_rows(sh.Y[-1]), tr.Y[-1]) < 1e-5
    assert ((l_sh - l_ref).abs() / l_ref.abs()).max().item() < 1e-5
    for l in range(tr.L):
        assert ((sh.gW[l] - tr.gW[l]).norm() / tr.gW[l].norm()).item() < 2e-3
    assert sh.exchange_bytes_per_step() == 0


def test_config0_plumbing_shape_two_layer_gcn():
    """BASELINE.json configs[0]: 2-layer GCN, synthetic 10k-node / 50k-edge graph, 64-d — the reference's CPU-runnable
    case, here as a parity case of the CUDA engine against the oracle (supervised step, dropout injected)."""
    from efficient_gnns_b200.synthetic import PLUMBING, make_node_dataset
    ds = make_node_dataset(PLUMBING, seed=0)
    n = ds.num_nodes
    row, col, _ = og.to_sparse_adj_t(ds.edge_index.numpy(), n)
    r, c = og.to_symmetric(row, col, n)
    adj = SparseTensor(row=torch.from_numpy(r).cuda(), col=torch.from_numpy(c).cuda(), sparse_sizes=(n, n), is_sorted=True)
    tr = GCNStudentTrainer(adj, [64, 64, 40], dropout=0.5, lr=0.01, seed=0)
    masks = [ops.dropout_mask(n, 64, 0.5, tr.seed, tr.dropout_offset(0, 0)).cpu().bool()]
    x, y, idx = ds.x, ds.y.squeeze(1), ds.split_idx["train"]
    ref_logits, ref_hidden, ref_loss, ref_grads, _ = oracle_step(tr, (r, c), x, y, ds.teacher_logits, idx, masks, kd=False)
    loss = tr.train_step(x.cuda(), y.cuda(), idx.cuda(), None).cpu()
    assert rel_err(tr.Y[-1], ref_logits) < 1e-5 and rel_err(tr.A[-1], ref_hidden) < 1e-5
    assert abs(loss[0] - ref_loss[0]) < 1e-5 * abs(ref_loss[0])
    assert rel_err(tr.gW[1], ref_grads[4]) < 2e-5 and rel_err(tr.gW[0], ref_grads[0]) < 5e-5


@pytest.mark.parametrize("aux_name", ["lpw", "nce_head", "at"])
def test_gcn_step_with_auxiliary_distillation_loss_matches_oracle(aux_name):
    """kd + beta*aux, the train() of arxiv_pyg/gnn_kd_and_aux.py:100-189, on the fused engine: the auxiliary criterion runs on
    out_feat[train_idx] (LSP on the train-induced subgraph :150-160; a projection head + nce_criterion :161-171; AT :128-137),
    its gradient w.r.t. out_feat seeds the input-gradient GEMM (accumulating epilogue) and every engine gradient must match
    the fp64 restatement of the whole step; a torch-side head receives its gradient through autograd."""
    from efficient_gnns_b200 import criterion as C
    tr, rc, x, y, t, idx = build(p=0.0, dims=(32, 64, 64, 8))
    n, H = x.shape[0], tr.dims[-2]
    g = torch.Generator().manual_seed(3)
    t_feat = torch.randn(n, 24, generator=g).relu() + 0.01
    m = idx.numel()
    sub = torch.from_numpy(og.subgraph(idx.numpy(), np.stack(rc), True)[0])          # arxiv_pyg/gnn.py:249
    head_w = (torch.randn(16, H, generator=g) * 0.2)
    beta = 0.7
    yc, ic, tc, tfc = y.cuda(), idx.cuda(), t.cuda(), t_feat.cuda()
    head = torch.nn.Parameter(head_w.clone().cuda())

    def aux_gpu(f):
        z = tr.Y[-1][ic]
        if aux_name == "lpw":
            return C.lpw_criterion(z, yc[ic], f[ic], tfc[ic], sub.cuda(), "cosine", 1.0)[2]
        if aux_name == "at":
            return C.at_criterion(z, yc[ic], f[ic], tfc[ic], 1.0)[2]
        return C.nce_criterion(z, yc[ic], f[ic] @ head.t(), tfc[ic][:, :16], 1.0, 0.075, 10 ** 9)[2]

    loss = tr.train_step(x.cuda(), yc, ic, tc, aux=aux_gpu, beta=beta).cpu()

    # fp64 restatement of the same step
    r, c, v = og.gcn_norm(rc[0], rc[1], n)
    ptr, c, v = torch.from_numpy(og.ind2ptr(r, n)), torch.from_numpy(c), torch.from_numpy(v).double()
    sd = {k: w.cpu().double() for k, w in tr.state_dict().items()}
    # (state_dict is read AFTER the step: rebuild the pre-step parameters from a twin engine with the same seed)
    tr0, *_ = build(p=0.0, dims=(32, 64, 64, 8))
    sd = {k: w.cpu().double() for k, w in tr0.state_dict().items()}
    L = tr.L
    W = [sd[f"convs.{i}.weight"].clone().requires_grad_(True) for i in range(L)]
    B = [sd[f"convs.{i}.bias"].clone().requires_grad_(True) for i in range(L)]
    ga = [sd[f"bns.{i}.weight"].clone().requires_grad_(True) for i in range(L - 1)]
    be = [sd[f"bns.{i}.bias"].clone().requires_grad_(True) for i in range(L - 1)]
    hw = head_w.double().clone().requires_grad_(True)
    logits, hidden = onn.gcn_forward(x.double(), ptr, c, v, W, B, ga, be, None, p=0.0)
    kd, lc, la = oc.kd_criterion(logits[idx], y[idx], t[idx].double(), tr.alpha, tr.kd_T)
    f, tf = hidden[idx], t_feat[idx].double()
    if aux_name == "lpw":
        aux = oc.lpw_criterion(logits[idx], y[idx], f, tf, sub, "cosine", 1.0)[2]
    elif aux_name == "at":
        aux = oc.at_criterion(logits[idx], y[idx], f, tf, 1.0)[2]
    else:
        aux = oc.nce_criterion(logits[idx], y[idx], f @ hw.t(), tf[:, :16], 1.0, 0.075, 10 ** 9)[2]
    total = kd + beta * aux
    params = []
    for i in range(L):
        params += [W[i], B[i]]
        if i < L - 1:
            params += [ga[i], be[i]]
    grads = torch.autograd.grad(total, params + [hw], allow_unused=True)
    assert abs(loss[0] - float(total)) < 2e-5 * abs(float(total))
    assert abs(float(tr.loss_aux) - float(aux)) < 2e-5 * abs(float(aux)) + 1e-9
    got = []
    for l in range(L):
        got += [tr.gW[l], tr.gb[l]]
        if l < L - 1:
            got += [tr.ggamma[l], tr.gbeta[l]]
    scale = max(gr.abs().max().item() for gr in grads[:-1])
    for i, (a, b) in enumerate(zip(got, grads[:-1])):
        if (i % 4 == 1) and i < 4 * (L - 1):
            assert a.abs().max().item() < 1e-5 * scale
        else:
            # InfoNCE at T = 0.075 multiplies the logits by 13: its own gradient test runs at 5e-5, through the network 1e-4
            assert rel_err(a, b) < (1e-4 if aux_name == "nce_head" else 5e-5), f"grad {i}"
    if aux_name == "nce_head":
        assert rel_err(head.grad, grads[-1] ) < 5e-5


@pytest.mark.parametrize("dims", [(32, 64, 64, 8), (64, 256, 128, 40)])
def test_fused_row_passes_equal_the_separate_passes(dims):
    """fuse_row_passes (BatchNorm statistics / backward reductions inside GEMM epilogues) changes WHERE the column sums are
    taken, not what is computed: same masks, the same logits, loss and gradients up to the summation order of the partials."""
    n, e = 6000, 50_000
    ei = skewed_edges(n, e, 3)
    row, col, _ = og.to_sparse_adj_t(ei.numpy(), n)
    r, c = og.to_symmetric(row, col, n)
    adj = SparseTensor(row=torch.from_numpy(r).cuda(), col=torch.from_numpy(c).cuda(), sparse_sizes=(n, n), is_sorted=True)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(n, dims[0], generator=g).cuda()
    y = torch.randint(0, dims[-1], (n,), generator=g).cuda()
    t = (torch.randn(n, dims[-1], generator=g) * 2).cuda()
    idx = torch.randperm(n, generator=g)[: n // 2].sort().values.cuda()
    outs = []
    for fuse in (True, False):
        tr = GCNStudentTrainer(adj, list(dims), dropout=0.5, lr=0.01, seed=1, fuse_row_passes=fuse)
        assert bool(tr._gemm_part) == fuse
        tr.train_step(x, y, idx, t)
        torch.cuda.synchronize()
        outs.append((tr.Y[-1].clone(), tr.loss_out.clone(), tr.grads.clone(), [a.clone() for a in tr.A]))
    (lf, ll, gf, af), (lu, lo_, gu, au) = outs
    # same dropout masks; a pre-activation within ~1e-7 of zero may land on the other side of the ReLU because the layer-0
    # batch statistics are summed in a different order — count such flips instead of assuming there are none
    flips = sum(int(((a > 0) != (b > 0)).sum()) for a, b in zip(af, au))
    assert flips <= 3, flips
    tol = 5e-6 if flips == 0 else 5e-4
    assert rel_err(lf, lu) < tol and rel_err(ll, lo_) < tol
    # (the updated parameters are not compared: the conv biases in front of a BatchNorm have analytically ZERO gradient, what
    # both engines hold there is rounding noise, and Adam's g / (|g| + eps) turns noise of either sign into +-lr)
    assert rel_err(gf, gu) < tol, (rel_err(gf, gu), flips)
