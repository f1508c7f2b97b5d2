"""Round-2 fixtures (tests/golden/make_golden_r2.py: the reference's own files on hub graphs, K=256 / K=750, and the PPI
criterion file) reproduced by the oracle on the CPU and — marked gpu — by the CUDA path."""
import sys
from pathlib import Path

import pytest
import torch
import torch.nn.functional as F

import efficient_gnns_b200  # noqa: F401
from conftest import GOLDEN, rel_err
from oracle import criterion as oc, nn as onn

sys.path.insert(0, str(GOLDEN))
import inputs_r2 as R2  # noqa: E402


@pytest.fixture(scope="module")
def ppi():
    return torch.load(GOLDEN / "criterion_ppi.pt")


@pytest.fixture(scope="module")
def hub():
    return torch.load(GOLDEN / "model_arxiv_hub.pt"), R2.hub_model()


@pytest.fixture(scope="module")
def lspw():
    return torch.load(GOLDEN / "lsp_wide.pt"), R2.lsp_wide()


@pytest.fixture(scope="module")
def gatw():
    return torch.load(GOLDEN / "gat_wide.pt"), R2.gat_wide()


def _sub_err(full, rows, ref_rows, ref_norm):
    """max-norm error on the stored rows + relative error of the Frobenius norm of the whole tensor."""
    full = full.detach().double().cpu()
    return max(rel_err(full[rows], ref_rows), abs(full.norm().item() - float(ref_norm)) / float(ref_norm))


# ------------------------------------------------------------------------------------------------ oracle (CPU)
def test_seeded_inputs_are_the_ones_the_fixtures_were_made_from(hub, lspw, gatw):
    (G, hm), (L, lw), (A, gw) = hub, lspw, gatw
    assert R2.checksum(hm["x"], hm["y"], hm["train_idx"], torch.from_numpy(hm["c"])) == pytest.approx(G["input_checksum"], rel=1e-12)
    assert R2.checksum(lw["feat"], lw["t_feat"], lw["sub_edge_index"]) == pytest.approx(L["input_checksum"], rel=1e-12)
    assert R2.checksum(gw["x"], gw["w"], gw["col"]) == pytest.approx(A["input_checksum"], rel=1e-12)
    assert G["max_degree"] > 1000 and L["max_segment"] > 256 and A["max_degree"] > 256


def test_oracle_ppi_criteria_reproduce_reference_file(ppi):
    i, cases = ppi["inputs"], ppi["cases"]
    z0, y, t, f0, tf, ei = i["logits"], i["labels"], i["t_logits"], i["feat"], i["t_feat"], i["edge_index"]

    def aux(fn):
        def run(z, f):
            cls = oc.bce_with_logits(z, y)
            a = fn(z, f)
            return cls, a
        return run

    def check(name, beta, loss_fn):
        z, f = z0.clone().requires_grad_(True), f0.clone().requires_grad_(True)
        if name == "kd":
            out = oc.kd_criterion_ppi(z, y, t)
        else:
            cls = oc.bce_with_logits(z, y)
            a = loss_fn(z, f)
            out = (cls + beta * a, cls, a)
        c = cases[name]
        for a, b in zip(out, (c["loss"], c["loss_cls"], c["loss_aux"])):
            assert abs(float(a) - float(b)) <= 1e-5 * abs(float(b)), name
        gz = torch.autograd.grad(out[0], [z], retain_graph=True)[0]
        assert rel_err(gz, c["d_logits"]) < 1e-5, name
        if c["d_feat"] is not None:
            assert rel_err(torch.autograd.grad(out[0], [f])[0], c["d_feat"]) < 2e-5, name

    dummy = torch.zeros(z0.shape[0], dtype=torch.long)
    check("kd", 0, None)
    check("fitnet", 1000, lambda z, f: oc.fitnet_criterion(z0[:, :2], dummy, f, tf)[2])
    check("lpw_cosine", 100, lambda z, f: oc.lpw_criterion(z0[:, :2], dummy, f, tf, ei, "cosine")[2])
    check("nce_full", 0.5, lambda z, f: oc.nce_criterion(z0[:, :2], dummy, f, tf, 0.5, 0.075, 10 ** 9)[2])


def test_oracle_gcn_reproduces_hub_fixture(hub):
    G, hm = hub
    m = G["models"]["gcn"]
    n = hm["n"]
    from oracle import graph as og
    r, c, v = og.gcn_norm(hm["r"], hm["c"], n)
    ptr, c, v = torch.from_numpy(og.ind2ptr(r, n)), torch.from_numpy(c), torch.from_numpy(v)
    st = m["state"]
    W = [st[f"convs.{i}.weight"].clone().requires_grad_(True) for i in range(3)]
    B = [st[f"convs.{i}.bias"].clone().requires_grad_(True) for i in range(3)]
    ga = [st[f"bns.{i}.weight"].clone().requires_grad_(True) for i in range(2)]
    be = [st[f"bns.{i}.bias"].clone().requires_grad_(True) for i in range(2)]
    logits, hidden = onn.gcn_forward(hm["x"], ptr, c, v, W, B, ga, be, None)
    assert _sub_err(logits, hm["rows"], m["logits_rows"], m["logits_norm"]) < 1e-5
    assert _sub_err(hidden, hm["rows"], m["out_feat_rows"], m["out_feat_norm"]) < 1e-5
    loss = oc.cross_entropy(logits[hm["train_idx"]], hm["y"][hm["train_idx"]])
    assert abs(float(loss) - float(m["loss"])) < 1e-5 * float(m["loss"])
    loss.backward()
    for i in range(3):
        assert rel_err(W[i].grad, m["grads"][f"convs.{i}.weight"]) < 5e-5


def test_oracle_lsp_reproduces_wide_fixture(lspw):
    L, lw = lspw
    for k, (sf, stc) in lw["scales"].items():
        f = (lw["feat"] * sf).clone().requires_grad_(True)
        out = oc.lpw_criterion(lw["logits"], lw["labels"], f, lw["t_feat"] * stc, lw["sub_edge_index"], k, 100)
        c = L["cases"][k]
        for a, b in zip(out, (c["loss"], c["loss_cls"], c["loss_aux"])):
            assert abs(float(a) - float(b)) <= 2e-5 * max(abs(float(b)), 1e-4), k
        (gf,) = torch.autograd.grad(out[0], [f])
        assert _sub_err(gf, lw["rows"], c["d_feat_rows"], c["d_feat_norm"]) < 5e-5, k


def test_oracle_gat_reproduces_wide_fixture(gatw):
    A, gw = gatw
    # fp64 restatement: the fixture itself is an fp32 computation over rows with > 1,000 neighbours
    st = {k: t.double().requires_grad_(True) for k, t in A["state"].items()}
    x = gw["x"].double().requires_grad_(True)
    out = onn.dgl_gat_conv(x, gw["row"], gw["col"], gw["n"], st["fc.weight"], st["attn_l"], st["attn_r"], st["res_fc.weight"], 3)
    assert _sub_err(out, gw["rows"], A["out_rows"], A["out_norm"]) < 1e-5
    (out * gw["w"].double()).sum().backward()
    assert rel_err(x.grad, A["d_x"]) < 5e-5
    for k, t in st.items():
        assert rel_err(t.grad, A["grads"][k]) < 5e-5, k


# ------------------------------------------------------------------------------------------------ CUDA path
gpu = pytest.mark.gpu


@gpu
def test_cuda_ppi_criteria_reproduce_reference_file(ppi):
    from efficient_gnns_b200 import criterion_ppi as bp
    i, cases = ppi["inputs"], ppi["cases"]
    y, t, tf, ei = i["labels"].cuda(), i["t_logits"].cuda(), i["t_feat"].cuda(), i["edge_index"].cuda()

    def run(name, fn, tol=2e-5):
        z, f = i["logits"].cuda().requires_grad_(True), i["feat"].cuda().requires_grad_(True)
        out = fn(z, f)
        out[0].backward()
        c = cases[name]
        for a, b in zip(out, (c["loss"], c["loss_cls"], c["loss_aux"])):
            assert abs(float(a) - float(b)) <= tol * abs(float(b)), name
        assert rel_err(z.grad, c["d_logits"]) < tol, name
        if c["d_feat"] is not None:
            assert rel_err(f.grad, c["d_feat"]) < 5e-5, name
    run("kd", lambda z, f: bp.kd_criterion(z, y, t))
    run("fitnet", lambda z, f: bp.fitnet_criterion(z, y, f, tf))
    run("lpw_cosine", lambda z, f: bp.lpw_criterion(z, y, f, tf, ei, "cosine", 100))
    run("nce_full", lambda z, f: bp.nce_criterion(z, y, f, tf, 0.5, 0.075, 10 ** 9))


@gpu
def test_cuda_engine_reproduces_reference_gcn_on_hub_graph(hub):
    """Fused engine (hub-split rows, pipelined K=256 aggregation fwd + bwd) vs the reference's GCN class."""
    from efficient_gnns_b200.engine import GCNStudentTrainer
    from efficient_gnns_b200.sparse import SparseTensor
    G, hm = hub
    m, n = G["models"]["gcn"], hm["n"]
    adj = SparseTensor(row=torch.from_numpy(hm["r"]).cuda(), col=torch.from_numpy(hm["c"]).cuda(), sparse_sizes=(n, n),
                       is_sorted=True)
    for agg_first in (False, True):
        tr = GCNStudentTrainer(adj, [16, 256, 256, 8], dropout=0.0, aggregate_first=agg_first)
        assert tr.G.n_hub > 0
        tr.load_state_dict({k: v.cuda() for k, v in m["state"].items() if "num_batches" not in k})
        loss = tr.train_step(hm["x"].cuda(), hm["y"].cuda(), hm["train_idx"].cuda(), None).cpu()
        assert _sub_err(tr.Y[-1], hm["rows"], m["logits_rows"], m["logits_norm"]) < 1e-5
        assert _sub_err(tr.out_feat(), hm["rows"], m["out_feat_rows"], m["out_feat_norm"]) < 1e-5
        assert abs(loss[0] - m["loss"]) < 1e-5 * abs(m["loss"])
        for l in range(3):
            assert rel_err(tr.gW[l], m["grads"][f"convs.{l}.weight"]) < 5e-5, (agg_first, l)
        for l in range(2):
            assert rel_err(tr.ggamma[l], m["grads"][f"bns.{l}.weight"]) < 5e-5
            assert rel_err(tr.gbeta[l], m["grads"][f"bns.{l}.bias"]) < 5e-5


@gpu
@pytest.mark.parametrize("name", ["gcn", "sage"])
def test_cuda_module_path_reproduces_reference_on_hub_graph(hub, name):
    from efficient_gnns_b200 import nn as bnn
    from efficient_gnns_b200.sparse import SparseTensor
    from test_shim_gpu import Student
    G, hm = hub
    m, n = G["models"][name], hm["n"]
    adj = SparseTensor(row=torch.from_numpy(hm["r"]).cuda(), col=torch.from_numpy(hm["c"]).cuda(), sparse_sizes=(n, n),
                       is_sorted=True)
    cls, kw = (bnn.GCNConv, dict(cached=True)) if name == "gcn" else (bnn.SAGEConv, {})
    model = Student(cls, [16, 256, 256, 8], **kw).cuda()
    model.load_state_dict(m["state"])
    model.train()
    out = model(hm["x"].cuda(), adj)
    assert _sub_err(out, hm["rows"], m["logits_rows"], m["logits_norm"]) < 1e-5
    assert _sub_err(model.out_feat, hm["rows"], m["out_feat_rows"], m["out_feat_norm"]) < 1e-5
    idx = hm["train_idx"].cuda()
    loss = F.cross_entropy(out[idx], hm["y"].cuda()[idx])
    assert abs(loss.item() - m["loss"].item()) < 1e-5 * m["loss"].item()
    loss.backward()
    for k, p in model.named_parameters():
        if k.endswith("weight") and ("convs" in k):
            assert rel_err(p.grad, m["grads"][k]) < 5e-5, k


@gpu
def test_cuda_lsp_reproduces_wide_fixture(lspw):
    from efficient_gnns_b200 import criterion as bc
    L, lw = lspw
    sub = lw["sub_edge_index"].cuda()
    for k, (sf, stc) in lw["scales"].items():
        f = (lw["feat"] * sf).cuda().requires_grad_(True)
        out = bc.lpw_criterion(lw["logits"].cuda(), lw["labels"].cuda(), f, (lw["t_feat"] * stc).cuda(), sub, k, 100)
        out[0].backward()
        c = L["cases"][k]
        for a, b in zip(out, (c["loss"], c["loss_cls"], c["loss_aux"])):
            assert abs(float(a) - float(b)) <= 2e-5 * max(abs(float(b)), 1e-4), k
        assert _sub_err(f.grad, lw["rows"], c["d_feat_rows"], c["d_feat_norm"]) < 5e-5, k


@gpu
def test_cuda_gat_reproduces_wide_fixture(gatw):
    from efficient_gnns_b200 import nn as bnn
    from efficient_gnns_b200.sparse import SparseTensor
    A, gw = gatw
    n = gw["n"]
    adj = SparseTensor(row=gw["row"].cuda(), col=gw["col"].cuda(), sparse_sizes=(n, n), is_sorted=True)
    layer = bnn.DGLGATConv(32, 250, num_heads=3, residual=True, use_symmetric_norm=True, use_attn_dst=True).cuda()
    layer.load_state_dict(A["state"])
    x = gw["x"].cuda().requires_grad_(True)
    out = layer(adj, x)
    assert _sub_err(out, gw["rows"], A["out_rows"], A["out_norm"]) < 1e-5
    (out * gw["w"].cuda()).sum().backward()
    assert rel_err(x.grad, A["d_x"]) < 5e-5
    for k, p in layer.named_parameters():
        assert rel_err(p.grad, A["grads"][k]) < 5e-5, k
