"""The fused criteria (efficient_gnns_b200.criterion, same API as the reference's criterion.py) vs
 (a) the fp64 CPU oracle on random inputs and (b) the fixtures the reference's own criterion.py produced."""
import numpy as np
import pytest
import torch

import efficient_gnns_b200  # noqa: F401
from conftest import rel_err
from efficient_gnns_b200 import criterion as C
from oracle import criterion as oc, graph as og
from efficient_gnns_b200.synthetic import skewed_edges

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def data(n=700, Cn=12, Fs=64, Ft=96, seed=0):
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(n, Cn, generator=g)
    y = torch.randint(0, Cn, (n,), generator=g)
    t = torch.randn(n, Cn, generator=g) * 2
    f = torch.randn(n, Fs, generator=g).relu() + 0.01
    tf = torch.randn(n, Ft, generator=g).relu() + 0.01
    tf_same = torch.randn(n, Fs, generator=g)
    return z, y, t, f, tf, tf_same


def check(fn_gpu, fn_ref, z, f, tf, tol=2e-5, teacher_grad=False):
    zr, fr = z.double().requires_grad_(True), f.double().requires_grad_(True)
    tr = tf.double().requires_grad_(teacher_grad)
    out_r = fn_ref(zr, fr, tr)
    out_r[0].backward()
    zc, fc = z.cuda().requires_grad_(True), f.cuda().requires_grad_(True)
    tc = tf.cuda().requires_grad_(teacher_grad)
    out_c = fn_gpu(zc, fc, tc)
    out_c[0].backward()
    for a, b in zip(out_c, out_r):
        assert abs(a.item() - b.item()) <= tol * max(abs(b.item()), 1e-6), (a.item(), b.item())
    assert rel_err(zc.grad, zr.grad) < tol
    if fr.grad is not None:
        assert rel_err(fc.grad, fr.grad) < tol
    if teacher_grad:
        assert rel_err(tc.grad, tr.grad) < tol


def test_kd_and_ce():
    z, y, t, f, tf, _ = data()
    check(lambda a, b, c: C.kd_criterion(a, y.cuda(), t.cuda(), 0.9, 4.0), lambda a, b, c: oc.kd_criterion(a, y, t.double(), 0.9, 4.0), z, f, tf)
    zc = z.cuda().requires_grad_(True)
    l = C.cross_entropy(zc, y.cuda()); l.backward()
    zr = z.double().requires_grad_(True); lr = oc.cross_entropy(zr, y); lr.backward()
    assert abs(l.item() - lr.item()) < 1e-5 * lr.item() and rel_err(zc.grad, zr.grad) < 1e-5


def test_fitnet_and_at():
    z, y, t, f, tf, tf_same = data()
    check(lambda a, b, c: C.fitnet_criterion(a, y.cuda(), b, c, 1000), lambda a, b, c: oc.fitnet_criterion(a, y, b, c, 1000),
          z, f, tf_same, teacher_grad=True)
    check(lambda a, b, c: C.at_criterion(a, y.cuda(), b, c, 1000), lambda a, b, c: oc.at_criterion(a, y, b, c, 1000), z, f, tf,
          teacher_grad=True)


@pytest.mark.parametrize("kernel", ["cosine", "poly", "l2", "rbf"])
def test_gsp(kernel):
    z, y, t, f, tf, _ = data(n=500, Fs=32, Ft=48)
    f, tf = f * 0.3, tf * 0.3                      # keep rbf away from underflow
    inds = torch.randperm(500, generator=torch.Generator().manual_seed(1))[:200]
    check(lambda a, b, c: C.gpw_criterion(a, y.cuda(), b, c, kernel, 1.0, 200, inds.cuda()),
          lambda a, b, c: oc.gpw_criterion(a, y, b, c, kernel, 1.0, 200, inds), z, f, tf, tol=5e-5, teacher_grad=True)
    check(lambda a, b, c: C.gpw_criterion(a, y.cuda(), b, c, kernel, 1.0, 10 ** 9),
          lambda a, b, c: oc.gpw_criterion(a, y, b, c, kernel, 1.0, 10 ** 9), z, f, tf, tol=5e-5)


@pytest.mark.parametrize("kernel", ["cosine", "poly", "l2", "rbf"])
@pytest.mark.parametrize("crit", ["kld", "mse"])
def test_lsp(kernel, crit):
    n = 800
    z, y, t, f, tf, _ = data(n=n, Fs=64, Ft=750)
    if kernel in ("l2", "rbf"):
        f, tf = f * 0.2, tf * 0.05
    ei = skewed_edges(n, 6000, 3).numpy()
    row, col, _ = og.to_sparse_adj_t(ei, n)
    r, c = og.to_symmetric(row, col, n)
    sub = torch.arange(0, n, 2)
    ei_sub = torch.from_numpy(og.subgraph(sub.numpy(), np.stack([r, c]), True)[0])     # unsorted dst, like the reference's
    m = sub.numel()
    check(lambda a, b, cc: C.lpw_criterion(a[:m], y[:m].cuda(), b[:m], cc[:m], ei_sub.cuda(), kernel, 100, crit),
          lambda a, b, cc: oc.lpw_criterion(a[:m], y[:m], b[:m], cc[:m], ei_sub, kernel, 100, crit), z, f, tf, tol=5e-5)


def test_nce():
    z, y, t, f, tf, tf_same = data(n=900, Fs=64)
    inds = torch.randperm(900, generator=torch.Generator().manual_seed(2))[:512]
    check(lambda a, b, c: C.nce_criterion(a, y.cuda(), b, c, 0.5, 0.075, 512, inds.cuda()),
          lambda a, b, c: oc.nce_criterion(a, y, b, c, 0.5, 0.075, 512, inds), z, f, tf_same, tol=5e-5, teacher_grad=True)


def test_numpy_rng_sampling_matches_reference_draw():
    z, y, t, f, tf, tf_same = data(n=900, Fs=64)
    np.random.seed(7); expect = np.random.choice(900, 300, replace=False)
    np.random.seed(7)
    a = C.nce_criterion(z.cuda(), y.cuda(), f.cuda(), tf_same.cuda(), 0.5, 0.075, 300)
    b = C.nce_criterion(z.cuda(), y.cuda(), f.cuda(), tf_same.cuda(), 0.5, 0.075, 300, torch.from_numpy(expect))
    assert torch.equal(a[2], b[2])


def test_criteria_reproduce_reference_fixtures(golden_criterion):
    G = golden_criterion
    i, cases = G["inputs"], G["cases"]
    z, y, t = i["logits"], i["labels"].cuda(), i["t_logits"].cuda()
    f, tf, same = i["feat"], i["t_feat"].cuda(), i["same_t_feat"].cuda()
    sub, draw = i["sub_edge_index"].cuda(), i["np_draw"].cuda()

    def run(name, fn, tol=5e-5):
        zc, fc = z.cuda().requires_grad_(True), f.cuda().requires_grad_(True)
        out = fn(zc, fc)
        out[0].backward()
        c = cases[name]
        for a, b in zip(out, (c["loss"], c["loss_cls"], c["loss_aux"])):
            assert abs(a.item() - b.item()) <= tol * max(abs(b.item()), 1e-4), (name, a.item(), b.item())  # aux terms ~1e-5 are fp32 cancellation residue
        assert rel_err(zc.grad, c["d_logits"]) < tol, name
        if c["d_feat"] is not None:
            assert rel_err(fc.grad, c["d_feat"]) < tol, name

    run("kd", lambda a, b: C.kd_criterion(a, y, t, 0.9, 4.0))
    run("fitnet", lambda a, b: C.fitnet_criterion(a, y, b, same, 1000))
    run("at", lambda a, b: C.at_criterion(a, y, b, tf, 1000))
    for k in ("cosine", "poly", "l2", "rbf"):
        run(f"gpw_{k}", lambda a, b, k=k: C.gpw_criterion(a, y, b, tf, k, 1.0, 10 ** 9))
        run(f"lpw_{k}", lambda a, b, k=k: C.lpw_criterion(a, y, b, tf, sub, k, 100))
    run("gpw_cosine_sampled", lambda a, b: C.gpw_criterion(a, y, b, tf, "cosine", 1.0, 64, draw))
    run("nce_sampled", lambda a, b: C.nce_criterion(a, y, b, same, 0.5, 0.075, 64, draw))
    run("nce_full", lambda a, b: C.nce_criterion(a, y, b, same, 0.5, 0.075, 10 ** 9))
