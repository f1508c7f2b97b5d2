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


@pytest.mark.parametrize("kernel", ["cosine", "rbf"])
def test_lsp_backward_is_bitwise_repeatable_and_handles_hubs_and_self_loops(kernel):
    """The LSP gradient is one sparse product with a fixed summation order (b200gnn_lsp_bwd_values_f32 + the row-segmented
    SpMM): two runs give identical bits (the reference's index backward is atomic), a 3000-edge hub destination, duplicate
    edges and self loops included; values against the oracle."""
    n = 4000
    z, y, t, f, tf, _ = data(n=n, Fs=256, Ft=64)
    if kernel == "rbf":
        f, tf = f * 0.05, tf * 0.1
    g = torch.Generator().manual_seed(9)
    src = torch.cat([torch.randint(0, n, (20_000,), generator=g), torch.randint(0, n, (3000,), generator=g), torch.arange(50),
                     torch.tensor([5, 5, 5])])
    dst = torch.cat([torch.randint(0, n, (20_000,), generator=g), torch.full((3000,), 17), torch.arange(50), torch.tensor([9, 9, 9])])
    ei = torch.stack([src, dst])
    grads = []
    for _ in range(2):
        fc = f.cuda().requires_grad_(True)
        out = C.lpw_criterion(z.cuda(), y.cuda(), fc, tf.cuda(), ei.cuda(), kernel, 100, "kld")
        out[0].backward()
        grads.append(fc.grad.clone())
    assert torch.equal(grads[0], grads[1])
    fr = f.double().requires_grad_(True)
    ref = oc.lpw_criterion(z.double(), y, fr, tf.double(), ei, kernel, 100, "kld")
    ref[0].backward()
    assert rel_err(grads[0], fr.grad) < 5e-5


@pytest.mark.parametrize("S,F_", [(901, 64), (1536, 250)])
def test_nce_streams_row_chunks_without_the_full_logits(S, F_, monkeypatch):
    """G-CRD (criterion.py:129-149) with the chunk budget forced small: several row chunks, a ragged last chunk, S and F not
    multiples of 4, student and teacher gradients; and the peak memory stays far below the S x S logits."""
    monkeypatch.setattr(C, "NCE_CHUNK_BYTES", 4 * 256 * ((S + 3) // 4 * 4))        # R = 256 rows per chunk
    z, y, t, f, tf, _ = data(n=S, Fs=F_, Ft=8)
    ts = torch.randn(S, F_, generator=torch.Generator().manual_seed(4))
    check(lambda a, b, c: C.nce_criterion(a, y.cuda(), b, c, 0.5, 0.075, 10 ** 9),
          lambda a, b, c: oc.nce_criterion(a, y, b, c, 0.5, 0.075, 10 ** 9), z, f, ts, tol=5e-5, teacher_grad=True)
    # student-only gradient (teacher detached, as at the reference's call sites) takes the cheaper path
    check(lambda a, b, c: C.nce_criterion(a, y.cuda(), b, c.detach(), 0.5, 0.075, 10 ** 9),
          lambda a, b, c: oc.nce_criterion(a, y, b, c.detach(), 0.5, 0.075, 10 ** 9), z, f, ts, tol=5e-5)


def test_nce_peak_memory_is_linear_in_samples():
    S, F_ = 8192, 256
    g = torch.Generator().manual_seed(0)
    fs = torch.randn(S, F_, generator=g).cuda().requires_grad_(True)
    ft = torch.randn(S, F_, generator=g).cuda()
    z, y = torch.randn(S, 8, generator=g).cuda().requires_grad_(True), torch.randint(0, 8, (S,), generator=g).cuda()
    torch.cuda.synchronize(); torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    out = C.nce_criterion(z, y, fs, ft, 0.5, 0.075, 10 ** 9)
    out[0].backward()
    torch.cuda.synchronize()
    peak = torch.cuda.max_memory_allocated() - base
    assert peak < 0.6 * S * S * 4, (peak, S * S * 4)          # the reference's [S,S] logits alone are S*S*4 = 256 MB
    assert torch.isfinite(fs.grad).all()
