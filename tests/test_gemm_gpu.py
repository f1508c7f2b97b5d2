"""tcgen05 3xTF32 GEMM (through the C ABI) vs fp64: fp32-level accuracy on the tensor cores."""
import pytest
import torch

import efficient_gnns_b200  # noqa: F401
from conftest import rel_err
from efficient_gnns_b200 import ops

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(120)]


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (128, 128, 128), (1000, 256, 128), (5000, 256, 256), (777, 40, 256),
                                   (3000, 256, 40), (129, 130, 36), (20_000, 128, 750 // 2 * 2 + 2)])
@pytest.mark.parametrize("bias", [False, True])
def test_gemm_matches_fp64(M, N, K, bias):
    K = (K + 3) // 4 * 4
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g) if bias else None
    ref = a.double() @ w.double().t() + (b.double() if bias else 0)
    hi, lo = ops.split_tf32(w.cuda())
    assert torch.equal((hi.cpu().view(torch.int32) & 0x1FFF), torch.zeros(N, K, dtype=torch.int32))
    out = ops.gemm_tf32x3(a.cuda(), hi, lo, b.cuda() if bias else None)
    torch.cuda.synchronize()
    e = rel_err(out, ref)
    e32 = rel_err(a @ w.t() + (b if bias else 0), ref)      # what plain fp32 achieves on the CPU
    assert e < 1e-5, (e, e32)
    assert e < 20 * max(e32, 1e-7), (e, e32)


def test_split_transpose():
    w = torch.randn(96, 200, generator=torch.Generator().manual_seed(0))
    hi, lo = ops.split_tf32(w.cuda(), transpose=True)
    assert hi.shape == (200, 96)
    assert (hi.cpu().double() + lo.cpu().double() - w.t().double()).abs().max().item() < 2.0 ** -21 * w.abs().max().item()
    h2, l2 = ops.split_tf32(w.cuda())
    assert torch.equal(h2.t().contiguous(), hi)


def test_gemm_deterministic_and_reusable():
    g = torch.Generator().manual_seed(1)
    a, w = torch.randn(4000, 256, generator=g).cuda(), torch.randn(256, 256, generator=g).cuda()
    hi, lo = ops.split_tf32(w)
    o1 = ops.gemm_tf32x3(a, hi, lo)
    o2 = ops.gemm_tf32x3(a, hi, lo)
    assert torch.equal(o1, o2)


@pytest.mark.parametrize("Nn,Kin,Nout", [(16, 128, 32), (1000, 128, 256), (5003, 256, 256), (40_000, 256, 64), (7, 128, 128),
                                          (3000, 256, 40), (999, 128, 4), (2000, 256, 100)])
def test_wgrad_matches_fp64(Nn, Kin, Nout):
    g = torch.Generator().manual_seed(Nn + Kin + Nout)
    x = torch.randn(Nn, Kin, generator=g)
    d = torch.randn(Nn, Nout, generator=g)
    ref = x.double().t() @ d.double()
    out = ops.gemm_wgrad_tf32x3(x.cuda(), d.cuda())
    torch.cuda.synchronize()
    e, e32 = rel_err(out, ref), rel_err(x.t() @ d, ref)
    assert e < 1e-5, (e, e32)
    out2 = ops.gemm_wgrad_tf32x3(x.cuda(), d.cuda())
    assert torch.equal(out, out2)


def test_wgrad_unsupported_shapes_are_reported():
    from efficient_gnns_b200 import lib
    assert not ops.wgrad_supported(64, 40) and ops.wgrad_supported(128, 256) and ops.wgrad_supported(256, 40)
    with pytest.raises(lib.B200GnnError):
        ops.gemm_wgrad_tf32x3(torch.randn(100, 64, device="cuda"), torch.randn(100, 40, device="cuda"))
