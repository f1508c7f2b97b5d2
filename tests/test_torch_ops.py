"""torch.ops.b200gnn.* (the dispatcher binding of the C ABI): registration, meta/fake tracing on CPU; values,
gradients and torch.library.opcheck on the GPU."""
import pytest
import torch

import efficient_gnns_b200  # noqa: F401
from efficient_gnns_b200 import lib, torch_ops
from efficient_gnns_b200.synthetic import skewed_edges


def csr(n=500, e=4000, seed=0, device="cpu", weighted=True):
    ei = skewed_edges(n, e, seed)
    key = torch.unique(ei[1] * n + ei[0])
    row, col = key // n, key % n
    rowptr = torch.zeros(n + 1, dtype=torch.long)
    rowptr[1:] = torch.bincount(row, minlength=n).cumsum(0)
    val = torch.rand(col.numel(), generator=torch.Generator().manual_seed(seed)) if weighted else None
    return rowptr.to(device), col.to(device), None if val is None else val.to(device)


def test_ops_are_registered_with_schemas():
    for name in torch_ops.OPS:
        assert hasattr(torch.ops.b200gnn, name)
    s = str(torch.ops.b200gnn.spmm_sum.default._schema)
    assert "Tensor rowptr" in s and "Tensor? value" in s and "Tensor mat" in s
    assert "Tensor? bias" in str(torch.ops.b200gnn.gemm_tf32x3.default._schema)


def test_fake_tensor_tracing_on_cpu():
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        rowptr, col = torch.empty(11, dtype=torch.long), torch.empty(50, dtype=torch.long)
        mat = torch.empty(10, 24)
        assert torch.ops.b200gnn.spmm_sum(rowptr, col, None, mat).shape == (10, 24)
        assert torch.ops.b200gnn.spmm_mean(rowptr, col, torch.empty(50), mat).shape == (10, 24)
        hi, lo = torch.ops.b200gnn.split_tf32(torch.empty(24, 16), True)
        assert hi.shape == lo.shape == (16, 24)
        assert torch.ops.b200gnn.gemm_tf32x3(mat, torch.empty(16, 24), torch.empty(16, 24), None).shape == (10, 16)
        assert torch.ops.b200gnn.ind2ptr(torch.empty(50, dtype=torch.long), 10).shape == (11,)


def test_cpu_tensors_raise_no_fallback():
    rowptr, col, val = csr(50, 200)
    with pytest.raises(lib.B200GnnError):
        torch.ops.b200gnn.spmm_sum(rowptr, col, val, torch.randn(50, 8))


@pytest.mark.gpu
@pytest.mark.parametrize("weighted", [True, False])
@pytest.mark.parametrize("op,reduce", [("spmm_sum", "sum"), ("spmm_mean", "mean")])
def test_spmm_ops_values_and_gradients(op, reduce, weighted):
    n, K = 3000, 64
    rowptr, col, val = csr(n, 30_000, 1, "cuda", weighted)
    x = torch.randn(n, K, device="cuda", requires_grad=True)
    w = torch.randn(n, K, device="cuda")
    y = getattr(torch.ops.b200gnn, op)(rowptr, col, val, x)
    (y * w).sum().backward()
    # fp64 reference from the same CSR
    row = torch.repeat_interleave(torch.arange(n, device="cuda"), rowptr[1:] - rowptr[:-1])
    v = (val if weighted else torch.ones(col.numel(), device="cuda")).double()
    xd = x.detach().double().requires_grad_(True)
    ref = torch.zeros(n, K, dtype=torch.float64, device="cuda").index_add_(0, row, xd[col] * v[:, None])
    if reduce == "mean":
        ref = ref / (rowptr[1:] - rowptr[:-1]).clamp(min=1).double()[:, None]
    (ref * w.double()).sum().backward()
    assert (y.double() - ref).abs().max().item() < 1e-5 * ref.abs().max().item()
    assert (x.grad.double() - xd.grad).abs().max().item() < 1e-5 * xd.grad.abs().max().item()
    # ... and the plan cache serves the second call (same tensors) and notices an in-place change of the graph
    n_cached = len(torch_ops._CACHE)
    getattr(torch.ops.b200gnn, op)(rowptr, col, val, x.detach())
    assert len(torch_ops._CACHE) == n_cached
    if weighted:
        val.mul_(2.0)
        y2 = getattr(torch.ops.b200gnn, op)(rowptr, col, val, x.detach())
        assert (y2 - 2 * y.detach()).abs().max().item() < 1e-5 * y.abs().max().item()


@pytest.mark.gpu
def test_dense_ops_and_opcheck():
    a = torch.randn(300, 64, device="cuda")
    w = torch.randn(48, 64, device="cuda") / 8
    hi, lo = torch.ops.b200gnn.split_tf32(w, False)
    out = torch.ops.b200gnn.gemm_tf32x3(a, hi, lo, None)
    ref = a.double() @ w.double().t()
    assert (out.double() - ref).abs().max().item() < 1e-5 * ref.abs().max().item()
    rowptr, col, val = csr(200, 1500, 2, "cuda")
    assert torch.equal(torch.ops.b200gnn.ind2ptr(torch.ops.b200gnn.ptr2ind(rowptr, col.numel()), 200), rowptr)
    x = torch.randn(200, 32, device="cuda", requires_grad=True)
    torch.library.opcheck(torch.ops.b200gnn.spmm_sum.default, (rowptr, col, val, x),
                          test_utils=("test_schema", "test_faketensor", "test_autograd_registration"))
    torch.library.opcheck(torch.ops.b200gnn.gemm_tf32x3.default, (a, hi, lo, None), test_utils=("test_schema", "test_faketensor"))


@pytest.mark.gpu
def test_ops_trace_under_torch_compile():
    """The binding is visible to the compiler stack (the hot path itself never runs under torch.compile)."""
    rowptr, col, val = csr(400, 3000, 3, "cuda")
    x = torch.randn(400, 16, device="cuda")

    def f(m):
        return torch.relu(torch.ops.b200gnn.spmm_sum(rowptr, col, val, m)) * 2

    out = torch.compile(f, backend="aot_eager", fullgraph=True)(x)
    assert torch.equal(out, f(x))
