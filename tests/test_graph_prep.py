"""Integer graph preparation: the product's SparseTensor host logic (torch ops, device-agnostic) must be
bit-exact with the numpy oracle, on CPU here and on the GPU under -m gpu."""
import numpy as np
import pytest
import torch

import efficient_gnns_b200  # noqa: F401
from efficient_gnns_b200.sparse import SparseTensor
from efficient_gnns_b200.synthetic import skewed_edges
from oracle import graph as og

DEVICES = ["cpu", pytest.param("cuda", marks=pytest.mark.gpu)]


def to_sparse_tensor(ei: torch.Tensor, n: int) -> SparseTensor:
    """T.ToSparseTensor(): perm = argsort(col*N+row); SparseTensor(row=col, col=row, sorted)."""
    row, col = ei
    perm = (col * n + row).argsort()
    return SparseTensor(row=col[perm], col=row[perm], sparse_sizes=(n, n), is_sorted=True)


@pytest.mark.parametrize("device", DEVICES)
@pytest.mark.parametrize("n,e,seed", [(50, 120, 0), (1000, 6000, 1), (10_000, 50_000, 2)])
def test_to_sparse_symmetric_coo_bit_exact(device, n, e, seed):
    ei = skewed_edges(n, e, seed)
    adj = to_sparse_tensor(ei.to(device), n)
    r0, c0, ptr0 = og.to_sparse_adj_t(ei.numpy(), n)
    row, col, _ = adj.coo()
    assert np.array_equal(row.cpu().numpy(), r0) and np.array_equal(col.cpu().numpy(), c0)
    assert np.array_equal(adj.storage.rowptr().cpu().numpy(), ptr0)

    sym = adj.to_symmetric()
    r1, c1 = og.to_symmetric(r0, c0, n)
    row, col, _ = sym.coo()
    assert row.dtype == torch.int64
    assert np.array_equal(row.cpu().numpy(), r1) and np.array_equal(col.cpu().numpy(), c1)
    # csr2csc / colptr
    perm = sym.storage.csr2csc().cpu().numpy()
    assert np.array_equal(perm, og.csr2csc(r1, c1))
    assert np.array_equal(sym.storage.colptr().cpu().numpy(), og.ind2ptr(c1[perm], n))
    # transpose of a symmetric pattern is itself
    rt, ct, _ = sym.t().coo()
    assert torch.equal(rt, row) and torch.equal(ct, col)


@pytest.mark.parametrize("device", DEVICES)
def test_fill_diag_replaces_existing_self_loops(device):
    n = 6
    row = torch.tensor([0, 0, 1, 2, 2, 4], device=device)
    col = torch.tensor([0, 3, 2, 2, 5, 1], device=device)
    adj = SparseTensor(row=row, col=col, value=torch.tensor([9., 1., 1., 7., 1., 1.], device=device), sparse_sizes=(n, n))
    r, c, v = adj.fill_diag(1.0).coo()
    r0, c0, v0 = og.fill_diag(row.cpu().numpy(), col.cpu().numpy(), np.array([9., 1., 1., 7., 1., 1.], dtype=np.float32), n)
    assert np.array_equal(r.cpu().numpy(), r0) and np.array_equal(c.cpu().numpy(), c0)
    assert np.array_equal(v.cpu().numpy(), v0)


@pytest.mark.parametrize("device", DEVICES)
def test_coalesce_duplicates_and_empty(device):
    row = torch.tensor([2, 0, 2, 0, 1], device=device)
    col = torch.tensor([1, 3, 1, 3, 1], device=device)
    adj = SparseTensor(row=row, col=col, sparse_sizes=(4, 4)).coalesce()
    r, c, _ = adj.coo()
    assert r.tolist() == [0, 1, 2] and c.tolist() == [3, 1, 1]
    empty = SparseTensor(row=torch.zeros(0, dtype=torch.long, device=device),
                         col=torch.zeros(0, dtype=torch.long, device=device), sparse_sizes=(3, 3))
    assert empty.nnz() == 0 and empty.storage.rowptr().tolist() == [0, 0, 0, 0]
    assert empty.to_symmetric().nnz() == 0


def test_int32_narrowing_rejects_large_indices():
    from efficient_gnns_b200 import lib
    from efficient_gnns_b200.sparse import _narrow_i32
    with pytest.raises(lib.B200GnnError):
        _narrow_i32(torch.tensor([0, 2 ** 31]), "col")


# ---------------------------------------------------------------------------------------------- device ingestion kernels (f2)
@pytest.mark.gpu
@pytest.mark.parametrize("n,major_size,minor_size", [(1, 5, 5), (31, 7, 3), (2049, 100, 100), (100_000, 169_343, 169_343),
                                                     (300_000, 1_939_743, 1_939_743), (50_000, 3, 2 ** 40)])
def test_device_radix_argsort_is_torch_stable_argsort(n, major_size, minor_size):
    """b200gnn_graph_argsort_i64 (hand-written stable LSD radix sort) == torch.argsort(stable=True), duplicates included,
    at tile boundaries, for MAG-sized ids (42-bit keys) and for a key range that needs 6 digit passes."""
    from efficient_gnns_b200.sparse import device_argsort
    g = torch.Generator().manual_seed(n)
    major = torch.randint(0, major_size, (n,), generator=g)
    minor = torch.randint(0, min(minor_size, 2 ** 62), (n,), generator=g)
    major[n // 2:] = major[: n - n // 2].clone()  # plenty of equal majors; equal keys too
    minor[n // 3: n // 3 + n // 4] = minor[: n // 4].clone()
    want = torch.argsort(major * minor_size + minor, stable=True)
    got = device_argsort(major.cuda(), minor.cuda(), major_size, minor_size)
    assert got.dtype == torch.int64 and torch.equal(got.cpu(), want)


@pytest.mark.gpu
@pytest.mark.parametrize("n,e,seed", [(1000, 6000, 1), (169_343, 1_166_243, 0)])
def test_device_coalesce_matches_numpy_oracle(n, e, seed):
    """to_symmetric through b200gnn_graph_coalesce_i64 at the real ARXIV size: rows, columns, row pointers and the source
    index of every kept entry are bit-exact with oracle/graph.py."""
    from efficient_gnns_b200.sparse import device_coalesce
    ei = skewed_edges(n, e, seed)
    r0, c0, _ = og.to_sparse_adj_t(ei.numpy(), n)
    r1, c1 = og.to_symmetric(r0, c0, n)
    rr = torch.cat([torch.from_numpy(r0), torch.from_numpy(c0)])
    cc = torch.cat([torch.from_numpy(c0), torch.from_numpy(r0)])
    ro, co, rowptr, src = device_coalesce(rr.cuda(), cc.cuda(), n, n)
    assert np.array_equal(ro.cpu().numpy(), r1) and np.array_equal(co.cpu().numpy(), c1)
    assert np.array_equal(rowptr.cpu().numpy(), og.ind2ptr(r1, n))
    key = rr * n + cc
    first = {}
    order = torch.argsort(key, stable=True)
    ks = key[order]
    keep = torch.ones_like(ks, dtype=torch.bool); keep[1:] = ks[1:] != ks[:-1]
    assert torch.equal(src.cpu(), order[keep])
