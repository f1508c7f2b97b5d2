"""Node-parallel host logic on CPU with a real world_size-2 gloo group: partition plan, relabelling, row shards,
padding and the all-gather assembly must reproduce the full (single-process) oracle aggregation."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import efficient_gnns_b200  # noqa: F401
from efficient_gnns_b200 import dist as D
from efficient_gnns_b200.sparse import SparseTensor
from efficient_gnns_b200.synthetic import skewed_edges
from oracle import graph as og, ops as oo


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, e, K, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ei = skewed_edges(n, e, 0).numpy()
    row, col, _ = og.to_sparse_adj_t(ei, n)
    r, c = og.to_symmetric(row, col, n)
    r, c, v = og.gcn_norm(r, c, n)
    r, c, v = map(torch.from_numpy, (r, c, v))
    adj = SparseTensor(row=r, col=c, value=v, sparse_sizes=(n, n), is_sorted=True)
    plan = D.make_plan(adj.storage.rowcount(), world)
    rel = D.relabel_adjacency(adj, plan)
    rowptr, colx, val = D.shard_rows(rel, plan, rank)
    n_real = plan.real_rows(rank)
    g = torch.Generator().manual_seed(5)
    h = torch.randn(n, K, generator=g, dtype=torch.float64)        # same on every rank (original order)
    h_pad = D.scatter_rows(h, plan)
    r0, r1 = plan.rows_of(rank)
    blk = torch.zeros(plan.block, K, dtype=torch.float64)
    blk[:n_real] = h_pad[r0:r1]
    full = torch.empty(plan.n_pad, K, dtype=torch.float64)
    dist.all_gather_into_tensor(full, blk)                          # the per-layer exchange
    assert torch.equal(full, h_pad)
    y_loc = oo.spmm_csr(rowptr, colx, val.double(), full, n_real, "sum")
    out_blk = torch.zeros(plan.block, K, dtype=torch.float64); out_blk[:n_real] = y_loc
    y_full = torch.empty(plan.n_pad, K, dtype=torch.float64)
    dist.all_gather_into_tensor(y_full, out_blk)
    y = y_full[plan.inv]
    ref = oo.spmm_scatter(r, c, v.double(), h, n, "sum")
    err = (y - ref).abs().max().item() / ref.abs().max().item()
    # nnz balance and bookkeeping
    nnz = torch.tensor([colx.numel()]); lst = [torch.zeros(1, dtype=torch.long) for _ in range(world)]
    dist.all_gather(lst, nnz)
    out_q.put((rank, err, [int(t) for t in lst], n_real))
    dist.destroy_process_group()


@pytest.mark.parametrize("n,e", [(1001, 6000), (4096, 30_000)])
def test_sharded_aggregation_matches_full_oracle_gloo(n, e):
    world, K = 2, 12
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, e, K, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, nnzs, n_real in res:
        assert err < 1e-12
        assert abs(nnzs[0] - nnzs[1]) <= 0.10 * max(nnzs) + 64      # degree-balanced blocks (one hub apart at most)
    assert sum(r[3] for r in res) == n


def test_plan_properties():
    rc = torch.randint(0, 100, (1003,), generator=torch.Generator().manual_seed(0))
    for world in (1, 2, 4, 8):
        plan = D.make_plan(rc, world)
        assert torch.equal(plan.perm[plan.inv], torch.arange(1003))
        assert sum(plan.real_rows(r) for r in range(world)) == 1003
        for r in range(world):
            blk = plan.perm[r * plan.block:(r + 1) * plan.block]
            k = plan.real_rows(r)
            assert (blk[:k] >= 0).all() and (blk[k:] < 0).all()
        loads = [int(rc[plan.perm[r * plan.block:r * plan.block + plan.real_rows(r)]].sum()) for r in range(world)]
        assert max(loads) - min(loads) <= 100


@pytest.mark.parametrize("world", [3, 4, 8])
def test_all_ranks_simulated_in_one_process_match_full_oracle(world):
    """The same shard/gather/aggregate pipeline as the gloo test, with every rank's part computed in this process:
    covers the world sizes the scaling run uses (padding rows appear when world does not divide N)."""
    n, e, K = 1003, 7000, 8
    ei = skewed_edges(n, e, 1).numpy()
    row, col, _ = og.to_sparse_adj_t(ei, n)
    r, c = og.to_symmetric(row, col, n)
    r, c, v = map(torch.from_numpy, og.gcn_norm(r, c, n))
    adj = SparseTensor(row=r, col=c, value=v, sparse_sizes=(n, n), is_sorted=True)
    plan = D.make_plan(adj.storage.rowcount(), world)
    rel = D.relabel_adjacency(adj, plan)
    h = torch.randn(n, K, generator=torch.Generator().manual_seed(2), dtype=torch.float64)
    full = D.scatter_rows(h, plan)                       # what the all-gather of every rank's block assembles
    y_pad = torch.zeros(plan.n_pad, K, dtype=torch.float64)
    nnz = []
    for rank in range(world):
        rowptr, colx, val = D.shard_rows(rel, plan, rank)
        r0, r1 = plan.rows_of(rank)
        assert rowptr.numel() == plan.real_rows(rank) + 1 and r1 - r0 == plan.real_rows(rank) <= plan.block
        y_pad[r0:r1] = oo.spmm_csr(rowptr, colx, val.double(), full, plan.real_rows(rank), "sum")
        nnz.append(int(colx.numel()))
    ref = oo.spmm_scatter(r, c, v.double(), h, n, "sum")
    assert (y_pad[plan.inv] - ref).abs().max().item() < 1e-12 * max(1.0, ref.abs().max().item())
    assert sum(nnz) == r.numel() and max(nnz) <= 1.25 * sum(nnz) / world    # degree-balanced (one hub dominates at this size)
    pad_rows = torch.ones(plan.n_pad, dtype=torch.bool); pad_rows[plan.inv] = False
    assert int(pad_rows.sum()) == plan.n_pad - n and torch.all(full[pad_rows] == 0)
