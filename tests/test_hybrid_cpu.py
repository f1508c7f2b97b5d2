"""Hybrid-layout (feature-parallel aggregation / node-parallel dense) host logic on CPU, real world_size-2 and -3 gloo
groups: the dense partition plan, the R<->C layout exchanges and the row all-gather must reproduce the single-process
oracle aggregation exactly, for both aggregation modes (efficient-gnns_b200/hybrid.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import efficient_gnns_b200  # noqa: F401
from efficient_gnns_b200 import hybrid as H
from efficient_gnns_b200.sparse import SparseTensor
from efficient_gnns_b200.synthetic import skewed_edges
from oracle import graph as og, ops as oo


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _graph(n, e):
    ei = skewed_edges(n, e, 0).numpy()
    row, col, _ = og.to_sparse_adj_t(ei, n)
    r, c = og.to_symmetric(row, col, n)
    r, c, v = og.gcn_norm(r, c, n)
    return tuple(map(torch.from_numpy, (r, c, v)))


def _worker(rank, world, port, n, e, K, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    r, c, v = _graph(n, e)
    adj = SparseTensor(row=r, col=c, value=v, sparse_sizes=(n, n), is_sorted=True)
    plan = H.make_dense_plan(adj.storage.rowcount(), world)
    rel = H.relabel(adj, plan)
    ex = H.TorchExchange(plan, rank)
    r0, r1 = plan.rows_of(rank)
    n_p = r1 - r0
    g = torch.Generator().manual_seed(5)
    h = torch.randn(n, K, generator=g)                     # original node order, identical on every rank
    h_rel = h[plan.perm]
    h_R = h_rel[r0:r1].contiguous()
    kc = K // world
    ref = oo.spmm_scatter(r, c, v.double(), h.double(), n, "sum")[plan.perm]      # relabelled order

    # R -> C: my columns of every node
    h_C = torch.empty(n, kc)
    ex.r2c(h_R, h_C)
    ok_r2c = torch.equal(h_C, h_rel[:, rank * kc:(rank + 1) * kc])
    # feature-parallel aggregation with the whole relabelled matrix, then C -> R
    frp, fcol, fval = rel.csr()
    y_C = oo.spmm_csr(frp, fcol, fval.double(), h_C.double(), n, "sum").float()
    y_R = torch.empty(n_p, K)
    ex.c2r(y_C, y_R)
    err_col = (y_R.double() - ref[r0:r1]).abs().max().item() / ref.abs().max().item()
    # round trip
    back = torch.empty(n_p, K)
    ex.c2r(h_C, back)
    ok_round = torch.equal(back, h_R)
    # node-parallel aggregation: all-gather + row shard
    full = torch.empty(n, K)
    ex.allgather_rows(h_R, full)
    ok_gather = torch.equal(full, h_rel)
    rp, colx, val = H.row_shard(rel, plan, rank)
    y_row = oo.spmm_csr(rp, colx, val.double(), full.double(), n_p, "sum")
    err_row = (y_row - ref[r0:r1]).abs().max().item() / ref.abs().max().item()
    vec = torch.arange(8, dtype=torch.float32) + 100 * rank
    allv = torch.empty(world, 8)
    ex.allgather_vec(vec, allv)
    ok_vec = all(torch.equal(allv[q], torch.arange(8, dtype=torch.float32) + 100 * q) for q in range(world))
    out_q.put((rank, ok_r2c, ok_round, ok_gather, ok_vec, err_col, err_row, int(colx.numel()), n_p))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n,e,K", [(2, 1001, 6000, 8), (3, 2050, 16_000, 12)])
def test_hybrid_exchanges_reproduce_full_oracle_gloo(world, n, e, K):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, e, K, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    nnzs = [r[7] for r in res]
    for rank, ok_r2c, ok_round, ok_gather, ok_vec, err_col, err_row, nnz, n_p in res:
        assert ok_r2c and ok_round and ok_gather and ok_vec, (rank, ok_r2c, ok_round, ok_gather, ok_vec)
        assert err_col < 1e-6 and err_row < 1e-12
    assert max(nnzs) - min(nnzs) <= 0.10 * max(nnzs) + 64
    assert sum(r[8] for r in res) == n


def test_dense_plan_properties():
    rc = torch.randint(0, 100, (1003,), generator=torch.Generator().manual_seed(0))
    for world in (1, 2, 3, 4, 8):
        plan = H.make_dense_plan(rc, world)
        assert sum(plan.counts) == 1003 and max(plan.counts) - min(plan.counts) <= 1
        assert plan.offsets[-1] == 1003 and plan.block == max(plan.counts)
        assert torch.equal(plan.perm[plan.inv], torch.arange(1003)) and torch.equal(plan.inv[plan.perm], torch.arange(1003))
        # degree balance: each rank's share of the total degree is within one hub of the mean
        deg_new = rc[plan.perm]
        shares = [int(deg_new[plan.offsets[p]:plan.offsets[p + 1]].sum()) for p in range(world)]
        assert max(shares) - min(shares) <= int(rc.max()) + world
