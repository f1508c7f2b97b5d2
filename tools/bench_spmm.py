"""Micro-benchmark of the SpMM kernel alone on the ARXIV-shape graph (CUDA events, L2 flushed between runs).
Not the contract benchmark (that is bench.py); used to tune and to feed profiles/."""
import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import efficient_gnns_b200  # noqa: E402,F401
from efficient_gnns_b200 import ops, sparse, synthetic  # noqa: E402


def build_adj(n, e, device="cuda", gcn=True, p_local=0.0):
    ei = synthetic.skewed_edges(n, e, 0, p_local).to(device)
    row, col = ei
    perm = (col * n + row).argsort()
    adj = sparse.SparseTensor(row=col[perm], col=row[perm], sparse_sizes=(n, n), is_sorted=True).to_symmetric()
    if gcn:
        adj = adj.fill_value(1.0).fill_diag(1.0)
        deg = adj.sum(1)
        dis = deg.pow(-0.5)
        dis[torch.isinf(dis)] = 0
        r, c, v = adj.coo()
        adj = adj.set_value(dis[r] * v * dis[c])
    return adj


def time_fn(fn, iters, flush):
    evs = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--widths", type=int, nargs="+", default=[256, 128, 40])
    ap.add_argument("--thresholds", type=int, nargs="+", default=[256])
    ap.add_argument("--chunks", type=int, nargs="+", default=[128])
    ap.add_argument("--variants", type=int, nargs="+", default=[0])
    ap.add_argument("--out", default="")
    ap.add_argument("--p-local", type=float, default=0.0, help="fraction of edges redrawn inside the source's id block (locality)")
    ap.add_argument("--weighted-only", action="store_true")
    a = ap.parse_args()
    S = synthetic.ARXIV
    n = S["num_nodes"]
    res = []
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    adjs = {gcn: build_adj(n, S["num_edges"], gcn=gcn, p_local=a.p_local) for gcn in ((True,) if a.weighted_only else (True, False))}
    for variant in a.variants:
     ops.set_spmm_variant(variant)
     for thr in a.thresholds:
      for chunk in a.chunks:
        for gcn in adjs:
            st = adjs[gcn].storage
            G = st.engine_csr() if gcn else st.engine_csr_unweighted()
            G.build_plan(hub_threshold=thr, seg_len=thr, chunk_nnz=chunk)
            nnz = G.nnz
            for K in a.widths:
                x = torch.randn(n, K, device="cuda")
                out = torch.empty(n, K, device="cuda")
                red = "sum" if gcn else "mean"
                for _ in range(3):
                    ops.spmm_csr(G, x, red, out=out)
                med, best = time_fn(lambda: ops.spmm_csr(G, x, red, out=out), a.iters, flush)
                alg = 2 * n * K * 4 + nnz * (4 + (4 if gcn else 0)) + (n + 1) * 4
                gather = nnz * (K * 4 + 8) + n * K * 4
                rec = dict(kernel="spmm", p_local=a.p_local, variant=variant, weighted=gcn, reduce=red, K=K, hub_threshold=thr, chunk_nnz=chunk, n_chunks=G.n_chunks, n_hub=G.n_hub, n_seg=G.n_seg,
                           nnz=nnz, ms_median=med, ms_best=best, alg_GBps=alg / med / 1e6, gather_GBps=gather / med / 1e6,
                           edges_per_s=nnz / med * 1e3)
                print(json.dumps(rec), flush=True)
                res.append(rec)
    if a.out:
        Path(a.out).write_text("\n".join(json.dumps(r) for r in res) + "\n")


if __name__ == "__main__":
    main()
