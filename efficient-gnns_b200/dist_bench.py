"""bench.py's N>1 arm: the same GCN + logit-KD training step over N GPUs (torchrun, one rank per GPU).

Default engine: hybrid.HybridGCNTrainer with peer-memory exchanges (B200GNN_DIST_MODE=hybrid-peer); other modes for A/B
runs: hybrid-nccl (same layout, torch.distributed all-to-all) and allgather (round-1 node-parallel engine, dist.py)."""
from __future__ import annotations

import json
import os

import torch
import torch.distributed as dist


def run(args):
    import bench as B
    from . import lib, sparse, synthetic
    from .dist import ShardedGCNTrainer
    from .hybrid import HybridGCNTrainer

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    lib.load()
    ds = synthetic.make_node_dataset(synthetic.ARXIV, seed=0)
    n = ds.num_nodes
    ei = ds.edge_index.to(dev)
    perm = (ei[1] * n + ei[0]).argsort()
    adj = sparse.SparseTensor(row=ei[1][perm], col=ei[0][perm], sparse_sizes=(n, n), is_sorted=True).to_symmetric()
    mode = os.environ.get("B200GNN_DIST_MODE", "hybrid-peer")
    if mode.startswith("hybrid"):
        try:
            tr = HybridGCNTrainer(adj, B.DIMS, dropout=0.5, lr=0.01, seed=0, exchange="peer" if mode == "hybrid-peer" else "nccl")
        except Exception as e:  # noqa: BLE001  (e.g. CUDA IPC unavailable in this container): same layout over NCCL
            if mode != "hybrid-peer":
                raise
            if rank == 0:
                print(f"[dist_bench] peer exchange unavailable ({type(e).__name__}: {e}); falling back to hybrid-nccl", flush=True)
            mode = "hybrid-nccl"
            tr = HybridGCNTrainer(adj, B.DIMS, dropout=0.5, lr=0.01, seed=0, exchange="nccl")
        nnz = tr.Gfull.nnz
    else:
        tr = ShardedGCNTrainer(adj, B.DIMS, dropout=0.5, lr=0.01, seed=0)
        nnz_global = torch.tensor([tr.nnz], device=dev, dtype=torch.long)
        dist.all_reduce(nnz_global)
        nnz = int(nnz_global.item())
    x_pad, y_loc, tr_loc, t_loc = tr.shard_inputs(ds.x, ds.y.squeeze(1), ds.split_idx["train"], ds.teacher_logits)

    graph = None
    for _ in range(2):
        tr.train_step(x_pad, y_loc, tr_loc, t_loc)
    torch.cuda.synchronize()
    if not args.no_graph:
        try:   # NCCL collectives are capturable; fall back to eager launches if this build refuses
            tr.capture(x_pad, y_loc, tr_loc, t_loc, warmup=1)
            graph = True
        except Exception as e:  # noqa: BLE001
            if rank == 0:
                print(f"[dist_bench] CUDA-graph capture failed ({type(e).__name__}: {e}); running eager", flush=True)
            graph = None
            torch.cuda.synchronize()
    step = tr.replay if graph else (lambda: tr.train_step(x_pad, y_loc, tr_loc, t_loc))
    lib.reset_launch_count()
    tr.train_step(x_pad, y_loc, tr_loc, t_loc)
    launches = lib.launch_count()

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with B.ClockSampler(local) as clk:
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
    dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item())

    # end to end: every step re-uploads this rank's inputs from pinned host memory and reads the loss back; two input
    # sets + two captured graphs so that the upload of step k+1 overlaps step k (as in the 1-GPU arm)
    host = {"x": x_pad.cpu().pin_memory(), "y": y_loc.cpu().pin_memory(), "t": t_loc.cpu().pin_memory(),
            "i": tr_loc.cpu().pin_memory()}
    sets = [{"x": x_pad, "y": y_loc, "t": t_loc, "i": tr_loc}]
    loss_host = torch.empty(3).pin_memory()
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    overlap = False
    if graph:
        try:
            s2 = {k: v.clone() for k, v in sets[0].items()}
            tr.capture(s2["x"], s2["y"], s2["i"], s2["t"], warmup=1, key=1)
            sets.append(s2)
            overlap = True
        except Exception:  # noqa: BLE001
            torch.cuda.synchronize()
    copy_stream, main = torch.cuda.Stream(), torch.cuda.current_stream()
    uploaded = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def upload(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i])
            for k in sets[i]:
                sets[i][k].copy_(host[k], non_blocking=True)
            uploaded[i].record(copy_stream)

    def e2e_loop(n_steps):
        if not overlap:
            for _ in range(n_steps):
                for k in sets[0]:
                    sets[0][k].copy_(host[k], non_blocking=True)
                step()
                loss_host.copy_(tr.loss_out, non_blocking=True)
            return
        for i in (0, 1):
            consumed[i].record(main)
        upload(0)
        for it in range(n_steps):
            i = it & 1
            if it + 1 < n_steps:
                upload(1 - i)
            main.wait_event(uploaded[i])
            tr.replay(i)
            consumed[i].record(main)
            loss_host.copy_(tr.loss_out, non_blocking=True)

    e2e_loop(4)
    torch.cuda.synchronize(); dist.barrier()
    e0.record()
    e2e_loop(args.steps)
    e1.record(); torch.cuda.synchronize(); dist.barrier()
    t2 = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    ms_e2e = float(t2.item())

    if rank == 0:
        peak, peak_src = B.peaks()
        ex = tr.exchange_bytes_per_step()
        line = {"metric": B.METRIC, "value": 6 * nnz / (ms_step * 1e-3), "unit": B.UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": B.workload_config(ds, nnz),
                "engine": {
                    "parallelism": (f"hybrid layout x{world}: node-parallel dense ops, feature-parallel wide aggregations over the "
                                    f"replicated graph, R<->C exchanges by {'peer-memory stores + flag barrier' if mode == 'hybrid-peer' else 'NCCL all-to-all'}; "
                                    "narrow (40-wide) aggregations row-sharded with an all-gather") if mode.startswith("hybrid")
                    else f"node-parallel x{world}: degree-balanced row blocks, one NCCL all-gather per aggregation",
                    "dist_mode": mode,
                    "cuda_graph": bool(graph), "exchange_bytes_received_per_rank_per_step": ex,
                    "nvlink_floor_ms": ex / 770e9 * 1e3,
                    "aggregations_executed": tr.aggregations_per_step()},
                "roofline": {"bound": "nvlink+hbm",
                             "note": ("multi-GPU point. achieved = bytes each rank RECEIVES over NVLink per step / step time against the "
                                      "measured 770 GB/s peer bandwidth (nvlink_floor_ms = the same bytes at that rate). For the hybrid "
                                      "layout the exchanges are stores issued by the producing kernels' epilogues and the floor is far "
                                      "below the step: the limiter is the per-rank compute that does not shrink with N (every rank "
                                      "walks all edges at width K/N; profiles/r2_hybrid_rank_w*_summary.txt), not the links")
                             if mode.startswith("hybrid") else
                             "multi-GPU point: the all-gathers bound the step; see nvlink_floor_ms",
                             "achieved": ex / (ms_step * 1e-3) / 1e9, "peak": 770.0, "unit": "GB/s",
                             "frac": ex / (ms_step * 1e-3) / 1e9 / 770.0, "traffic": None},
                "cpu_baseline": None,
                "e2e": {"value": 6 * nnz / (ms_e2e * 1e-3), "unit": B.UNIT, "ms_per_step": ms_e2e,
                        "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": 12 * world},
                "gpu_launches": launches * args.steps * world, "gpu_launches_per_step_per_rank": launches,
                "clocks": clk.summary(), "loss": tr.loss_out.tolist()}
        B.emit_json_line(line)
    # NCCL teardown with live CUDA graphs that captured collectives can dead-lock; results are out, leave hard.
    torch.cuda.synchronize()
    import sys
    sys.stdout.flush()
    os._exit(0)
