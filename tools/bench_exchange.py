"""Times the exchange primitives of the hybrid-layout engine on N GPUs (torchrun): R->C, C->R, row all-gather, small vector
all-gather and the bare barrier, peer-memory stores vs torch.distributed (NCCL).  CUDA events, max over ranks.
   torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tools/bench_exchange.py"""
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import efficient_gnns_b200  # noqa: E402,F401
from efficient_gnns_b200.hybrid import DensePlan, PeerExchange, TorchExchange, make_dense_plan  # noqa: E402


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", os.environ["RANK"])), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    N = 169_343
    plan = make_dense_plan(torch.ones(N, dtype=torch.long), world)
    n_p, B = plan.counts[rank], plan.block
    for mode in ("peer", "nccl"):
        ex = PeerExchange(plan, rank, 4 * (N * 256 // world * 4 + B * 256 * 4 + N * 40 * 4 + world * 1024 * 4) + (1 << 20)) if mode == "peer" \
            else TorchExchange(plan, rank)
        for K in (256, 128):
            if K % (4 * world):
                continue
            src_r = torch.randn(n_p, K, device=dev)
            dst_c = ex.buffer(f"c{K}", (N, K // world), dev)
            src_c = torch.randn(N, K // world, device=dev)
            dst_r = ex.buffer(f"r{K}", (B, K), dev)[:n_p]
            t1 = timed(lambda: ex.r2c(src_r, dst_c, f"c{K}"))
            t2 = timed(lambda: ex.c2r(src_c, dst_r, f"r{K}"))
            byt = n_p * K * 4 * (world - 1) // world
            if rank == 0:
                print(json.dumps(dict(mode=mode, world=world, op="r2c", K=K, ms=t1, MB_sent_per_rank=byt / 1e6, GBps=byt / t1 / 1e6)))
                print(json.dumps(dict(mode=mode, world=world, op="c2r", K=K, ms=t2, MB_sent_per_rank=byt / 1e6, GBps=byt / t2 / 1e6)), flush=True)
        src = torch.randn(n_p, 40, device=dev)
        full = ex.buffer("full40", (N, 40), dev)
        t3 = timed(lambda: ex.allgather_rows(src, full, "full40"))
        vec = torch.randn(1024, device=dev)
        allv = ex.buffer("vec", (world, 1024), dev)
        t4 = timed(lambda: ex.allgather_vec(vec, allv, "vec"))
        if rank == 0:
            print(json.dumps(dict(mode=mode, world=world, op="allgather_rows", K=40, ms=t3, MB_recv_per_rank=(N - n_p) * 160 / 1e6)))
            print(json.dumps(dict(mode=mode, world=world, op="allgather_vec", floats=1024, ms=t4)), flush=True)
        if mode == "peer":
            t5 = timed(lambda: ex.arena.barrier())
            if rank == 0:
                print(json.dumps(dict(mode=mode, world=world, op="barrier", ms=t5)), flush=True)
            ex.check()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
