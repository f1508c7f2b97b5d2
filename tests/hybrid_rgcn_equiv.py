"""Feature-parallel multi-GPU R-GCN inference (efficient_gnns_b200/rgcn.py) vs the single-GPU engine (torchrun; spawned by
tests/test_multigpu_gpu.py):  torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tests/hybrid_rgcn_equiv.py [peer|nccl]
A 3-type, 5-relation synthetic heterogeneous graph (the structure of the MAG fixture, 40x larger), 2 layers 64 -> 64 -> 16."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import efficient_gnns_b200  # noqa: E402,F401
from efficient_gnns_b200.hybrid import PeerExchange, TorchExchange  # noqa: E402
from efficient_gnns_b200.rgcn import RGCNInference  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "peer"
    rank, local, world = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", os.environ["RANK"])), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    g = torch.Generator().manual_seed(0)
    num_nodes = {0: 6001, 1: 4999, 2: 803}
    rel_list = [(1, "r0", 2, 9000), (2, "r1", 1, 9000), (1, "r2", 0, 40_000), (0, "r3", 1, 40_000), (0, "r4", 0, 50_000)]
    eid, key2int = {}, {0: 0, 1: 1, 2: 2}
    for i, (s, name, d, e) in enumerate(rel_list):
        src = torch.randint(0, num_nodes[s], (e,), generator=g)
        dst = (torch.rand(e, generator=g) ** 3 * num_nodes[d]).long().clamp_(max=num_nodes[d] - 1)     # skewed: hub destinations
        eid[(s, name, d)] = torch.stack([src, dst])
        key2int[(s, name, d)] = i
    F_in, F_h, F_out = 64, 64, 16
    state = {"emb_dict.1": torch.randn(num_nodes[1], F_in, generator=g) * 0.3, "emb_dict.2": torch.randn(num_nodes[2], F_in, generator=g) * 0.3}
    for i, (a, b) in enumerate(((F_in, F_h), (F_h, F_out))):
        for r in range(5):
            state[f"convs.{i}.rel_lins.{r}.weight"] = torch.randn(b, a, generator=g) * 0.2
        for t in range(3):
            state[f"convs.{i}.root_lins.{t}.weight"] = torch.randn(b, a, generator=g) * 0.2
            state[f"convs.{i}.root_lins.{t}.bias"] = torch.randn(b, generator=g) * 0.1
    x0 = torch.randn(num_nodes[0], F_in, generator=g)
    one = RGCNInference(state, num_nodes, eid, key2int, device=dev)
    ref = one({0: x0})

    def factory(plan):
        if mode == "peer":
            need = 8 * (plan.n * (F_in // world) * 4 + plan.block * F_in * 4) * 6 + (1 << 20)
            return PeerExchange(plan, rank, need)
        return TorchExchange(plan, rank)
    par = RGCNInference(state, num_nodes, eid, key2int, device=dev, exchange_factory=factory, rank=rank, world=world)
    out = par.gather(par({0: x0}))
    err = max(((out[t] - ref[t]).abs().max() / ref[t].abs().max()).item() for t in range(3))
    if mode == "peer":
        for e in par.ex.values():
            e.check()
    ok = err < 1e-5
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"[rgcn {mode} P={world}] max rel err vs single GPU {err:.2e}", flush=True)
        print("HYBRID_RGCN_EQUIV", mode, f"P={world}", "PASS" if flag.item() == 1 else "FAIL", flush=True)
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
