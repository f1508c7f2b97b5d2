"""Head-parallel multi-GPU GAT layer vs the 1-GPU DGL-style GATConv (run under torchrun; spawned by tests/test_multigpu_gpu.py):
   torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tests/hybrid_gat_equiv.py [peer|nccl]
Forward output and every gradient (fc, attn_l, attn_r, res_fc, input features) of one GATConv layer, H=8, D=32 (BASELINE.json
configs[3]) with the symmetric degree normalisation and the residual of the reference teacher, must equal the single-GPU layer
to 1e-5 / 5e-5; and the layer forward+backward is timed at the ARXIV shape."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import efficient_gnns_b200  # noqa: E402,F401
from efficient_gnns_b200 import nn as bnn, sparse, synthetic  # noqa: E402
from efficient_gnns_b200.hybrid import PeerExchange, TorchExchange, make_dense_plan, relabel  # noqa: E402
from efficient_gnns_b200.hybrid_gat import HeadParallelGATConv  # noqa: E402


def rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def build_graph(n, e, dev):
    ei = synthetic.skewed_edges(n, e, 0).to(dev)
    perm = (ei[1] * n + ei[0]).argsort()
    adj = sparse.SparseTensor(row=ei[1][perm], col=ei[0][perm], sparse_sizes=(n, n), is_sorted=True).to_symmetric()
    r, c, _ = adj.coo()                                              # + self loops (arxiv_dgl/gat.py:61,66)
    off = r != c
    d = torch.arange(n, device=dev)
    return sparse.SparseTensor(row=torch.cat([r[off], d]), col=torch.cat([c[off], d]), sparse_sizes=(n, n), is_sorted=False)


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "peer"
    rank, local, world = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", os.environ["RANK"])), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    Fin, H, D = 128, 8, 32
    for n, e, timed in ((20_011, 150_000, False), (169_343, 1_166_243, True)):
        adj = build_graph(n, e, dev)
        plan = make_dense_plan(adj.storage.rowcount(), world)
        adj_rel = relabel(adj, plan)
        K = H * D
        arena = 4 * (n * (K // world) * 4 * 2 + plan.block * K * 4 * 2) + (1 << 20)
        ex = PeerExchange(plan, rank, arena) if mode == "peer" else TorchExchange(plan, rank)
        torch.manual_seed(0)
        ref = bnn.DGLGATConv(Fin, D, num_heads=H, residual=True, use_symmetric_norm=True).to(dev)
        par = HeadParallelGATConv(Fin, D, H, plan, rank, ex, residual=True, use_symmetric_norm=True).to(dev)
        par.load_state_dict(ref.state_dict())
        g = torch.Generator().manual_seed(1)
        x = torch.randn(n, Fin, generator=g).to(dev)
        w = torch.randn(n, H, D, generator=g).to(dev)
        r0, r1 = plan.rows_of(rank)
        mine = plan.perm.to(dev)[r0:r1]
        if not timed:
            xr = x.clone().requires_grad_(True)
            out_ref = ref(adj, xr)
            (out_ref * w).sum().backward()
            xp = x[mine].clone().requires_grad_(True)
            out = par(adj_rel, xp)
            (out * w[mine]).sum().backward()
            par.allreduce_grads()
            e_out = rel(out, out_ref[mine])
            e_x = rel(xp.grad, xr.grad[mine])
            e_par = max(rel(pp.grad, pr.grad) for (_, pp), (_, pr) in zip(par.named_parameters(), ref.named_parameters()))
            if rank == 0:
                print(f"[gat {mode} P={world}] out {e_out:.2e}  d_feat {e_x:.2e}  param grads {e_par:.2e}", flush=True)
            ok &= e_out < 1e-5 and e_x < 5e-5 and e_par < 5e-5
        else:
            xp = x[mine].clone().requires_grad_(True)

            def step():
                for p_ in par.parameters():
                    p_.grad = None
                xp.grad = None
                o = par(adj_rel, xp)
                (o * w[mine]).sum().backward()
            for _ in range(3):
                step()
            torch.cuda.synchronize(); dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                step()
            e1.record(); torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / 10], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)

            xr = x.clone().requires_grad_(True)

            def step1():
                for p_ in ref.parameters():
                    p_.grad = None
                xr.grad = None
                o = ref(adj, xr)
                (o * w).sum().backward()
            for _ in range(3):
                step1()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(10):
                step1()
            e1.record(); torch.cuda.synchronize()
            if rank == 0:
                print(f"[gat {mode} P={world}] ARXIV-shape layer fwd+bwd (H=8, D=32): {float(t):.3f} ms on {world} GPUs, "
                      f"{e0.elapsed_time(e1) / 10:.3f} ms on 1 GPU (autograd module path)", flush=True)
        if mode == "peer":
            ex.check()
        del ref, par
        torch.cuda.synchronize()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("HYBRID_GAT_EQUIV", mode, f"P={world}", "PASS" if flag.item() == 1 else "FAIL", flush=True)
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
