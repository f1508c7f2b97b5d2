"""2+ GPU equivalence check (run under torchrun, not collected by pytest):
   torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tests/dist_equiv.py
The node-parallel step must reproduce the 1-GPU engine: logits, loss and updated parameters (dropout off so that
no mask bookkeeping is involved; a second pass checks dropout statistics only)."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import efficient_gnns_b200  # noqa: E402,F401
from efficient_gnns_b200 import sparse, synthetic  # noqa: E402
from efficient_gnns_b200.dist import ShardedGCNTrainer  # noqa: E402
from efficient_gnns_b200.engine import GCNStudentTrainer  # noqa: E402


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", os.environ["RANK"]))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n, e, dims = 20_011, 150_000, [128, 256, 256, 40]
    ei = synthetic.skewed_edges(n, e, 0).to(dev)
    perm = (ei[1] * n + ei[0]).argsort()
    adj = sparse.SparseTensor(row=ei[1][perm], col=ei[0][perm], sparse_sizes=(n, n), is_sorted=True).to_symmetric()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n, dims[0], generator=g).to(dev)
    y = torch.randint(0, dims[-1], (n,), generator=g).to(dev)
    t = (torch.randn(n, dims[-1], generator=g) * 2).to(dev)
    idx = torch.randperm(n, generator=g)[: n // 2].sort().values.to(dev)

    ref = GCNStudentTrainer(adj, dims, dropout=0.0, seed=3)
    sh = ShardedGCNTrainer(adj, dims, dropout=0.0, seed=3)
    xp, yl, il, tl = sh.shard_inputs(x, y, idx, t)
    # Criteria.  Forward quantities are smooth: max-norm 1e-5.  Gradients pass through ReLU masks: an activation
    # within rounding distance of 0 may flip between the two summation orders and moves a handful of gradient
    # entries by O(1e-3) relative, and Adam (update = lr*sign(g) on step 1) amplifies any sign change of a tiny
    # entry to 2*lr — so gradients are compared in the Frobenius norm and parameters by the loss they produce.
    ok = True

    def fro(a, b):
        return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
    for step in range(3):
        l_ref = ref.train_step(x, y, idx, t).clone()
        l_sh = sh.train_step(xp, yl, il, tl).clone()
        logits = sh.gather_rows(sh.Y[-1])
        e_logit = ((logits - ref.Y[-1]).abs().max() / ref.Y[-1].abs().max()).item()
        e_loss = ((l_sh - l_ref).abs() / l_ref.abs().clamp_min(1e-12)).max().item()
        e_grad = 0.0
        for l in range(ref.L):
            e_grad = max(e_grad, fro(sh.gW[l], ref.gW[l]))
        e_grad = max(e_grad, fro(sh.gb[-1], ref.gb[-1]))
        for l in range(ref.L - 1):
            e_grad = max(e_grad, fro(sh.ggamma[l], ref.ggamma[l]), fro(sh.gbeta[l], ref.gbeta[l]))
        if rank == 0:
            print(f"step {step}: logits max-rel err {e_logit:.2e}  loss rel err {e_loss:.2e}  grad Frobenius rel err {e_grad:.2e}",
                  flush=True)
        if step == 0:
            ok &= e_logit < 1e-5 and e_loss < 1e-5 and e_grad < 2e-3
        else:   # trajectories separate slowly (Adam turns sign flips of tiny entries into 2*lr parameter moves)
            ok &= e_loss < 1e-3 and e_grad < 0.3
    # replicas stay bit-identical across ranks
    p0 = sh.params.clone()
    dist.broadcast(p0, 0)
    same = torch.equal(p0, sh.params)
    # dropout on: just run and check the loss is finite and identical on all ranks
    sh2 = ShardedGCNTrainer(adj, dims, dropout=0.5, seed=3)
    l2 = sh2.train_step(*sh2.shard_inputs(x, y, idx, t)).clone()
    l2b = l2.clone(); dist.broadcast(l2b, 0)
    ok &= bool(torch.isfinite(l2).all()) and torch.equal(l2, l2b) and same
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_EQUIV", "PASS" if flag.item() == 1 else "FAIL", "replicas identical:", same, flush=True)
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
