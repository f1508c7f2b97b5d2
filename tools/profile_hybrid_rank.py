"""One rank's share of a P-GPU hybrid-layout step on ONE GPU with the exchanges skipped (timing diagnostics; results are wrong):
   ncu --metrics gpu__time_duration.sum ... python tools/profile_hybrid_rank.py --world 8 --steps 3
gives the per-kernel launch list of what every rank executes."""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import efficient_gnns_b200  # noqa: E402,F401
from efficient_gnns_b200 import sparse, synthetic  # noqa: E402
from efficient_gnns_b200.hybrid import HybridGCNTrainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=2)
    ap.add_argument("--rank", type=int, default=0)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--graph", action="store_true")
    a = ap.parse_args()
    ds = synthetic.make_node_dataset(synthetic.ARXIV, seed=0)
    n = ds.num_nodes
    ei = ds.edge_index.cuda()
    perm = (ei[1] * n + ei[0]).argsort()
    adj = sparse.SparseTensor(row=ei[1][perm], col=ei[0][perm], sparse_sizes=(n, n), is_sorted=True).to_symmetric()
    tr = HybridGCNTrainer(adj, [128, 256, 256, 40], dropout=0.5, lr=0.01, seed=0, exchange="null", _fake=(a.rank, a.world))
    xin, yl, il, tl = tr.shard_inputs(ds.x, ds.y.squeeze(1), ds.split_idx["train"], ds.teacher_logits)
    for _ in range(2):
        tr.train_step(xin, yl, il, tl)
    torch.cuda.synchronize()
    if a.graph:
        tr.capture(xin, yl, il, tl, warmup=1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(a.steps):
        tr.replay() if a.graph else tr.train_step(xin, yl, il, tl)
    e1.record()
    torch.cuda.synchronize()
    print(f"world {a.world} rank {a.rank}: {e0.elapsed_time(e1) / a.steps:.3f} ms/step (exchanges skipped)")


if __name__ == "__main__":
    main()
