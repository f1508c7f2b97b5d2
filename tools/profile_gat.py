"""Three GAT layer steps (default H=8, D=32; argv: H D) on the ARXIV-shape graph, for an ncu launch list."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import efficient_gnns_b200  # noqa
from efficient_gnns_b200 import nn as bnn, sparse, synthetic
ds = synthetic.make_node_dataset(synthetic.ARXIV, seed=0)
n = ds.num_nodes
ei = ds.edge_index.cuda()
perm = (ei[1] * n + ei[0]).argsort()
adj = bnn._fill_diag_pattern(sparse.SparseTensor(row=ei[1][perm], col=ei[0][perm], sparse_sizes=(n, n), is_sorted=True).to_symmetric())
x = ds.x.cuda()
H, D = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (8, 32)
layer = bnn.DGLGATConv(128, D, num_heads=H, use_symmetric_norm=True).cuda()
for _ in range(3):
    out = layer(adj, x); layer.zero_grad(); out.sum().backward()
torch.cuda.synchronize()
