"""Seeded inputs of the round-2 fixtures, shared by the generator (make_golden_r2.py, which feeds them to the reference's
files) and by the tests (which feed them to oracle/ and to the CUDA path).  Regenerated instead of stored: the fixtures
keep only what the reference computed, plus a checksum of these inputs so that generator drift is detected."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
from oracle import graph as og  # noqa: E402


def hub_graph(n, e, seed):
    """(directed edge_index, symmetric row, col, rowptr) of the skewed synthetic graph (heavy rows: thousands of edges)."""
    from efficient_gnns_b200.synthetic import skewed_edges
    ei = skewed_edges(n, e, seed).numpy()
    row, col, _ = og.to_sparse_adj_t(ei, n)
    r, c = og.to_symmetric(row, col, n)
    return ei, r, c, og.ind2ptr(r, n)


def hub_model():
    n, Fin, Hd, C = 3000, 16, 256, 8
    ei_d, r, c, rowptr = hub_graph(n, 40_000, 7)
    g = torch.Generator().manual_seed(41)
    x = torch.randn(n, Fin, generator=g)
    y = torch.randint(0, C, (n,), generator=g)
    train_idx = torch.randperm(n, generator=g)[:1800].sort().values
    return dict(n=n, dims=(Fin, Hd, C), edge_index_directed=torch.from_numpy(ei_d), r=r, c=c, rowptr=rowptr, x=x, y=y,
                train_idx=train_idx, rows=torch.arange(0, n, 10))


def lsp_wide():
    hm = hub_model()
    edge_index = torch.from_numpy(np.stack([hm["r"], hm["c"]]))
    sub_ei = torch.from_numpy(og.subgraph(hm["train_idx"].numpy(), edge_index.numpy(), True)[0])
    nt, C = hm["train_idx"].numel(), 8
    g = torch.Generator().manual_seed(43)
    z = torch.randn(nt, C, generator=g)
    yy = torch.randint(0, C, (nt,), generator=g)
    fs = torch.randn(nt, 256, generator=g).relu() + 0.01
    ft = torch.randn(nt, 750, generator=g).relu() + 0.01
    return dict(logits=z, labels=yy, feat=fs, t_feat=ft, sub_edge_index=sub_ei, rows=torch.arange(0, nt, 6),
                scales={"cosine": (1.0, 1.0), "rbf": (0.05, 0.03)})     # rbf: keep exp(-d^2/2) away from underflow


def gat_wide():
    n = 1200
    _, r, c, rp = hub_graph(n, 12_000, 9)
    rs, cs, _ = og.fill_diag(r, c, np.ones(r.shape[0], dtype=np.float32), n)    # row = destination, col = source
    g = torch.Generator().manual_seed(47)
    x = torch.randn(n, 32, generator=g)
    w = torch.randn(n, 3, 250, generator=g)
    return dict(n=n, row=torch.from_numpy(rs), col=torch.from_numpy(cs), x=x, w=w, rows=torch.arange(0, n, 8),
                max_degree=int(np.diff(rp).max()) + 1)


def checksum(*tensors) -> float:
    return float(sum(t.double().abs().sum() for t in tensors))
