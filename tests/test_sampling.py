"""Mini-batch regimes: GraphSAINT random-walk sampling and small-graph batching.
CPU: the oracle's own invariants + the host-side batching; GPU: the kernels bit-exact against the oracle."""
import numpy as np
import pytest
import torch

import efficient_gnns_b200  # noqa: F401
from efficient_gnns_b200 import sampling
from efficient_gnns_b200.graphdata import Data
from efficient_gnns_b200.synthetic import skewed_edges
from oracle import sampling as osamp


def graph(n=3000, e=20_000, seed=0, isolated=True):
    ei = skewed_edges(n, e, seed).numpy()
    if isolated:                                   # a block of nodes without out-edges: walkers that land there stay
        ei = ei[:, ei[0] < n - 50]
    return ei


# ------------------------------------------------------------------ oracle (CPU)
def test_oracle_philox_known_answers():
    """Philox4x32-10 with a zero key and counter, and the all-ones test vector of the Random123 distribution."""
    assert osamp.philox4x32(0, 0, np.array([0], dtype=np.uint64))[0].tolist() == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    r = osamp.philox4x32(0xFFFFFFFFFFFFFFFF, 0xFFFFFFFFFFFFFFFF, np.array([0xFFFFFFFFFFFFFFFF], dtype=np.uint64))[0].tolist()
    assert r == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]


def test_oracle_walks_follow_edges_and_subgraph_is_induced():
    n = 800
    ei = graph(n, 6000, 1)
    rowptr, col, eid = osamp.csr_by_source(ei, n)
    start = np.random.default_rng(0).integers(0, n, 500)
    walks = osamp.random_walk(rowptr, col, start, 3, seed=5, offset=2)
    edges = set(zip(ei[0].tolist(), ei[1].tolist()))
    deg = np.diff(rowptr)
    for w in walks:
        for a, b in zip(w[:-1], w[1:]):
            assert (a, b) in edges or (deg[a] == 0 and a == b)
    nodes = np.unique(walks)
    r, c, e = osamp.saint_subgraph(rowptr, col, eid, nodes)
    sel = np.zeros(n, bool); sel[nodes] = True
    want = np.nonzero(sel[ei[0]] & sel[ei[1]])[0]
    assert sorted(e.tolist()) == sorted(want.tolist())                 # exactly the edges with both endpoints sampled
    assert np.array_equal(nodes[r], ei[0][e]) and np.array_equal(nodes[c], ei[1][e])
    key = ei[0][e] * n + ei[1][e]
    assert np.all(np.diff(key) >= 0)                                   # parent (row, col) order


def test_oracle_walk_is_uniform_over_neighbours():
    rowptr = np.array([0, 4, 4, 4, 4, 4]); col = np.array([1, 2, 3, 4])
    w = osamp.random_walk(rowptr, col, np.zeros(40_000, dtype=np.int64), 1, seed=9, offset=0)
    freq = np.bincount(w[:, 1], minlength=5)[1:] / 40_000
    assert np.abs(freq - 0.25).max() < 0.01


def test_sampler_without_a_gpu_raises_instead_of_sampling_on_the_cpu():
    if torch.cuda.is_available():
        pytest.skip("needs a CPU-only environment")
    d = Data(edge_index=torch.randint(0, 10, (2, 30)))
    d.num_nodes = 10
    with pytest.raises(Exception, match="no CPU fallback"):
        sampling.GraphSAINTRandomWalkSampler(d, batch_size=4)


def test_small_graph_batches_cpu():
    """DataLoader over PPI-like graphs (ppi_pyg/gnn.py:305-310): disjoint union, offsets, graph ids."""
    gs = []
    for i, n in enumerate((5, 7, 3)):
        g = torch.Generator().manual_seed(i)
        gs.append(Data(x=torch.randn(n, 4, generator=g), y=torch.randint(0, 2, (n, 3), generator=g).float(),
                       edge_index=torch.randint(0, n, (2, 2 * n), generator=g)))
    loader = sampling.DataLoader(gs, batch_size=2, shuffle=False)
    batches = list(loader)
    assert len(loader) == 2 and len(batches) == 2
    b = batches[0]
    assert b.num_nodes == 12 and b.num_graphs == 2 and b.x.shape == (12, 4) and b.y.shape == (12, 3)
    assert torch.equal(b.batch, torch.tensor([0] * 5 + [1] * 7))
    assert torch.equal(b.edge_index[:, :10], gs[0].edge_index) and torch.equal(b.edge_index[:, 10:], gs[1].edge_index + 5)
    assert batches[1].num_nodes == 3
    order = [int(b.x[0, 0] * 1e6) for b in sampling.DataLoader(gs, batch_size=1, shuffle=True, seed=3)]
    assert sorted(order) == sorted(int(g.x[0, 0] * 1e6) for g in gs)


def test_shim_exports_the_loaders():
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(efficient_gnns_b200.__file__).resolve().parent / "shim"))
    try:
        from torch_geometric.data import Data as D, DataLoader, GraphSAINTRandomWalkSampler
        assert D is Data and DataLoader is sampling.DataLoader and GraphSAINTRandomWalkSampler is sampling.GraphSAINTRandomWalkSampler
    finally:
        sys.path.pop(0)


# ------------------------------------------------------------------ kernels (GPU)
@pytest.mark.gpu
@pytest.mark.parametrize("walk_length", [1, 2, 5, 9])
def test_random_walk_kernel_replays_bit_exact(walk_length):
    n = 3000
    ei = graph(n, 20_000, 2)
    rowptr, col, _ = osamp.csr_by_source(ei, n)
    g = sampling.SaintGraph(torch.from_numpy(ei).cuda(), n)
    assert np.array_equal(g.rowptr.cpu().numpy(), rowptr) and np.array_equal(g.col.cpu().numpy(), col)
    start = torch.randint(0, n, (7001,), generator=torch.Generator().manual_seed(1))
    got = sampling.random_walk(g.rowptr, g.col, start.cuda(), walk_length, seed=1234567, offset=17).cpu().numpy()
    want = osamp.random_walk(rowptr, col, start.numpy(), walk_length, 1234567, 17)
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_saint_subgraph_kernel_bit_exact_and_workspace_restored():
    n = 5000
    ei = graph(n, 60_000, 3)
    rowptr, col, eid = osamp.csr_by_source(ei, n)
    g = sampling.SaintGraph(torch.from_numpy(ei).cuda(), n)
    assert np.array_equal(g.eid.cpu().numpy(), eid)
    for k, seed in ((1, 0), (37, 1), (900, 2), (5000, 3)):
        nodes = np.sort(np.random.default_rng(seed).choice(n, size=k, replace=False))
        ei_loc, e_id = g.subgraph(torch.from_numpy(nodes).cuda())
        r, c, e = osamp.saint_subgraph(rowptr, col, eid, nodes)
        assert np.array_equal(ei_loc[0].cpu().numpy(), r) and np.array_equal(ei_loc[1].cpu().numpy(), c)
        assert np.array_equal(e_id.cpu().numpy(), e)
        assert int((g.node_map != -1).sum()) == 0
    ei_loc, e_id = g.subgraph(torch.empty(0, dtype=torch.long, device="cuda"))
    assert ei_loc.shape == (2, 0) and e_id.numel() == 0


@pytest.mark.gpu
def test_graphsaint_sampler_batches_like_the_reference_loop():
    """The attributes mag_pyg/gnn.py:187-190 reads from a batch: edge_index, edge_attr, node_type, local_node_idx, train_mask, y."""
    n = 4000
    ei = torch.from_numpy(graph(n, 30_000, 4, isolated=False))
    E = ei.size(1)
    g = torch.Generator().manual_seed(0)
    data = Data(edge_index=ei, edge_attr=torch.randint(0, 7, (E,), generator=g), node_type=torch.randint(0, 4, (n,), generator=g),
                local_node_idx=torch.arange(n), train_mask=torch.rand(n, generator=g) < 0.3, y=torch.randint(0, 9, (n,), generator=g),
                num_classes=9)
    data.num_nodes = n
    host_loader = sampling.GraphSAINTRandomWalkSampler(data, batch_size=10, walk_length=1, num_steps=1)   # host-resident parent, as
    b0 = next(iter(host_loader))                                         # mag_pyg/gnn.py:361 builds it: uploaded once,
    assert b0.edge_index.is_cuda and b0.y.is_cuda and not data.edge_index.is_cuda   # batches are born on the device
    data = data.to("cuda")
    loader = sampling.GraphSAINTRandomWalkSampler(data, batch_size=500, walk_length=2, num_steps=3, sample_coverage=0, seed=7)
    assert len(loader) == 3
    epoch1 = list(loader)
    assert len(epoch1) == 3
    rowptr, col, eid = osamp.csr_by_source(ei.numpy(), n)
    for step, b in enumerate(epoch1):
        nodes = b.n_id.cpu().numpy()
        assert np.all(np.diff(nodes) > 0) and b.num_nodes == len(nodes) and 500 <= len(nodes) * 3
        # the batch is the oracle's induced sub-graph of its node set, attributes sliced by node / edge id
        r, c, e = osamp.saint_subgraph(rowptr, col, eid, nodes)
        assert np.array_equal(b.edge_index.cpu().numpy(), np.stack([r, c])) and np.array_equal(b.e_id.cpu().numpy(), e)
        assert torch.equal(b.edge_attr, data.edge_attr[b.e_id]) and torch.equal(b.node_type, data.node_type[b.n_id])
        assert torch.equal(b.y, data.y[b.n_id]) and torch.equal(b.train_mask, data.train_mask[b.n_id]) and b.num_classes == 9
        # ... and the node set is the oracle's replay of the walks from the same roots
        walks = osamp.random_walk(rowptr, col, loader_roots(loader, step), 2, 7, step)
        assert np.array_equal(np.unique(walks), nodes)
    epoch2 = list(loader)
    assert not torch.equal(epoch2[0].n_id, epoch1[0].n_id)                   # a new epoch draws new walks
    again = sampling.GraphSAINTRandomWalkSampler(data, batch_size=500, walk_length=2, num_steps=3, seed=7)
    assert torch.equal(next(iter(again)).n_id, epoch1[0].n_id)               # same seed, same batches


def loader_roots(loader, step):
    g = torch.Generator(device="cuda")
    g.manual_seed(loader.seed * 1_000_003 + step)
    return torch.randint(0, loader.N, (loader.batch_size,), generator=g, device="cuda").cpu().numpy()


@pytest.mark.gpu
@pytest.mark.timeout(300)
def test_graphsaint_epochs_train_an_rgcn_on_sampled_batches():
    """The reference's MAG training loop (mag_pyg/gnn.py:174-204) on this sampler: every batch is a sub-graph of the grouped
    heterogeneous graph with its edge types / node types / local ids sliced along; R-GCN forward+backward on the mirrored
    MessagePassing surface; the loss falls over a few epochs."""
    import torch.nn.functional as F
    from test_rgcn_gpu import RelNet
    g = torch.Generator().manual_seed(0)
    n_paper, n_author = 1500, 900
    n = n_paper + n_author
    node_type = torch.cat([torch.zeros(n_paper, dtype=torch.long), torch.ones(n_author, dtype=torch.long)])
    local_idx = torch.cat([torch.arange(n_paper), torch.arange(n_author)])
    cites = torch.randint(0, n_paper, (2, 6000), generator=g)
    writes = torch.stack([torch.randint(0, n_author, (5000,), generator=g) + n_paper, torch.randint(0, n_paper, (5000,), generator=g)])
    edge_index = torch.cat([cites, writes, writes.flip(0)], 1)
    edge_type = torch.cat([torch.zeros(6000), torch.ones(5000), torch.full((5000,), 2.0)]).long()
    x_paper = torch.randn(n_paper, 16, generator=g)
    w_true = torch.randn(16, 5, generator=g)
    y = torch.full((n, 1), -1, dtype=torch.long)
    y[:n_paper, 0] = (x_paper @ w_true).argmax(1)                        # learnable labels on the paper nodes
    train_mask = torch.zeros(n, dtype=torch.bool)
    train_mask[:n_paper] = torch.rand(n_paper, generator=g) < 0.6
    data = Data(edge_index=edge_index, edge_attr=edge_type, node_type=node_type, local_node_idx=local_idx, y=y, train_mask=train_mask)
    data.num_nodes = n
    loader = sampling.GraphSAINTRandomWalkSampler(data.to("cuda"), batch_size=400, walk_length=2, num_steps=4, sample_coverage=0, seed=1)
    torch.manual_seed(0)
    model = RelNet(16, 32, 5, {0: n_paper, 1: n_author}, [0], 3).cuda()
    for p in model.parameters():
        torch.nn.init.normal_(p, std=0.1)
    opt = torch.optim.Adam(model.parameters(), lr=0.01)
    x_dict = {0: x_paper.cuda()}
    losses = []
    for epoch in range(6):
        tot = n_ex = 0
        for batch in loader:
            opt.zero_grad()
            out = model(x_dict, batch.edge_index, batch.edge_attr, batch.node_type, batch.local_node_idx)[batch.train_mask]
            tgt = batch.y[batch.train_mask].squeeze(1)
            assert (tgt >= 0).all()                                    # only paper nodes are ever in the train mask
            loss = F.nll_loss(F.log_softmax(out, dim=-1), tgt)
            loss.backward()
            opt.step()
            tot += float(loss) * tgt.numel(); n_ex += tgt.numel()
        losses.append(tot / n_ex)
    assert all(np.isfinite(losses)) and losses[-1] < 0.8 * losses[0], losses


def test_ppi_shaped_dataset_on_the_shim_cpu():
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(efficient_gnns_b200.__file__).resolve().parent / "shim"))
    try:
        from torch_geometric.data import DataLoader
        from torch_geometric.datasets import PPI
        val = PPI("data/PPI/", split="val")
        assert len(val) == 2 and val.num_features == 50 and val.num_classes == 121
        b = next(iter(DataLoader(val, batch_size=2, shuffle=False)))
        assert b.x.shape[1] == 50 and b.y.shape == (b.num_nodes, 121) and b.num_graphs == 2
        assert set(b.y.unique().tolist()) <= {0.0, 1.0} and 0.05 < float(b.y.mean()) < 0.6
        ei = val[0].edge_index
        assert torch.equal(torch.unique(ei[0] * val[0].num_nodes + ei[1]), torch.unique(ei[1] * val[0].num_nodes + ei[0]))  # symmetric
    finally:
        sys.path.pop(0)


@pytest.mark.gpu
@pytest.mark.timeout(300)
def test_ppi_style_epochs_small_graph_batches_gat_student():
    """ppi_pyg/gnn.py:185-274: batch_size-1 loader over small graphs, PyG-style GAT student (GATConv + skip Linear, ELU), BCE on
    multi-hot labels, micro-F1 — on the mirrored surface; the loss falls and F1 rises over a few epochs."""
    import torch.nn.functional as F
    from efficient_gnns_b200 import nn as bnn, synthetic
    from efficient_gnns_b200.criterion import bce_with_logits
    graphs = [Data(x=x, y=y, edge_index=ei).to("cuda") for x, y, ei in synthetic.make_ppi_graphs("train", scale=0.25)[:6]]
    loader = sampling.DataLoader(graphs, batch_size=1, shuffle=True, seed=0)

    class Student(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.c1, self.l1 = bnn.GATConv(50, 16, heads=4), torch.nn.Linear(50, 64)
            self.c2, self.l2 = bnn.GATConv(64, 121, heads=4, concat=False), torch.nn.Linear(64, 121)

        def forward(self, x, ei):
            x = F.elu(self.c1(x, ei) + self.l1(x))
            return self.c2(x, ei) + self.l2(x)

    torch.manual_seed(0)
    model = Student().cuda()
    opt = torch.optim.Adam(model.parameters(), lr=0.01)

    def micro_f1():
        model.eval()
        tp = fp = fn = 0
        with torch.no_grad():
            for b in sampling.DataLoader(graphs, batch_size=2):
                pred = model(b.x, b.edge_index) > 0
                tp += int((pred & (b.y > 0)).sum()); fp += int((pred & (b.y == 0)).sum()); fn += int((~pred & (b.y > 0)).sum())
        model.train()
        return 2 * tp / max(2 * tp + fp + fn, 1)

    f0, losses = micro_f1(), []
    for epoch in range(8):
        tot = 0.0
        for b in loader:
            opt.zero_grad()
            loss = bce_with_logits(model(b.x, b.edge_index), b.y)
            loss.backward()
            opt.step()
            tot += float(loss)
        losses.append(tot / len(loader))
    assert losses[-1] < 0.9 * losses[0] and micro_f1() > f0, (losses, f0)
