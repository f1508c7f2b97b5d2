"""The shim packages expose every name the reference imports, importable without a GPU."""
import importlib
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
SHIM = ROOT / "efficient-gnns_b200" / "shim"


@pytest.fixture()
def shim_path():
    sys.path.insert(0, str(SHIM))
    yield
    sys.path.remove(str(SHIM))
    for m in [m for m in sys.modules if m.split(".")[0] in ("torch_geometric", "torch_sparse", "torch_scatter", "ogb")]:
        del sys.modules[m]


def test_reference_import_surface(shim_path):
    import torch_geometric
    import torch_geometric.transforms as T
    from torch_geometric.nn import GCNConv, SAGEConv, MessagePassing
    from torch_geometric.utils import softmax, to_dense_adj, subgraph, negative_sampling, add_self_loops, to_undirected
    from torch_geometric.utils.hetero import group_hetero_graph
    from torch_geometric.data import Data, GraphSAINTRandomWalkSampler, DataLoader
    from torch_geometric.datasets import PPI
    from torch_sparse import SparseTensor
    from torch_scatter import scatter
    from ogb.nodeproppred import PygNodePropPredDataset, Evaluator
    assert callable(T.ToSparseTensor) and torch_geometric.__version__


def test_dataset_stub_and_transform_on_cpu(shim_path):
    import torch_geometric.transforms as T
    from ogb.nodeproppred import PygNodePropPredDataset, Evaluator
    ds = PygNodePropPredDataset(name="ogbn-arxiv-plumbing", transform=T.ToSparseTensor())
    data = ds[0]
    assert data.edge_index is None and data.adj_t.nnz() == 50_000
    data.adj_t = data.adj_t.to_symmetric()
    row, col, _ = data.adj_t.coo()
    assert bool((row[1:] >= row[:-1]).all())
    split = ds.get_idx_split()
    assert set(split) == {"train", "valid", "test"} and ds.num_classes == 40
    acc = Evaluator("ogbn-arxiv").eval({"y_true": data.y[split["train"]], "y_pred": data.y[split["train"]]})["acc"]
    assert acc == 1.0


def test_subgraph_and_hetero_match_oracle(shim_path):
    import numpy as np
    from torch_geometric.utils import subgraph, to_undirected
    from torch_geometric.utils.hetero import group_hetero_graph
    from oracle import graph as og
    g = torch.Generator().manual_seed(0)
    ei = torch.randint(0, 50, (2, 400), generator=g)
    subset = torch.randperm(50, generator=g)[:20]
    got, _ = subgraph(subset, ei, relabel_nodes=True)
    ref, _ = og.subgraph(subset.numpy(), ei.numpy(), True)
    assert np.array_equal(got.numpy(), ref)
    und = to_undirected(ei[:, ei[0] != ei[1]], 50)
    assert np.array_equal(und.numpy(), og.to_undirected(ei[:, ei[0] != ei[1]].numpy(), 50))
    eid = {("a", "r", "b"): torch.tensor([[0, 1], [2, 0]]), ("b", "s", "a"): torch.tensor([[1], [1]])}
    e, et, nt, li, l2g, k2i = group_hetero_graph(eid, {"a": 2, "b": 3})   # the 6-tuple mag_pyg/gnn.py:346-347 unpacks
    assert e.tolist() == [[0, 1, 3], [4, 2, 1]] and et.tolist() == [0, 0, 1] and nt.tolist() == [0, 0, 1, 1, 1]
    assert li.tolist() == [0, 1, 0, 1, 2] and k2i["b"] == 1 and k2i[("b", "s", "a")] == 1
    assert l2g["a"].tolist() == [0, 1] and l2g["b"].tolist() == [2, 3, 4]


def test_mag_shaped_dataset_and_grouping_as_the_reference_main_uses_them():
    """mag_pyg/gnn.py:308-356 up to the sampler: dataset fields, reverse relations, to_undirected, group_hetero_graph, the
    homogeneous Data with labels and train mask."""
    import sys
    from pathlib import Path
    import torch
    import efficient_gnns_b200
    sys.path.insert(0, str(Path(efficient_gnns_b200.__file__).resolve().parent / "shim"))
    try:
        from ogb.nodeproppred import PygNodePropPredDataset
        from torch_geometric.data import Data
        from torch_geometric.utils.hetero import group_hetero_graph
        dataset = PygNodePropPredDataset(name="ogbn-mag-plumbing")
        data = dataset[0]
        split_idx = dataset.get_idx_split()
        assert dataset.num_classes == 349 and set(data.x_dict) == {"paper"} and data.x_dict["paper"].shape[1] == 128
        eid = data.edge_index_dict
        assert set(eid) == {("author", "affiliated_with", "institution"), ("author", "writes", "paper"),
                            ("paper", "cites", "paper"), ("paper", "has_topic", "field_of_study")}
        r, c = eid[("author", "writes", "paper")]
        assert int(r.max()) < data.num_nodes_dict["author"] and int(c.max()) < data.num_nodes_dict["paper"]
        eid[("paper", "to", "author")] = torch.stack([c, r])
        out = group_hetero_graph(eid, data.num_nodes_dict)
        edge_index, edge_type, node_type, local_node_idx, local2global, key2int = out
        n = sum(data.num_nodes_dict.values())
        assert node_type.numel() == n and int(edge_index.max()) < n and int(edge_type.max()) == 4
        homo = Data(edge_index=edge_index, edge_attr=edge_type, node_type=node_type, local_node_idx=local_node_idx, num_nodes=n)
        homo.y = node_type.new_full((n, 1), -1)
        homo.y[local2global["paper"]] = data.y_dict["paper"]
        homo.train_mask = torch.zeros(n, dtype=torch.bool)
        homo.train_mask[local2global["paper"][split_idx["train"]["paper"]]] = True
        assert homo.num_nodes == n and int(homo.train_mask.sum()) == split_idx["train"]["paper"].numel()
        assert bool((homo.y[homo.train_mask] >= 0).all())
    finally:
        sys.path.pop(0)
