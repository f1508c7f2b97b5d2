"""ogb.nodeproppred.{PygNodePropPredDataset, Evaluator} (arxiv_pyg/gnn.py:16,236-244,268)."""
import torch

import efficient_gnns_b200  # noqa: F401
from efficient_gnns_b200 import synthetic
from torch_geometric.data import Data

_SHAPES = {"ogbn-arxiv": synthetic.ARXIV, "ogbn-arxiv-plumbing": synthetic.PLUMBING}


class PygNodePropPredDataset:
    def __init__(self, name, root="dataset", transform=None, pre_transform=None):
        if name not in _SHAPES:
            raise NotImplementedError(f"{name}: only the ARXIV-shape synthetic is served (no network)")
        self.name, self.transform = name, transform
        self._ds = synthetic.make_node_dataset(_SHAPES[name], seed=0)
        self.num_classes = self._ds.num_classes
        self.processed_dir = root

    def __len__(self):
        return 1

    def __getitem__(self, idx):
        d = self._ds
        data = Data(x=d.x, y=d.y, edge_index=d.edge_index)
        data.num_nodes = d.num_nodes
        return self.transform(data) if self.transform is not None else data

    def get_idx_split(self):
        return dict(self._ds.split_idx)


class Evaluator:
    def __init__(self, name):
        self.name = name

    def eval(self, input_dict):
        y_true, y_pred = input_dict["y_true"], input_dict["y_pred"]
        y_true = torch.as_tensor(y_true).view(-1).cpu()
        y_pred = torch.as_tensor(y_pred).view(-1).cpu()
        return {"acc": float((y_true == y_pred).float().mean())}
