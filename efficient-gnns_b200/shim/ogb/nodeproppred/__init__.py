"""ogb.nodeproppred.{PygNodePropPredDataset, Evaluator} (arxiv_pyg/gnn.py:16,236-244,268)."""
import torch

import efficient_gnns_b200  # noqa: F401
from efficient_gnns_b200 import synthetic
from torch_geometric.data import Data

_SHAPES = {"ogbn-arxiv": synthetic.ARXIV, "ogbn-arxiv-plumbing": synthetic.PLUMBING}
_MAG_SCALE = {"ogbn-mag": 1.0, "ogbn-mag-plumbing": 0.01}        # heterogeneous: mag_pyg/gnn.py:308-321


class PygNodePropPredDataset:
    def __init__(self, name, root="dataset", transform=None, pre_transform=None):
        self.name, self.transform = name, transform
        self.processed_dir = root
        if name in _MAG_SCALE:
            self._mag = synthetic.make_mag_dataset(_MAG_SCALE[name])
            self.num_classes = self._mag["num_classes"]
            return
        if name not in _SHAPES:
            raise NotImplementedError(f"{name}: only ARXIV- and MAG-shape synthetics are served (no network)")
        self._mag = None
        self._ds = synthetic.make_node_dataset(_SHAPES[name], seed=0)
        self.num_classes = self._ds.num_classes

    def __len__(self):
        return 1

    def __getitem__(self, idx):
        if self._mag is not None:
            m = self._mag
            data = Data(num_nodes_dict=dict(m["num_nodes_dict"]), edge_index_dict=dict(m["edge_index_dict"]),
                        x_dict=dict(m["x_dict"]), y_dict=dict(m["y_dict"]), node_year_dict={}, edge_reltype_dict={})
            return self.transform(data) if self.transform is not None else data
        d = self._ds
        data = Data(x=d.x, y=d.y, edge_index=d.edge_index)
        data.num_nodes = d.num_nodes
        return self.transform(data) if self.transform is not None else data

    def get_idx_split(self):
        if self._mag is not None:
            return {k: dict(v) for k, v in self._mag["split_idx"].items()}
        return dict(self._ds.split_idx)


class Evaluator:
    def __init__(self, name):
        self.name = name

    def eval(self, input_dict):
        y_true, y_pred = input_dict["y_true"], input_dict["y_pred"]
        y_true = torch.as_tensor(y_true).view(-1).cpu()
        y_pred = torch.as_tensor(y_pred).view(-1).cpu()
        return {"acc": float((y_true == y_pred).float().mean())}
