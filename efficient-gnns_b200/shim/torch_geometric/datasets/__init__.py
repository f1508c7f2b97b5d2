"""torch_geometric.datasets.PPI (ppi_pyg/gnn.py:305-307).  There is no network here, so what is served is a synthetic dataset
of the PPI SHAPE (24 graphs: 20/2/2, ~2.4 k nodes, 50 features, 121 multi-hot labels — efficient_gnns_b200.synthetic.PPI),
with the interface the reference reads: len(), indexing -> Data(x, y, edge_index), num_features, num_classes."""
from efficient_gnns_b200 import synthetic
from efficient_gnns_b200.graphdata import Data


class PPI:
    def __init__(self, root=None, split="train", transform=None, pre_transform=None, pre_filter=None):
        if split not in ("train", "val", "test"):
            raise ValueError(f"split={split!r}")
        self.root, self.split, self.transform = root, split, transform
        self._graphs = [Data(x=x, y=y, edge_index=ei) for x, y, ei in synthetic.make_ppi_graphs(split)]
        self.num_features = self.num_node_features = synthetic.PPI["num_features"]
        self.num_classes = synthetic.PPI["num_classes"]

    def __len__(self):
        return len(self._graphs)

    def __getitem__(self, idx):
        g = self._graphs[idx]
        return self.transform(g) if self.transform is not None else g
