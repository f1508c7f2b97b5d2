"""The attribute bag the reference scripts pass around (torch_geometric.data.Data): .to(device) also moves SparseTensors."""
import torch

from .sparse import SparseTensor


class Data:
    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def num_nodes(self):
        if getattr(self, "_num_nodes", None) is not None:
            return self._num_nodes
        return self.x.size(0)

    @num_nodes.setter
    def num_nodes(self, v):
        self._num_nodes = v

    @property
    def num_features(self):
        return self.x.size(1)

    def to(self, device, *a, **k):
        for name, v in list(self.__dict__.items()):
            if isinstance(v, (torch.Tensor, SparseTensor)):
                setattr(self, name, v.to(device))
        return self

    def __repr__(self):
        return "Data(" + ", ".join(f"{k}={tuple(v.shape) if isinstance(v, torch.Tensor) else v}" for k, v in self.__dict__.items()) + ")"
