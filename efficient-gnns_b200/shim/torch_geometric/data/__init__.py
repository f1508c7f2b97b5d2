"""torch_geometric.data: Data (attribute bag) and the device-side mini-batch loaders."""
from efficient_gnns_b200.graphdata import Data  # noqa: F401
from efficient_gnns_b200.sampling import Batch, DataLoader, GraphSAINTRandomWalkSampler  # noqa: F401
