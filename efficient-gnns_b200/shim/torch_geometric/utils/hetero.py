"""torch_geometric.utils.hetero.group_hetero_graph (mag_pyg/gnn.py:16,346): merge a heterogeneous graph into one
homogeneous index space (node types numbered in dict order, global ids by cumulative offsets).  Returns, like the
PyG version the reference unpacks at mag_pyg/gnn.py:346-347,
``(edge_index, edge_type, node_type, local_node_idx, local2global, key2int)``."""
import torch


def group_hetero_graph(edge_index_dict, num_nodes_dict=None):
    if num_nodes_dict is None:
        num_nodes_dict = {}
        for (s, _, d), ei in edge_index_dict.items():
            num_nodes_dict[s] = max(num_nodes_dict.get(s, 0), int(ei[0].max()) + 1)
            num_nodes_dict[d] = max(num_nodes_dict.get(d, 0), int(ei[1].max()) + 1)
    key2int, offset, cum = {}, {}, 0
    node_types, local_idx, local2global = [], [], {}
    for i, (key, n) in enumerate(num_nodes_dict.items()):
        key2int[key] = i
        offset[key] = cum
        cum += n
        node_types.append(torch.full((n,), i, dtype=torch.long))
        local_idx.append(torch.arange(n))
        local2global[key] = torch.arange(n) + offset[key]
    eis, ets = [], []
    for i, (keys, ei) in enumerate(edge_index_dict.items()):
        key2int[keys] = i
        off = torch.tensor([[offset[keys[0]]], [offset[keys[-1]]]], device=ei.device)
        eis.append(ei + off)
        ets.append(torch.full((ei.size(1),), i, dtype=torch.long, device=ei.device))
    return (torch.cat(eis, 1), torch.cat(ets), torch.cat(node_types), torch.cat(local_idx), local2global, key2int)
