from efficient_gnns_b200.nn import GCNConv, SAGEConv, MessagePassing  # noqa: F401
