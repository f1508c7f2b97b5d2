from efficient_gnns_b200.nn import GCNConv, SAGEConv, GATConv, MessagePassing  # noqa: F401
