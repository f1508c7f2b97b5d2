from efficient_gnns_b200.nn import GCNConv, SAGEConv, GATConv, GINConv, MessagePassing  # noqa: F401
