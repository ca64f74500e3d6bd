from .aggr import (Aggregation, FusedAggregation, MaxAggregation, MeanAggregation, MinAggregation,  # noqa: F401
                   MultiAggregation, SoftmaxAggregation, StdAggregation, SumAggregation, VarAggregation,
                   aggregation_resolver)
from .conv import (FastRGCNConv, GATConv, GATv2Conv, GCNConv, GINConv, GraphConv, HeteroLinear, RGCNConv,  # noqa: F401
                   SAGEConv, TransformerConv)
