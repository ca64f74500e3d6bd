from .aggr import (Aggregation, FusedAggregation, MaxAggregation, MeanAggregation, MinAggregation,  # noqa: F401
                   MultiAggregation, SoftmaxAggregation, StdAggregation, SumAggregation, VarAggregation,
                   aggregation_resolver)
from .conv import (FastRGCNConv, GATConv, GATv2Conv, GCNConv, GINConv, GraphConv, RGCNConv, SAGEConv,  # noqa: F401
                   TransformerConv)
