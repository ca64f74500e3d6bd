from .aggr import (Aggregation, FusedAggregation, MaxAggregation, MeanAggregation, MinAggregation,  # noqa: F401
                   MultiAggregation, SoftmaxAggregation, StdAggregation, SumAggregation, VarAggregation,
                   aggregation_resolver)
from .conv import GATConv, GCNConv, GINConv, RGCNConv, SAGEConv  # noqa: F401
