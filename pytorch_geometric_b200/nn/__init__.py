from .aggr import (Aggregation, MaxAggregation, MeanAggregation, MinAggregation, SoftmaxAggregation,  # noqa: F401
                   SumAggregation, aggregation_resolver)
from .conv import GATConv, GCNConv, GINConv, RGCNConv, SAGEConv  # noqa: F401
