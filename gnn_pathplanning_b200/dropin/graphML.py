"""`import utils.graphUtils.graphML as gml` (graphs/models/decentralplanner.py:9 of the
reference) -> the B200 GraphFilterBatch / BatchLSIGF.  The names the planner path constructs plus
the recurrent layers built on the same primitive; the rest of the Alelab layer zoo is
out of scope."""
from gnn_pathplanning_b200.graphml import (BatchLSIGF, GraphFilterBatch, GraphFilterL2ShareBatch,  # noqa: F401
                                           GraphFilterMoRNNBatch, GraphFilterRNNBatch, torchpermul)

__all__ = ["GraphFilterBatch", "BatchLSIGF", "GraphFilterRNNBatch", "GraphFilterMoRNNBatch", "GraphFilterL2ShareBatch",
           "torchpermul"]
