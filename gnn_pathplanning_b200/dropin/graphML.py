"""`import utils.graphUtils.graphML as gml` (graphs/models/decentralplanner.py:9 of the
reference) -> the B200 GraphFilterBatch / BatchLSIGF.  Only the two names the planner path
constructs are provided; the rest of the Alelab layer zoo is out of scope."""
from gnn_pathplanning_b200.graphml import BatchLSIGF, GraphFilterBatch  # noqa: F401

__all__ = ["GraphFilterBatch", "BatchLSIGF"]
