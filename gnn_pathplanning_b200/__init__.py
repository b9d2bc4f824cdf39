"""gnn_pathplanning_b200 -- B200 (sm_100a) implementation of the DecentralPlannerNet
forward path of proroklab/gnn_pathplanning behind the reference's own module API.

    from gnn_pathplanning_b200 import DecentralPlannerNet, GraphFilterBatch, BatchLSIGF

Importing the package never touches the GPU; the CUDA library (libgnnpp_b200.so, C ABI
in include/gnnpp_b200.h) is loaded on first use and every compute entry point raises if
it is missing -- there is no CPU fallback.  `install_dropin()` makes the reference's
dotted names (`graphs.models.decentralplanner`, `utils.graphUtils.graphML`,
`graphs.weights_initializer`) resolve to this package.
"""
from . import _lib
from .graphml import BatchLSIGF, GraphFilterBatch, graph_filter, FEATURE_MAJOR, NODE_MAJOR
from .planner import DecentralPlannerNet, weights_init

__all__ = ["DecentralPlannerNet", "GraphFilterBatch", "BatchLSIGF", "graph_filter", "weights_init",
           "FEATURE_MAJOR", "NODE_MAJOR", "install_dropin", "build"]


def build(force: bool = False, verbose: bool = False) -> str:
    return _lib.build(force=force, verbose=verbose)


def install_dropin() -> None:
    """Prepends gnn_pathplanning_b200/dropin to sys.path so that
    `from graphs.models.decentralplanner import *` and
    `import utils.graphUtils.graphML as gml` (agents/decentralplannerlocal.py:27,
    decentralplanner.py:9) pick up the B200 implementations."""
    import os
    import sys
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dropin")
    if d not in sys.path:
        sys.path.insert(0, d)
