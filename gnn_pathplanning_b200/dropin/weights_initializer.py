"""`from graphs.weights_initializer import weights_init` -> the package's re-statement of the
reference rule (graphs/weights_initializer.py:11-23)."""
from gnn_pathplanning_b200.planner import weights_init  # noqa: F401

__all__ = ["weights_init"]
