"""`from graphs.models.decentralplanner import *` (agents/decentralplannerlocal.py:27 of the
reference) -> the B200 DecentralPlannerNet."""
from gnn_pathplanning_b200.planner import DecentralPlannerNet  # noqa: F401

__all__ = ["DecentralPlannerNet"]
