"""`from graphs.models.decentralplanner import *` (agents/decentralplannerlocal.py:27 of the
reference) -> the B200 DecentralPlannerNet.  Like the reference module (which defines no
`__all__`), the star-import also carries the module-level names of decentralplanner.py:4-10."""
import numpy as np  # noqa: F401
import torch  # noqa: F401
import torch.nn as nn  # noqa: F401
import torch.nn.functional as F  # noqa: F401

from gnn_pathplanning_b200.dropin import graphML as gml  # noqa: F401
from gnn_pathplanning_b200.planner import DecentralPlannerNet, weights_init  # noqa: F401
