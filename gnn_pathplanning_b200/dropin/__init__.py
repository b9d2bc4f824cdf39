"""Import hook behind `install_dropin()`.

The reference's callers reach the hot path through three dotted module names:

    graphs.models.decentralplanner   (agents/decentralplannerlocal.py:27: `from ... import *`)
    utils.graphUtils.graphML         (graphs/models/decentralplanner.py:9)
    graphs.weights_initializer       (graphs/models/decentralplanner.py:7)

A `sys.meta_path` finder answers exactly those three names with the modules of this package and
declines everything else, so the reference's own `graphs`, `graphs.losses`, `utils`, `utils.metrics`,
`utils.multirobotsim_dcenlocal` ... keep resolving from the reference tree on `sys.path` (its
package `__init__`s eagerly import every sibling module -- they get the overrides too).  Nothing is
shadowed and nothing is prepended to `sys.path`.
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.util
import sys

OVERRIDES = {
    "graphs.models.decentralplanner": "gnn_pathplanning_b200.dropin.decentralplanner",
    "utils.graphUtils.graphML": "gnn_pathplanning_b200.dropin.graphML",
    "graphs.weights_initializer": "gnn_pathplanning_b200.dropin.weights_initializer",
}


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, target: str):
        self.target = target

    def create_module(self, spec):
        return importlib.import_module(self.target)

    def exec_module(self, module):      # already executed under its own name
        return None


class DropinFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        tgt = OVERRIDES.get(fullname)
        if tgt is None:
            return None
        return importlib.util.spec_from_loader(fullname, _AliasLoader(tgt))


def install() -> None:
    if not any(isinstance(f, DropinFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, DropinFinder())
    # a reference module imported before install() keeps its place in sys.modules: replace it, and
    # re-point the attribute on its (already imported) parent package
    for name, tgt in OVERRIDES.items():
        old = sys.modules.get(name)
        mod = importlib.import_module(tgt)
        if old is not None and old is not mod:
            sys.modules[name] = mod
            parent, _, leaf = name.rpartition(".")
            if parent in sys.modules:
                setattr(sys.modules[parent], leaf, mod)


def uninstall() -> None:
    sys.meta_path[:] = [f for f in sys.meta_path if not isinstance(f, DropinFinder)]
    for name in OVERRIDES:
        sys.modules.pop(name, None)
