"""B200-native implementation of the CGR-MPNN-3D hot path (encoder + readout + collate).

Public surface: :class:`GNN`, :class:`DMPNNConv` (reference module API), :func:`collate` /
:func:`build_plan` (device collation), the synthetic data helpers in :mod:`.data`.
"""
from .data import Batch, Graph, make_batch, make_reactions  # noqa: F401


def __getattr__(name):
    if name in ("GNN", "DMPNNConv", "global_add_pool"):
        from . import model
        return getattr(model, name)
    if name in ("collate", "build_plan", "plan_for", "GraphPlan"):
        from . import collate as _c
        return getattr(_c, name)
    raise AttributeError(name)
