"""``cgr_mpnn_3D.models.GNN`` — the module path the reference's callers import
(``train.py:8``, ``tests/test_trainer.py:11``) and that reference ``.pth`` pickles name
(``cgr_mpnn_3D.models.GNN.GNN`` / ``.DMPNNConv``, ``training/trainer.py:208``).

The classes are the B200-native implementations from :mod:`cgr_mpnn_3d_b200.model`; their
``__module__`` is rewritten to this path so ``torch.save(model)`` produces files the reference's
``torch.load`` call sites (``test.py:93``, ``cli_tool/activation_energy_predictor.py:62``) resolve.
"""
from cgr_mpnn_3d_b200.model import GNN, DMPNNConv, global_add_pool  # noqa: F401

GNN.__module__ = __name__
DMPNNConv.__module__ = __name__
global_add_pool.__module__ = __name__
