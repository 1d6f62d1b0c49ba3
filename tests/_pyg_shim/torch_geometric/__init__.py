"""TEST-ONLY stand-in for ``torch_geometric`` (absent from this image, not installable).

Provides just enough surface for the UNMODIFIED reference
``cgr_mpnn_3D/models/GNN.py`` and ``cgr_mpnn_3D/training/trainer.py`` to import and run:
``nn.MessagePassing`` with sum-aggregating ``propagate``, ``nn.global_add_pool``, ``loader.DataLoader``
and ``data.Data`` (semantics recalled from PyG 2.6).  Used by ``tests/golden/make_golden.py`` (build container
only) to generate the committed golden vectors.  Never shipped as product code.
"""
from . import data, loader, nn  # noqa: F401
