"""Synthetic CGR reactions and the duck-typed batch container the model consumes.

The reference featurises reactions with RDKit (absent here) into per-graph
``tg.data.Data{x, edge_index, edge_attr, y}`` (reference
``cgr_mpnn_3D/data/ChemDataset.py:81-94``) whose directed bonds are emitted as
adjacent ``(a1,a2),(a2,a1)`` pairs sorted by ``a1 < a2``
(``cgr_mpnn_3D/utils/graph_features.py:184-195``).  This module produces graphs
with exactly that contract from a seeded numpy generator:

* T1x-shaped: ``n ~ U{8..23}`` atoms, random spanning tree + 2 extra bonds.
* drug-like: ``n ~ U{80..120}`` atoms, spanning tree + 4 extra bonds.

Nothing here touches the GPU; it is host-side input generation only.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch

FA_CGR = 78     # reference graph_features.py:178-182 (39 reactant + 39 diff)
FB_CGR = 14     # reference graph_features.py:186-192 (7 reactant + 7 diff)
F3D_DEFAULT = 768  # synthetic MACE block (runtime parameter)


@dataclass
class Graph:
    """One reaction, same fields as the reference's per-item ``Data``."""
    x: np.ndarray           # [n, Fa] float32
    edge_index: np.ndarray  # [2, e] int64
    edge_attr: np.ndarray   # [e, Fb] float32
    y: np.ndarray           # [1] float32

    @property
    def num_nodes(self) -> int:
        return int(self.x.shape[0])

    @property
    def num_edges(self) -> int:
        return int(self.edge_index.shape[1])


class Batch:
    """Duck-typed stand-in for ``torch_geometric.data.Batch``.

    Carries the attributes ``GNN.forward`` reads (reference
    ``cgr_mpnn_3D/models/GNN.py:77-82``) plus ``ptr``/``y``.  ``to()`` mirrors
    ``Batch.to(device)`` as used at ``cgr_mpnn_3D/training/trainer.py:139``.
    """

    _TENSOR_FIELDS = ("x", "edge_index", "edge_attr", "batch", "ptr", "y")

    def __init__(self, x, edge_index, edge_attr, batch=None, ptr=None, y=None):
        self.x = x
        self.edge_index = edge_index
        self.edge_attr = edge_attr
        self.batch = batch
        self.ptr = ptr
        self.y = y

    @property
    def num_graphs(self) -> int:
        if self.ptr is not None:
            return int(self.ptr.numel()) - 1
        if self.batch is None:
            return 1
        return int(self.batch.max()) + 1

    @property
    def num_nodes(self) -> int:
        return int(self.x.shape[0])

    @property
    def num_edges(self) -> int:
        return int(self.edge_index.shape[1])

    def to(self, device, non_blocking: bool = False) -> "Batch":
        kw = {}
        for f in self._TENSOR_FIELDS:
            t = getattr(self, f)
            kw[f] = None if t is None else t.to(device, non_blocking=non_blocking)
        return Batch(**kw)

    def pin_memory(self) -> "Batch":
        kw = {}
        for f in self._TENSOR_FIELDS:
            t = getattr(self, f)
            kw[f] = None if t is None else t.pin_memory()
        return Batch(**kw)

    def clone(self) -> "Batch":
        kw = {}
        for f in self._TENSOR_FIELDS:
            t = getattr(self, f)
            kw[f] = None if t is None else t.clone()
        return Batch(**kw)


def _random_bonds(rng: np.random.Generator, n: int, n_extra: int) -> np.ndarray:
    """Undirected bond list, lexicographically sorted rows (a1 < a2)."""
    pairs = set()
    for v in range(1, n):
        u = int(rng.integers(0, v))
        pairs.add((u, v))
    tries = 0
    added = 0
    while added < n_extra and tries < 64:
        tries += 1
        a, b = (int(t) for t in rng.integers(0, n, size=2))
        if a == b:
            continue
        p = (min(a, b), max(a, b))
        if p in pairs:
            continue
        pairs.add(p)
        added += 1
    return np.array(sorted(pairs), dtype=np.int64).reshape(-1, 2)


def make_graph(rng: np.random.Generator, n_lo: int, n_hi: int, n_extra: int,
               fa: int, fb: int) -> Graph:
    n = int(rng.integers(n_lo, n_hi + 1))
    bonds = _random_bonds(rng, n, n_extra)
    nb = bonds.shape[0]
    ei = np.empty((2, 2 * nb), dtype=np.int64)
    ei[0, 0::2] = bonds[:, 0]
    ei[1, 0::2] = bonds[:, 1]
    ei[0, 1::2] = bonds[:, 1]
    ei[1, 1::2] = bonds[:, 0]
    x = rng.standard_normal((n, fa), dtype=np.float32)
    if fa > FA_CGR:
        x[:, FA_CGR:] *= np.float32(0.1)   # "synthetic MACE fingerprints"
    ea_u = rng.standard_normal((nb, fb), dtype=np.float32)
    ea = np.repeat(ea_u, 2, axis=0)        # both directions share the row
    y = rng.standard_normal((1,), dtype=np.float32)
    return Graph(x=x, edge_index=ei, edge_attr=ea, y=y)


def make_reactions(num: int, seed: int = 0, kind: str = "t1x",
                   fa: int = FA_CGR + F3D_DEFAULT, fb: int = FB_CGR) -> List[Graph]:
    """``num`` synthetic reactions.  ``kind``: "t1x" | "drug"."""
    rng = np.random.default_rng(seed)
    if kind == "t1x":
        lo, hi, extra = 8, 23, 2
    elif kind == "drug":
        lo, hi, extra = 80, 120, 4
    else:
        raise ValueError(f"unknown kind {kind!r}")
    return [make_graph(rng, lo, hi, extra, fa, fb) for _ in range(num)]


def collate_host(graphs: List[Graph]) -> Batch:
    """Reference-semantics collate on the host with torch ops only.

    Mirrors PyG ``Batch.from_data_list`` as used by the reference loaders
    (``cgr_mpnn_3D/training/trainer.py:105-118``): cat features on dim 0,
    ``edge_index`` on dim 1 with cumulative node offsets, ``batch`` and ``ptr``.
    This is the plain host path used to *feed* the device collate kernel's
    parity tests; the product collate is ``cgr_mpnn_3d_b200.collate``.
    """
    n = np.array([g.num_nodes for g in graphs], dtype=np.int64)
    ptr = np.zeros(len(graphs) + 1, dtype=np.int64)
    np.cumsum(n, out=ptr[1:])
    x = np.concatenate([g.x for g in graphs], axis=0)
    ea = np.concatenate([g.edge_attr for g in graphs], axis=0)
    ei = np.concatenate([g.edge_index + ptr[i] for i, g in enumerate(graphs)], axis=1)
    batch = np.repeat(np.arange(len(graphs), dtype=np.int64), n)
    y = np.concatenate([g.y for g in graphs], axis=0)
    return Batch(
        x=torch.from_numpy(x), edge_index=torch.from_numpy(ei),
        edge_attr=torch.from_numpy(ea), batch=torch.from_numpy(batch),
        ptr=torch.from_numpy(ptr), y=torch.from_numpy(y),
    )


def make_batch(num: int, seed: int = 0, kind: str = "t1x",
               fa: int = FA_CGR + F3D_DEFAULT, fb: int = FB_CGR) -> Batch:
    return collate_host(make_reactions(num, seed, kind, fa, fb))
