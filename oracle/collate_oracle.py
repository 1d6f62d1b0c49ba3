"""ORACLE — test infrastructure only, never imported by the product path.

numpy restatement of batch collation and of the derived CSR index arrays.

The collation itself is third-party code the reference calls
(``torch_geometric`` — un-pinned in ``requirements.txt:5``, resolved against the
torch-2.5.1 wheel index at ``requirements.txt:16`` => PyG 2.6.x; absent from
``/root/reference`` and from this image).  Its published algorithm
(``Batch.from_data_list`` / ``collate.py``) is restated here and anchored on the
reference's call sites: ``cgr_mpnn_3D/training/trainer.py:105-118``,
``test.py:85-90``, with the per-graph field contract of
``cgr_mpnn_3D/data/ChemDataset.py:81-94`` and the edge emission order of
``cgr_mpnn_3D/utils/graph_features.py:184-195``.

Parity pinning: "parity unpinned" by reference tests (they only pin the count
relation ``len(edge_index) == len(f_bonds) == 2*bonds``,
``tests/test_molgraph.py:39-58``); that relation is asserted in
``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np


def collate_indices(num_nodes: Sequence[int], edge_indices: Sequence[np.ndarray]) -> Dict[str, np.ndarray]:
    """PyG collate of the integer fields.

    ``edge_index = cat_1(edge_index_g + offset_g)``, ``offset_g = sum_{g'<g} n_g'``;
    ``batch = repeat_interleave(arange(B), n_g)``; ``ptr = [0, cumsum(n_g)]``.
    """
    n = np.asarray(num_nodes, dtype=np.int64)
    ptr = np.zeros(n.size + 1, dtype=np.int64)
    np.cumsum(n, out=ptr[1:])
    ei = [np.asarray(e, dtype=np.int64) + ptr[g] for g, e in enumerate(edge_indices)]
    edge_index = np.concatenate(ei, axis=1) if ei else np.zeros((2, 0), np.int64)
    batch = np.repeat(np.arange(n.size, dtype=np.int64), n)
    e = np.array([x.shape[1] for x in edge_indices], dtype=np.int64)
    eptr = np.zeros(n.size + 1, dtype=np.int64)
    np.cumsum(e, out=eptr[1:])
    return {"edge_index": edge_index, "batch": batch, "ptr": ptr, "edge_ptr": eptr}


def csr_arrays(edge_index: np.ndarray, num_nodes: int) -> Dict[str, np.ndarray]:
    """chemprop-style index arrays derived from the batched ``edge_index``.

    * ``b2a   = edge_index[0]``              (reference GNN.py:85,132: ``row``)
    * ``a2b``  = CSR (``in_ptr``, ``in_idx``) of bonds grouped by ``edge_index[1]`` — the
      rows ``propagate`` sums (GNN.py:134) — ascending bond id inside a group, which is the
      order CPU ``scatter_add_`` accumulates in.
    * ``b2revb[e] = e ^ 1``                  (GNN.py:136-138 ``view(E//2,2,-1).flip(1)``)
    """
    src = edge_index[0].astype(np.int32)
    dst = edge_index[1].astype(np.int32)
    e = src.size
    order = np.argsort(dst, kind="stable").astype(np.int32)
    counts = np.bincount(dst, minlength=num_nodes).astype(np.int64)
    in_ptr = np.zeros(num_nodes + 1, dtype=np.int32)
    np.cumsum(counts, out=in_ptr[1:])
    rev = (np.arange(e, dtype=np.int32) ^ 1).astype(np.int32)
    return {"src": src, "dst": dst, "in_ptr": in_ptr, "in_idx": order, "rev": rev}


def check_pairing(edge_index: np.ndarray) -> bool:
    """Reference precondition (GNN.py:136-138): E even and bond ``e^1`` is the reverse of ``e``."""
    e = edge_index.shape[1]
    if e % 2:
        return False
    return bool(np.all(edge_index[0, 0::2] == edge_index[1, 1::2]) and
                np.all(edge_index[1, 0::2] == edge_index[0, 1::2]))


def tile_plan(ptr: np.ndarray, edge_ptr: np.ndarray, tile_rows: int = 128) -> Dict[str, np.ndarray]:
    """Greedy packing of consecutive whole reactions into row tiles of ``tile_rows`` bonds.

    Reaction g goes to the current tile if its bonds AND atoms still fit, else a new tile
    is opened.  A reaction with more than ``tile_rows`` bonds gets ``tile = -1`` (not
    tileable; the layer-wise path is used).  Returns per-reaction tile id and first row /
    first atom slot inside the tile, plus per-tile reaction ranges.
    """
    b = ptr.size - 1
    tile_of = np.full(b, -1, dtype=np.int32)
    row0 = np.zeros(b, dtype=np.int32)
    atom0 = np.zeros(b, dtype=np.int32)
    t = -1
    used_r = tile_rows + 1
    used_a = tile_rows + 1
    tile_first = []
    for g in range(b):
        ne = int(edge_ptr[g + 1] - edge_ptr[g])
        na = int(ptr[g + 1] - ptr[g])
        if ne > tile_rows or na > tile_rows:
            return {"ok": np.array(0), "tile_of": tile_of, "row0": row0, "atom0": atom0,
                    "tile_first": np.zeros(1, np.int32)}
        if used_r + ne > tile_rows or used_a + na > tile_rows:
            t += 1
            used_r = 0
            used_a = 0
            tile_first.append(g)
        tile_of[g] = t
        row0[g] = used_r
        atom0[g] = used_a
        used_r += ne
        used_a += na
    tile_first.append(b)
    return {"ok": np.array(1), "tile_of": tile_of, "row0": row0, "atom0": atom0,
            "tile_first": np.array(tile_first, dtype=np.int32)}


def gather_batch(graphs, indices):
    """What ``ReactionStore.batch(indices)`` must return: ChemDataset.__getitem__ (data/ChemDataset.py:69-94) for every
    index followed by PyG collate (training/trainer.py:105-118) -- cat on dim 0, edge_index shifted by the cumulative
    atom count, batch vector, ptr.  ``graphs`` carry ``x``, ``edge_index`` (local ids), ``edge_attr``, ``y``."""
    sel = [graphs[int(i)] for i in indices]
    n = np.array([g.x.shape[0] for g in sel], dtype=np.int64)
    ptr = np.zeros(len(sel) + 1, dtype=np.int64)
    np.cumsum(n, out=ptr[1:])
    return {
        "x": np.concatenate([np.asarray(g.x, dtype=np.float32) for g in sel], axis=0),
        "edge_attr": np.concatenate([np.asarray(g.edge_attr, dtype=np.float32) for g in sel], axis=0),
        "edge_index": np.concatenate([np.asarray(g.edge_index, dtype=np.int64) + ptr[i] for i, g in enumerate(sel)], axis=1),
        "batch": np.repeat(np.arange(len(sel), dtype=np.int64), n),
        "ptr": ptr,
        "y": np.concatenate([np.asarray(g.y, dtype=np.float32).reshape(-1)[:1] for g in sel]),
    }
