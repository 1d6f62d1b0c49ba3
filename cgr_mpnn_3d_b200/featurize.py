"""CGR featurisation downstream of a SMILES parser.

The reference featurises every reaction with RDKit inside ``RxnGraph`` (``cgr_mpnn_3D/utils/graph_features.py:154-195``)
and does it again every epoch in its DataLoader workers.  RDKit is what turns a SMILES string into atoms, bonds and atom
maps; everything after it is table lookups -- the one-hot lists of ``atom_features`` / ``bond_features`` (``:4-63``), the
reactant ‖ (product − reactant) layout (``:177-195``), the union of reactant and product bonds and the edge order
``(a1, a2), (a2, a1)`` for ``a1 < a2`` ascending (``:184-195``).  This module takes the parser's output as compact integer
attribute arrays (:class:`ParsedMol`), does the integer part on the host (atom-map alignment, bond union and order) and
expands the float features on the GPU (``cgr_featurize_cgr``), straight into a collated batch whose ``x`` can carry the
MACE block next to the 78 CGR columns (``data/ChemDataset.py:83-86``).  No CPU implementation of the expansion exists
here: the oracle (``oracle/featurize_oracle.py``) is the CPU restatement the tests compare against.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .data import Batch

ATOM_FDIM, BOND_FDIM = 78, 14        # reference graph_features.py: 39 + 39, 7 + 7

# the reference's choice lists (graph_features.py:16-30) as the parser-side codes of :class:`ParsedMol`
SYMBOL_Z = [1, 6, 7, 8, 9, 14, 15, 16, 17, 35, 53]     # H C N O F Si P S Cl Br I
DEGREES = [0, 1, 2, 3, 4, 5]
CHARGES = [-1, -2, 1, 2, 0]
NUM_HS = [0, 1, 2, 3, 4]
HYB_CODES = {"SP": 0, "SP2": 1, "SP3": 2, "SP3D": 3, "SP3D2": 4}       # anything else: 5 ("unknown" slot)
BOND_CODES = {"SINGLE": 0, "DOUBLE": 1, "TRIPLE": 2, "AROMATIC": 3}     # anything else: 4 (a bond of another type)
_Z_OF = {"H": 1, "C": 6, "N": 7, "O": 8, "F": 9, "Si": 14, "P": 15, "S": 16, "Cl": 17, "Br": 35, "I": 53}


def default_tables() -> _lib.CgrFeatureTables:
    t = _lib.CgrFeatureTables()
    t.symbol_z[:] = SYMBOL_Z
    t.degrees[:] = DEGREES
    t.charges[:] = CHARGES
    t.num_hs[:] = NUM_HS
    t.hybridizations[:] = [0, 1, 2, 3, 4]
    return t


@dataclass
class ParsedMol:
    """What a SMILES parser hands over for one side of a reaction (explicit hydrogens kept, graph_features.py:106-118)."""
    attrs: np.ndarray      # [n, 6] int16: atomic number, total degree, formal charge, total #Hs, hybridisation code, aromatic
    mass: np.ndarray       # [n] float64 (atom.GetMass())
    map_num: np.ndarray    # [n] int64 atom map numbers
    bonds: np.ndarray      # [m, 2] int64 atom index pairs, a < b
    bond_attrs: np.ndarray  # [m, 3] int8: type code, conjugated, in ring

    @classmethod
    def from_records(cls, mol: dict) -> "ParsedMol":
        """From the record form the oracle uses (``oracle/featurize_oracle.py``): symbols / names instead of codes."""
        atoms = mol["atoms"]
        attrs = np.array([[_Z_OF.get(a[0], 0), a[1], a[2], int(a[3]), HYB_CODES.get(a[4], 5), 1 if a[5] else 0]
                          for a in atoms], dtype=np.int16).reshape(-1, 6)
        pairs = sorted(mol["bonds"])
        return cls(attrs=attrs, mass=np.array([a[6] for a in atoms], dtype=np.float64),
                   map_num=np.array([a[7] for a in atoms], dtype=np.int64),
                   bonds=np.array(pairs, dtype=np.int64).reshape(-1, 2),
                   bond_attrs=np.array([[BOND_CODES.get(mol["bonds"][p][0], 4), 1 if mol["bonds"][p][1] else 0,
                                         1 if mol["bonds"][p][2] else 0] for p in pairs], dtype=np.int8).reshape(-1, 3))


def _align(reac: ParsedMol, prod: Optional[ParsedMol]):
    """Integer part of ``RxnGraph`` (graph_features.py:170-195) / ``MolGraph`` (:138-151): product attributes in reactant
    atom order, the bond union in the reference's edge order, per-side bond codes (-1: absent on that side)."""
    n = reac.attrs.shape[0]
    if prod is None:                       # plain molecule graph: features of one side only
        key_r = reac.bonds[:, 0] * n + reac.bonds[:, 1]
        order = np.argsort(key_r, kind="stable")
        pairs = reac.bonds[order]
        return None, None, pairs, reac.bond_attrs[order], None
    # ri2pi (graph_features.py:83-103): a later product atom with the same map number wins, like the reference's dict
    prod_of_map = {int(m): i for i, m in enumerate(prod.map_num)}
    ri2pi = np.array([prod_of_map[int(m)] for m in reac.map_num], dtype=np.int64)
    pi2ri = {}
    for r, p in enumerate(ri2pi):
        pi2ri.setdefault(int(p), []).append(r)
    key_r = {int(a) * n + int(b): i for i, (a, b) in enumerate(reac.bonds)}
    # product bonds expressed in reactant indices: the reference looks up GetBondBetweenAtoms(ri2pi[a1], ri2pi[a2])
    key_p = {}
    for i, (a, b) in enumerate(prod.bonds):
        for ra in pi2ri.get(int(a), ()):
            for rb in pi2ri.get(int(b), ()):
                if ra != rb:
                    lo, hi = (ra, rb) if ra < rb else (rb, ra)
                    key_p[lo * n + hi] = i
    keys = np.array(sorted(set(key_r) | set(key_p)), dtype=np.int64)
    pairs = np.stack([keys // n, keys % n], axis=1) if keys.size else np.zeros((0, 2), dtype=np.int64)
    none = np.array([-1, 0, 0], dtype=np.int8)
    b_r = np.array([reac.bond_attrs[key_r[k]] if k in key_r else none for k in keys.tolist()], dtype=np.int8).reshape(-1, 3)
    b_p = np.array([prod.bond_attrs[key_p[k]] if k in key_p else none for k in keys.tolist()], dtype=np.int8).reshape(-1, 3)
    return prod.attrs[ri2pi], prod.mass[ri2pi], pairs, b_r, b_p


def featurize_batch(reactions: Sequence[Tuple[ParsedMol, Optional[ParsedMol]]], device="cuda",
                    mace: Optional[Sequence[np.ndarray]] = None, labels: Optional[Sequence[float]] = None,
                    tables: Optional[_lib.CgrFeatureTables] = None) -> Batch:
    """Featurise and collate reactions ``(reactant, product)`` (``product=None``: a plain molecule graph, the reference's
    ``mode="mol"``: 39 / 7 columns) into one device batch: ``x [N, 78 (+F3D)]``, ``edge_attr [E, 14]``, ``edge_index``
    with cumulative node offsets, ``batch``, ``ptr``, ``y``.  ``mace[i]`` is the reaction's ``[n_atoms, F3D]`` descriptor
    block (``arr_{i}`` of the reference's ``.npz``), copied next to the CGR columns in float32."""
    lib = _lib.load()
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("featurize_batch expands the features on the GPU: pass a CUDA device (no CPU path)")
    mol_mode = [p is None for _, p in reactions]
    if any(mol_mode) and not all(mol_mode):
        raise ValueError("mix of reaction and molecule graphs in one batch")
    if all(mol_mode) and reactions:
        raise NotImplementedError("mode='mol' (39 / 7 feature columns) is not used by the CGR model path")
    ar, ap, mr, mp, br, bp, ei, nn = [], [], [], [], [], [], [], []
    off = 0
    for reac, prod in reactions:
        a_p, m_p, pairs, b_r, b_p = _align(reac, prod)
        n = reac.attrs.shape[0]
        ar.append(reac.attrs); ap.append(a_p); mr.append(reac.mass); mp.append(m_p)
        br.append(np.repeat(b_r, 2, axis=0)); bp.append(np.repeat(b_p, 2, axis=0))     # both directions: same features
        e = np.empty((2, 2 * pairs.shape[0]), dtype=np.int64)
        e[0, 0::2], e[1, 0::2] = pairs[:, 0], pairs[:, 1]
        e[0, 1::2], e[1, 1::2] = pairs[:, 1], pairs[:, 0]
        ei.append(e + off)
        nn.append(n)
        off += n
    n_atoms = off
    cat = lambda xs, dt, w: (np.ascontiguousarray(np.concatenate(xs, axis=0), dtype=dt) if xs else np.zeros((0, w), dtype=dt))
    ar, ap = cat(ar, np.int16, 6), cat(ap, np.int16, 6)
    mr, mp = np.ascontiguousarray(np.concatenate(mr)), np.ascontiguousarray(np.concatenate(mp))
    br, bp = cat(br, np.int8, 3), cat(bp, np.int8, 3)
    edge_index = np.concatenate(ei, axis=1) if ei else np.zeros((2, 0), dtype=np.int64)
    n_bonds = int(edge_index.shape[1])
    f3d = 0 if mace is None else int(np.asarray(mace[0]).shape[1])
    up = lambda a: torch.from_numpy(a).to(dev, non_blocking=True)
    with torch.cuda.device(dev):
        x = torch.empty((n_atoms, ATOM_FDIM + f3d), dtype=torch.float32, device=dev)
        ea = torch.empty((n_bonds, BOND_FDIM), dtype=torch.float32, device=dev)
        d = [up(a) for a in (ar, ap, mr, mp, br, bp)]
        t = tables or default_tables()
        _lib.check(lib.cgr_featurize_cgr(C.byref(t), d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(),
                                         n_atoms, d[4].data_ptr(), d[5].data_ptr(), n_bonds, x.data_ptr(), x.stride(0),
                                         ea.data_ptr(), _lib.current_stream_handle()), "cgr_featurize_cgr")
        if f3d:
            m = np.concatenate([np.asarray(a, dtype=np.float32) for a in mace], axis=0)     # float64 would promote x
            if m.shape[0] != n_atoms:
                raise ValueError("MACE descriptor rows do not match the atoms of the reactions")
            x[:, ATOM_FDIM:] = up(np.ascontiguousarray(m))
        nn_a = np.array(nn, dtype=np.int64)
        ptr = np.zeros(len(nn) + 1, dtype=np.int64)
        np.cumsum(nn_a, out=ptr[1:])
        y = None if labels is None else up(np.asarray(labels, dtype=np.float32))
        return Batch(x, up(edge_index), ea, up(np.repeat(np.arange(len(nn), dtype=np.int64), nn_a)), up(ptr), y)
