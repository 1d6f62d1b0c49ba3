"""ORACLE — test infrastructure only, never imported by the product path.

CPU restatement (pure-Python loops, small cases only) of the reference's CGR featurisation
``cgr_mpnn_3D/utils/graph_features.py`` on PRE-PARSED molecules: RDKit (absent here) is what turns a SMILES string
into atoms and bonds; everything after that -- the one-hot tables, the reactant ‖ (product − reactant) layout, the
union of reactant and product bonds and the edge order -- is restated here line by line.

A parsed molecule is ``{"atoms": [(symbol, total_degree, formal_charge, total_num_hs, hybridization, is_aromatic, mass,
atom_map_num), ...], "bonds": {(a, b): (bond_type, is_conjugated, is_in_ring)}}`` with ``a < b`` atom indices,
``hybridization`` in {"SP", "SP2", "SP3", "SP3D", "SP3D2", other} and ``bond_type`` in {"SINGLE", "DOUBLE", "TRIPLE",
"AROMATIC", other}.

Pinning status: "parity unpinned" against the reference's own output (it needs RDKit); pinned instead against the
hand-derived vectors of SURVEY.md Appendix A in ``tests/golden/cgr_features_ethanol.json`` (the reference's own tests,
``tests/test_molgraph.py:39-58``, only pin counts: they are reproduced in tests/test_featurize.py).
"""
from __future__ import annotations

SYMBOLS = ["H", "C", "N", "O", "F", "Si", "P", "S", "Cl", "Br", "I"]          # graph_features.py:16-18
DEGREES = [0, 1, 2, 3, 4, 5]                                                  # :19
CHARGES = [-1, -2, 1, 2, 0]                                                   # :20
NUM_HS = [0, 1, 2, 3, 4]                                                      # :21
HYBRIDIZATIONS = ["SP", "SP2", "SP3", "SP3D", "SP3D2"]                        # :22-30
BOND_FDIM = 7                                                                 # :48


def onek_encoding_unk(value, choices):
    """graph_features.py:66-80: unknown values light the LAST slot."""
    encoding = [0] * (len(choices) + 1)
    index = choices.index(value) if value in choices else -1
    encoding[index] = 1
    return encoding


def atom_features(atom):
    """graph_features.py:4-35 on a parsed atom tuple."""
    symbol, degree, charge, num_hs, hyb, aromatic, mass, _ = atom
    return (onek_encoding_unk(symbol, SYMBOLS) + onek_encoding_unk(degree, DEGREES) + onek_encoding_unk(charge, CHARGES)
            + onek_encoding_unk(int(num_hs), NUM_HS) + onek_encoding_unk(hyb, HYBRIDIZATIONS)
            + [1 if aromatic else 0] + [mass * 0.01])


def bond_features(bond):
    """graph_features.py:38-63 on a parsed bond tuple (or None)."""
    if bond is None:
        return [1] + [0] * (BOND_FDIM - 1)
    bt, conj, ring = bond
    return [0, bt == "SINGLE", bt == "DOUBLE", bt == "TRIPLE", bt == "AROMATIC", (conj if bt is not None else 0),
            (ring if bt is not None else 0)]


def map_reac_to_prod(mol_reac, mol_prod):
    """graph_features.py:83-103."""
    prod_map_to_id = dict([(atom[7], i) for i, atom in enumerate(mol_prod["atoms"])])
    return dict([(i, prod_map_to_id[atom[7]]) for i, atom in enumerate(mol_reac["atoms"])])


def _bond(mol, a, b):
    return mol["bonds"].get((a, b) if a < b else (b, a))


def mol_graph(mol):
    """MolGraph.__init__, graph_features.py:126-151."""
    f_atoms, f_bonds, edge_index = [], [], []
    n = len(mol["atoms"])
    for a1 in range(n):
        f_atoms.append(atom_features(mol["atoms"][a1]))
        for a2 in range(a1 + 1, n):
            bond = _bond(mol, a1, a2)
            if bond is None:
                continue
            f_bond = bond_features(bond)
            f_bonds.append(f_bond)
            f_bonds.append(f_bond)
            edge_index.extend([(a1, a2), (a2, a1)])
    return f_atoms, f_bonds, edge_index


def rxn_graph(mol_reac, mol_prod):
    """RxnGraph.__init__, graph_features.py:159-195."""
    f_atoms, f_bonds, edge_index = [], [], []
    ri2pi = map_reac_to_prod(mol_reac, mol_prod)
    n = len(mol_reac["atoms"])
    for a1 in range(n):
        f_r = atom_features(mol_reac["atoms"][a1])
        f_p = atom_features(mol_prod["atoms"][ri2pi[a1]])
        f_atoms.append(f_r + [y - x for x, y in zip(f_r, f_p)])                      # :178-182
        for a2 in range(a1 + 1, n):
            b_r = _bond(mol_reac, a1, a2)
            b_p = _bond(mol_prod, ri2pi[a1], ri2pi[a2])
            if b_r is None and b_p is None:                                            # :187-188
                continue
            fb_r, fb_p = bond_features(b_r), bond_features(b_p)
            f_bond = fb_r + [y - x for x, y in zip(fb_r, fb_p)]                        # :189-192
            f_bonds.append(f_bond)
            f_bonds.append(f_bond)                                                     # :193-194
            edge_index.extend([(a1, a2), (a2, a1)])                                    # :195
    return f_atoms, f_bonds, edge_index
