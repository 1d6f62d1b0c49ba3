"""CGR featurisation: the oracle restatement of the reference's ``utils/graph_features.py`` against
hand-derived vectors, and the device featuriser (``cgr_featurize_cgr``) bit-exact against the oracle."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import featurize_oracle as fo
from tests.util import GOLDEN


def _mol(rec):
    return {"atoms": [tuple(a) for a in rec["atoms"]], "bonds": {(b[0], b[1]): (b[2], b[3], b[4]) for b in rec["bonds"]}}


def _fixture():
    with open(os.path.join(GOLDEN, "cgr_features_ethanol.json")) as fh:
        return json.load(fh)["reactions"]


def _f32(rows):
    return np.array(rows, dtype=np.float64).astype(np.float32)       # torch.tensor(list, dtype=torch.float), ChemDataset.py:82


def test_oracle_matches_hand_derived_vectors():
    for rx in _fixture():
        f_atoms, f_bonds, edge_index = fo.rxn_graph(_mol(rx["reac"]), _mol(rx["prod"]))
        assert len(f_atoms[0]) == 78 and len(f_bonds[0]) == 14                          # Fa = 78, Fb = 14
        assert np.array_equal(_f32(f_atoms), _f32(rx["f_atoms"]))
        assert np.array_equal(_f32(f_bonds), _f32(rx["f_bonds"]))
        assert [list(e) for e in edge_index] == rx["edge_index"]


def test_oracle_reproduces_the_reference_tests():
    """reference tests/test_molgraph.py:22-58 (counts) on the parsed form of the same molecules."""
    assert fo.onek_encoding_unk("A", ["A", "B", "C"]) == [1, 0, 0, 0] and fo.onek_encoding_unk("D", ["A", "B", "C"]) == [0, 0, 0, 1]
    f = fo.bond_features(("DOUBLE", False, False))
    assert len(f) == 7 and f[1] == 0 and f[2] == 1
    ethanol = _mol(_fixture()[0]["reac"])
    f_atoms, f_bonds, edge_index = fo.mol_graph(ethanol)                                # MolGraph("CCO")
    assert len(f_atoms) == 3 and len(f_bonds) == 4 and len(edge_index) == 4
    rx = _fixture()[0]
    f_atoms, f_bonds, edge_index = fo.rxn_graph(_mol(rx["reac"]), _mol(rx["prod"]))     # RxnGraph("CCO>>CC=O")
    assert len(f_atoms) == 3 and len(f_bonds) == 4 and len(edge_index) == 4 and f_atoms[0] != f_atoms[1]
    assert fo.map_reac_to_prod(_mol(_fixture()[1]["reac"]), _mol(_fixture()[1]["prod"])) == {0: 1, 1: 2, 2: 0}


def _random_mol(rng, n, maps):
    syms = ["H", "C", "N", "O", "F", "Si", "P", "S", "Cl", "Br", "I", "B", "Se"]
    hybs = ["SP", "SP2", "SP3", "SP3D", "SP3D2", "S", "OTHER"]
    atoms = [(syms[rng.integers(len(syms))], int(rng.integers(0, 8)), int(rng.integers(-3, 4)), int(rng.integers(0, 6)),
              hybs[rng.integers(len(hybs))], bool(rng.integers(2)), float(rng.uniform(1.0, 127.0)), int(m)) for m in maps]
    bonds = {}
    types = ["SINGLE", "DOUBLE", "TRIPLE", "AROMATIC", "DATIVE"]
    for v in range(1, n):
        bonds[(int(rng.integers(0, v)), v)] = (types[rng.integers(len(types))], bool(rng.integers(2)), bool(rng.integers(2)))
    for _ in range(3):
        a, b = sorted(int(t) for t in rng.integers(0, n, size=2))
        if a != b:
            bonds[(a, b)] = (types[rng.integers(len(types))], bool(rng.integers(2)), bool(rng.integers(2)))
    return {"atoms": atoms, "bonds": bonds}


@pytest.mark.gpu
def test_device_featuriser_bit_exact_against_oracle():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from cgr_mpnn_3d_b200.featurize import ParsedMol, featurize_batch
    rng = np.random.default_rng(0)
    recs = [(_mol(rx["reac"]), _mol(rx["prod"])) for rx in _fixture()]
    for _ in range(40):
        n = int(rng.integers(2, 24))
        maps = rng.permutation(n) + 1
        reac = _random_mol(rng, n, maps)
        prod = _random_mol(rng, n, rng.permutation(maps))          # product lists the same atoms in another order
        recs.append((reac, prod))
    f3d = 5
    mace = [rng.standard_normal((len(r["atoms"]), f3d)) for r, _ in recs]          # float64, like MACE descriptors
    labels = rng.standard_normal(len(recs)).astype(np.float32)
    batch = featurize_batch([(ParsedMol.from_records(r), ParsedMol.from_records(p)) for r, p in recs], mace=mace, labels=labels)
    xs, eas, eis, off = [], [], [], 0
    for (r, p), m in zip(recs, mace):
        f_atoms, f_bonds, edge_index = fo.rxn_graph(r, p)
        xs.append(np.concatenate([_f32(f_atoms), m.astype(np.float32)], axis=1))
        eas.append(_f32(f_bonds).reshape(-1, 14))
        eis.append(np.array(edge_index, dtype=np.int64).reshape(-1, 2).T + off)
        off += len(f_atoms)
    assert batch.x.dtype == torch.float32 and batch.x.shape[1] == 78 + f3d
    assert np.array_equal(batch.x.cpu().numpy(), np.concatenate(xs))                  # bit-exact
    assert np.array_equal(batch.edge_attr.cpu().numpy(), np.concatenate(eas))
    assert np.array_equal(batch.edge_index.cpu().numpy(), np.concatenate(eis, axis=1))
    assert np.array_equal(batch.ptr.cpu().numpy(), np.cumsum([0] + [len(r["atoms"]) for r, _ in recs]))
    assert np.array_equal(batch.y.cpu().numpy(), labels)
    # the batch is what the model consumes: pairing of directed bonds and the CSR build accept it
    from cgr_mpnn_3d_b200.collate import plan_for
    plan = plan_for(batch)
    assert (int(plan.status.item()) & 3) == 0        # reverse pairing / index range hold (isolated atoms may exist in random graphs)
