"""CPU: the C-ABI library loads, exports every symbol the header declares, and the host-side
module mirrors the reference's constructor / error behaviour.  No compute calls (no GPU here)."""
import os
import re

import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "cgr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cgr_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from cgr_mpnn_3d_b200 import _lib
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/cgr_b200.h but not exported"
        assert s in _lib.PROTOTYPES, f"{s} has no ctypes prototype"
    assert lib.cgr_version() == 100
    assert lib.cgr_csr_workspace(1000, 2000) > 0 and lib.cgr_collate_workspace(64) > 0


def test_argument_errors_are_reported_without_a_gpu():
    from cgr_mpnn_3d_b200 import _lib
    lib = _lib.load()
    rc = lib.cgr_csr_build(None, -1, 5, None, None, None, None, None, None, 0, None)
    assert rc < 0 and b"negative" in lib.cgr_last_error_string()
    with pytest.raises(RuntimeError):
        _lib.check(rc, "cgr_csr_build")


def test_constructor_matches_reference_signature():
    from cgr_mpnn_3D.models.GNN import GNN, DMPNNConv
    # reference tests/test_trainer.py:37-38 : positional, num_edge_features = 0
    m = GNN(5, 0)
    assert isinstance(m, torch.nn.Module) and m.depth == 3
    assert m.hidden_sizes == [300] * 3 and m.dropout_ps == [0.02] * 3
    assert m.edge_init.weight.shape == (300, 5)
    m = GNN(num_node_features=846, num_edge_features=14, depth=4, hidden_sizes=[400] * 4, dropout_ps=[0.1] * 4,
            activation_fn=F.relu, use_learnable_skip=True)
    keys = list(m.state_dict().keys())
    assert keys[:2] == ["edge_init.weight", "edge_init.bias"]
    assert keys[2:10] == [f"convs.{l}.lin.{k}" for l in range(4) for k in ("weight", "bias")]
    assert keys[10:14] == ["edge_to_node.weight", "edge_to_node.bias", "ffn.weight", "ffn.bias"]
    assert keys[14:] == [f"skip_weights.{l}" for l in range(4)]
    assert sum(p.numel() for p in m.parameters()) == 1485205
    assert isinstance(m.convs[0], DMPNNConv) and m.convs[0].lin.weight.shape == (400, 400)
    # README config (depth 4, three hidden sizes) raises IndexError exactly like GNN.py:59-60
    with pytest.raises(IndexError):
        GNN(846, 14, depth=4, hidden_sizes=[400, 400, 400])


def test_state_dict_roundtrip_with_oracle_layout():
    from cgr_mpnn_3D.models.GNN import GNN
    from oracle.gnn_oracle import OracleGNN
    o = OracleGNN(78, 14, depth=3, hidden_sizes=[64] * 3, use_learnable_skip=True)
    m = GNN(78, 14, depth=3, hidden_sizes=[64] * 3, use_learnable_skip=True)
    m.load_state_dict(o.state_dict())          # strict: identical keys and shapes
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), o.state_dict().values()))


def test_whole_module_pickle(tmp_path):
    """reference trainer.py:208 pickles the module; test.py:93 / CLI :62 unpickle it."""
    from cgr_mpnn_3D.models.GNN import GNN
    m = GNN(78, 14, depth=2, hidden_sizes=[32] * 2, use_learnable_skip=True)
    path = tmp_path / "model.pth"
    torch.save(m, path)
    m2 = torch.load(path, map_location="cpu", weights_only=False)
    assert type(m2).__module__ == "cgr_mpnn_3D.models.GNN"
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))


def test_no_cpu_fallback():
    from cgr_mpnn_3D.models.GNN import GNN
    from cgr_mpnn_3d_b200.data import make_batch
    if torch.cuda.is_available():
        pytest.skip("CPU-only behaviour")
    m = GNN(78, 14, depth=2, hidden_sizes=[32] * 2)
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        m(make_batch(2, seed=0, fa=78))


def test_fused_adam_has_torch_adam_interface_and_no_cpu_path():
    """FusedAdam mirrors torch.optim.Adam (train.py:117-119): same arguments, param_groups drive ExponentialLR
    (train.py:121), argument errors match; stepping CPU parameters fails loudly instead of falling back."""
    import pytest
    import torch
    from cgr_mpnn_3d_b200.optim import FusedAdam
    w = torch.nn.Parameter(torch.ones(4))
    opt = FusedAdam([w], lr=1e-2, weight_decay=1e-4, amsgrad=True)
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.5)
    assert set(opt.param_groups[0]) >= {"lr", "betas", "eps", "weight_decay", "amsgrad"}
    opt.step()                      # no gradients yet: nothing to do, like torch
    sched.step()
    assert opt.param_groups[0]["lr"] == pytest.approx(5e-3)
    w.grad = torch.ones(4)
    with pytest.raises(RuntimeError, match="CUDA"):
        opt.step()
    with pytest.raises(ValueError):
        FusedAdam([w], lr=-1.0)
    with pytest.raises(ValueError):
        FusedAdam([w], betas=(1.0, 0.999))


def test_reference_whole_module_checkpoint_loads_without_torch_geometric():
    """a checkpoint written by the reference (torch.save(self.model), trainer.py:208) -- classes
    named cgr_mpnn_3D.models.GNN.*, torch_geometric objects inside -- loads into the B200 module with no torch_geometric
    installed.  Fixture: tests/golden/ref_module_small.pth, written by the unmodified reference (make_ref_pickle.py)."""
    import os
    import sys
    import numpy as np
    import torch
    import torch.nn.functional as F
    from cgr_mpnn_3d_b200.checkpoint import load_reference_checkpoint
    from cgr_mpnn_3d_b200.model import GNN
    assert "torch_geometric" not in sys.modules or "_pyg_shim" in (getattr(sys.modules["torch_geometric"], "__file__", "") or "")
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    model = load_reference_checkpoint(os.path.join(here, "ref_module_small.pth"), map_location="cpu")
    z = np.load(os.path.join(here, "ref_module_small.npz"))
    assert type(model) is GNN and type(model).__module__ == "cgr_mpnn_3D.models.GNN"
    assert model.depth == 2 and list(model.hidden_sizes) == [32, 32] and list(model.dropout_ps) == [0.1, 0.2]
    assert model.activation_fn is F.relu and model.use_learnable_skip is True
    assert model.num_node_features == 78 and model.num_edge_features == 14
    sd = model.state_dict()
    keys = [k[2:] for k in z.files if k.startswith("p/")]
    assert list(sd.keys()) == keys
    for k in keys:
        assert torch.equal(sd[k], torch.from_numpy(z["p/" + k])), k
    # a bare state_dict file (model.state_dict() saved by a user) loads too; hyper-parameters come from the shapes
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "sd.pth")
        torch.save(sd, path)
        m2 = load_reference_checkpoint(path, dropout_ps=[0.1, 0.2])
        assert m2.depth == 2 and m2.use_learnable_skip and all(torch.equal(a, b) for a, b in zip(m2.state_dict().values(), sd.values()))


def test_reaction_store_is_device_only():
    """The packed reaction store lives in GPU memory; a CPU device is refused, not emulated."""
    import numpy as np
    import pytest
    from cgr_mpnn_3d_b200.data import make_reactions
    from cgr_mpnn_3d_b200.store import ReactionStore
    with pytest.raises(RuntimeError, match="CUDA"):
        ReactionStore.from_graphs(make_reactions(3, seed=0, kind="t1x", fa=8), device="cpu")


def test_host_tile_plan_matches_oracle():
    """cgr_tc_plan_host is pure host code (no device work): the greedy packing of whole reactions into 128-row tiles
    must equal the oracle's restatement, and untileable reactions are reported with -3."""
    import ctypes as C
    import numpy as np
    from cgr_mpnn_3d_b200 import _lib
    from oracle import collate_oracle
    lib = _lib.load()
    rng = np.random.default_rng(0)
    for trial in range(20):
        b = int(rng.integers(1, 200))
        na = rng.integers(2, 40, size=b)
        ne = 2 * rng.integers(1, 45, size=b)
        ptr = np.concatenate([[0], np.cumsum(na)]).astype(np.int64)
        eptr = np.concatenate([[0], np.cumsum(ne)]).astype(np.int64)
        tiles = np.zeros((b, 8), dtype=np.int32)
        n_tiles = C.c_int64(0)
        rc = lib.cgr_tc_plan_host(ptr.ctypes.data, eptr.ctypes.data, b, tiles.ctypes.data, C.byref(n_tiles))
        assert rc == 0
        tf = collate_oracle.tile_plan(ptr, eptr)["tile_first"]
        t = int(n_tiles.value)
        assert t == tf.size - 1
        assert np.array_equal(tiles[:t, 4], tf[:-1]) and np.array_equal(tiles[:t, 5], np.diff(tf))
        assert np.array_equal(tiles[:t, 0], eptr[tf[:-1]]) and np.array_equal(tiles[:t, 1], eptr[tf[1:]] - eptr[tf[:-1]])
        assert np.array_equal(tiles[:t, 2], ptr[tf[:-1]]) and np.array_equal(tiles[:t, 3], ptr[tf[1:]] - ptr[tf[:-1]])
        assert tiles[:t, 1].max() <= 128 and tiles[:t, 3].max() <= 128
    ptr = np.array([0, 10, 150], dtype=np.int64)          # second reaction: 140 atoms / 300 bonds
    eptr = np.array([0, 20, 320], dtype=np.int64)
    tiles = np.zeros((2, 8), dtype=np.int32)
    assert lib.cgr_tc_plan_host(ptr.ctypes.data, eptr.ctypes.data, 2, tiles.ctypes.data, C.byref(n_tiles)) == -3
    assert b"not tileable" in lib.cgr_last_error_string()


def test_store_pack_order_fills_tiles():
    """cgr_store_pack_order (pure host code): a permutation of the batch, under which the greedy plan needs fewer tiles
    than in the caller's order and comes close to the lower bound; both constraints (bonds, atoms <= 128) respected."""
    import ctypes as C
    import numpy as np
    from cgr_mpnn_3d_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(1)

    def n_tiles_of(na, ne):
        ptr = np.concatenate([[0], np.cumsum(na)]).astype(np.int64)
        eptr = np.concatenate([[0], np.cumsum(ne)]).astype(np.int64)
        tiles = np.zeros((na.size, 8), dtype=np.int32)
        t = C.c_int64(0)
        assert lib.cgr_tc_plan_host(ptr.ctypes.data, eptr.ctypes.data, na.size, tiles.ctypes.data, C.byref(t)) == 0
        assert tiles[: t.value, 1].max() <= 128 and tiles[: t.value, 3].max() <= 128
        return int(t.value)

    for kind, n_store, n in (("t1x", 5000, 2048), ("t1x", 300, 64), ("mixed", 4000, 1000), ("atoms", 500, 400)):
        if kind == "t1x":                                  # n ~ U{8..23} atoms, about n + 1 undirected bonds
            na = rng.integers(8, 24, size=n_store)
            ne = 2 * (na + rng.integers(0, 3, size=n_store))
        elif kind == "mixed":
            na = rng.integers(1, 60, size=n_store)
            ne = 2 * rng.integers(1, 65, size=n_store)
        else:                                              # the atom constraint binds: many atoms, few bonds
            na = rng.integers(30, 129, size=n_store)
            ne = 2 * rng.integers(1, 20, size=n_store)
        nptr = np.concatenate([[0], np.cumsum(na)]).astype(np.int64)
        eptr = np.concatenate([[0], np.cumsum(ne)]).astype(np.int64)
        ids = rng.integers(0, n_store, size=n).astype(np.int64)      # repeats allowed
        perm = np.full(n, -1, dtype=np.int32)
        assert lib.cgr_store_pack_order(nptr.ctypes.data, eptr.ctypes.data, n_store, ids.ctypes.data, n, perm.ctypes.data) == 0
        assert np.array_equal(np.sort(perm), np.arange(n))
        before = n_tiles_of(na[ids], ne[ids])
        after = n_tiles_of(na[ids[perm]], ne[ids[perm]])
        bound = max(int(np.ceil(ne[ids].sum() / 128)), int(np.ceil(na[ids].sum() / 128)))
        assert bound <= after <= before, (kind, bound, after, before)
        if kind == "t1x" and n >= 1024:
            assert after <= 0.92 * before and after <= 1.05 * bound, (bound, after, before)
    # a reaction beyond a tile: identity order and -3; ids out of range are refused
    nptr = np.array([0, 10, 150], dtype=np.int64)
    eptr = np.array([0, 20, 320], dtype=np.int64)
    ids = np.array([1, 0], dtype=np.int64)
    perm = np.zeros(2, dtype=np.int32)
    assert lib.cgr_store_pack_order(nptr.ctypes.data, eptr.ctypes.data, 2, ids.ctypes.data, 2, perm.ctypes.data) == -3
    assert perm.tolist() == [0, 1]
    ids[0] = 5
    assert lib.cgr_store_pack_order(nptr.ctypes.data, eptr.ctypes.data, 2, ids.ctypes.data, 2, perm.ctypes.data) < 0


def test_optimizer_entry_points_validate_arguments_without_a_gpu():
    import ctypes as C
    from cgr_mpnn_3d_b200 import _lib
    lib = _lib.load()
    arr = (_lib.CgrAdamTensor * 1)()
    assert lib.cgr_adam_step(None, 0, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1, 1, 1.0, None) < 0
    assert lib.cgr_adam_step(arr, 1, 1e-3, 0.9, 0.999, 1e-8, 0.0, 0, 1, 1.0, None) < 0          # steps count from 1
    assert lib.cgr_adam_step(arr, 1, 1e-3, 1.5, 0.999, 1e-8, 0.0, 1, 1, 1.0, None) < 0          # beta1 out of range
    assert b"range" in lib.cgr_last_error_string()
    assert lib.cgr_peer_allreduce_adam(arr, 1, None, None, None, 0, 2, 0, 1, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1, 1, 1.0, None) < 0
