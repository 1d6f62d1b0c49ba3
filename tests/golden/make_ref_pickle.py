"""Generate ``ref_module_small.pth``: a WHOLE-MODULE pickle written by the unmodified reference model, the way
``training/trainer.py:208`` saves checkpoints, plus the outputs it produces (``ref_module_small.npz``).

Build-container only (needs /root/reference).  The reference module is imported under its real name
``cgr_mpnn_3D.models.GNN`` so the pickle names the classes exactly as a checkpoint from a real training run does.  A
real run also leaves torch_geometric objects inside every DMPNNConv (MessagePassing internals); they are emulated by
attaching instances of classes from a fake ``torch_geometric.inspector`` module, so the loader's stand-in logic is
exercised by the fixture."""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests", "_pyg_shim"))
sys.path.insert(0, ROOT)
from cgr_mpnn_3d_b200.data import make_batch  # noqa: E402

REF = "/root/reference/cgr_mpnn_3D/models/GNN.py"


def main():
    for name in ("cgr_mpnn_3D", "cgr_mpnn_3D.models"):          # parent packages of the reference module
        sys.modules[name] = types.ModuleType(name)
    spec = importlib.util.spec_from_file_location("cgr_mpnn_3D.models.GNN", REF)
    ref = importlib.util.module_from_spec(spec)
    sys.modules["cgr_mpnn_3D.models.GNN"] = ref
    spec.loader.exec_module(ref)
    insp = types.ModuleType("torch_geometric.inspector")

    class Inspector:                                              # stands for MessagePassing.inspector of real PyG
        def __init__(self, tag):
            self.tag, self.params = tag, {"message": ["edge_attr"], "aggregate": ["index", "dim_size"]}
    Inspector.__module__ = "torch_geometric.inspector"
    Inspector.__qualname__ = "Inspector"
    insp.Inspector = Inspector
    sys.modules["torch_geometric.inspector"] = insp

    torch.manual_seed(21)
    torch.set_num_threads(1)
    model = ref.GNN(78, 14, depth=2, hidden_sizes=[32, 32], dropout_ps=[0.1, 0.2], use_learnable_skip=True)
    with torch.no_grad():
        for l, p in enumerate(model.skip_weights):
            p.fill_(0.9 - 0.2 * l)
    for l, conv in enumerate(model.convs):
        conv.inspector = Inspector(f"conv{l}")
        conv._explain = None
    data = make_batch(4, seed=9, kind="t1x", fa=78)
    model.eval()
    with torch.no_grad():
        out = model(data)
    torch.save(model, os.path.join(HERE, "ref_module_small.pth"))                # trainer.py:208
    arrs = {"out": out.numpy(), "dseed": np.int64(9), "nb": np.int64(4)}
    for k, v in model.state_dict().items():
        arrs["p/" + k] = v.numpy()
    np.savez(os.path.join(HERE, "ref_module_small.npz"), **arrs)
    print("wrote ref_module_small.pth / .npz; out =", out.numpy())


if __name__ == "__main__":
    main()
