"""Generate the committed golden vectors by running the UNMODIFIED reference model.

Build-container only (needs ``/root/reference``; the GPU box never runs this):

    python tests/golden/make_golden.py

Imports ``/root/reference/cgr_mpnn_3D/models/GNN.py`` verbatim with the test-only
``tests/_pyg_shim`` standing in for torch_geometric, feeds it seeded synthetic CGR
batches (``cgr_mpnn_3d_b200.data``) and stores inputs, the full state_dict, forward
outputs, the MSE(sum) loss (reference ``train.py:120``) and every parameter gradient.
Small cases store everything; the two BASELINE-sized cases store only outputs, loss
and per-parameter gradient checksums, with weights regenerated from the seed.
"""
import importlib.util
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests", "_pyg_shim"))
sys.path.insert(0, ROOT)

from cgr_mpnn_3d_b200.data import make_batch, make_reactions, collate_host, Batch  # noqa: E402

REF = "/root/reference/cgr_mpnn_3D/models/GNN.py"
ACTS = {"relu": F.relu, "silu": F.silu, "gelu": F.gelu}


def load_reference():
    spec = importlib.util.spec_from_file_location("_reference_gnn", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


CASES = [
    # name, fa, depth, H, act, skip, B, data_seed, weight_seed, full, batch_none
    ("small_relu",      78, 3,  64, "relu", False, 5, 1, 11, True, False),
    ("small_skip",     110, 4,  48, "relu", True,  6, 2, 12, True, False),
    ("small_silu",      78, 2,  32, "silu", True,  4, 3, 13, True, False),
    ("small_gelu",      94, 2,  40, "gelu", False, 3, 4, 14, True, False),
    ("single_nobatch",  78, 3,  64, "relu", True,  1, 5, 15, True, True),
    ("cfg1_d3_h300",    78, 3, 300, "relu", False, 32, 1, 0, False, False),
    ("cfg2_d4_h400",   846, 4, 400, "relu", True,  64, 0, 0, False, False),
]


def main():
    ref = load_reference()
    torch.set_num_threads(1)   # sequential fp32 accumulation everywhere
    for (name, fa, depth, hid, act, skip, nb, dseed, wseed, full, nobatch) in CASES:
        data = make_batch(nb, seed=dseed, kind="t1x", fa=fa)
        if nobatch:
            data = Batch(data.x, data.edge_index, data.edge_attr, None, None, data.y)
        torch.manual_seed(wseed)
        model = ref.GNN(fa, 14, depth=depth, hidden_sizes=[hid] * depth,
                        dropout_ps=[0.0] * depth, activation_fn=ACTS[act],
                        use_learnable_skip=skip)
        if skip:
            with torch.no_grad():
                for l, p in enumerate(model.skip_weights):
                    p.fill_(1.0 - 0.15 * l + 0.05 * (l % 2))
        model.train()
        out = model(data)
        loss = ((out - data.y) ** 2).sum()
        loss.backward()
        rec = {
            "meta": np.array([fa, 14, depth, hid, int(skip), nb, dseed, wseed, int(nobatch)], np.int64),
            "act": np.array(act),
            "out": out.detach().numpy(),
            "loss": loss.detach().numpy(),
            "state_keys": np.array(list(model.state_dict().keys())),
        }
        for k, p in model.named_parameters():
            g = p.grad.detach().double()
            rec["gsum/" + k] = np.array([float(g.sum()), float(g.abs().sum()), float(g.abs().max())])
        if full:
            rec.update({"x": data.x.numpy(), "edge_index": data.edge_index.numpy(),
                        "edge_attr": data.edge_attr.numpy(), "y": data.y.numpy()})
            if not nobatch:
                rec["batch"] = data.batch.numpy()
                rec["ptr"] = data.ptr.numpy()
            for k, v in model.state_dict().items():
                rec["w/" + k] = v.detach().numpy()
            for k, p in model.named_parameters():
                rec["g/" + k] = p.grad.detach().numpy()
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **rec)
        print(f"{name}: out[:3]={out[:3].tolist()} loss={float(loss):.6f} -> {os.path.getsize(path)/1024:.1f} KiB")

    # reference error behaviours worth pinning
    try:
        ref.GNN(78, 14, depth=4, hidden_sizes=[400, 400, 400])
        raised = False
    except IndexError:
        raised = True
    print("README config raises IndexError:", raised)


if __name__ == "__main__":
    main()
