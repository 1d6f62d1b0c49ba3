"""Shared helpers for the tests (golden loading, oracle construction)."""
import os

import numpy as np
import torch
import torch.nn.functional as F

from cgr_mpnn_3d_b200.data import Batch, make_batch
from oracle.gnn_oracle import OracleGNN

ACTS = {"relu": F.relu, "silu": F.silu, "gelu": F.gelu}
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SMALL_CASES = ["small_relu", "small_skip", "small_silu", "small_gelu", "single_nobatch"]
BIG_CASES = ["cfg1_d3_h300", "cfg2_d4_h400"]


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    fa, fb, depth, hid, skip, nb, dseed, wseed, nobatch = (int(v) for v in z["meta"])
    meta = dict(fa=fa, fb=fb, depth=depth, hidden=hid, skip=bool(skip), nb=nb, dseed=dseed,
                wseed=wseed, nobatch=bool(nobatch), act=str(z["act"]))
    return z, meta


def case_batch(z, meta):
    if "x" in z.files:
        t = lambda k: torch.from_numpy(z[k])
        batch = None if meta["nobatch"] else t("batch")
        ptr = None if meta["nobatch"] else t("ptr")
        return Batch(t("x"), t("edge_index"), t("edge_attr"), batch, ptr, t("y"))
    return make_batch(meta["nb"], seed=meta["dseed"], kind="t1x", fa=meta["fa"])


def case_state_dict(z):
    return {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w/")}


def build_oracle(meta, state=None, dtype=torch.float32):
    torch.manual_seed(meta["wseed"])
    m = OracleGNN(meta["fa"], meta["fb"], depth=meta["depth"], hidden_sizes=[meta["hidden"]] * meta["depth"],
                  dropout_ps=[0.0] * meta["depth"], activation_fn=ACTS[meta["act"]],
                  use_learnable_skip=meta["skip"])
    if state is not None:
        m.load_state_dict(state)
    elif meta["skip"]:
        with torch.no_grad():
            for l, p in enumerate(m.skip_weights):
                p.fill_(1.0 - 0.15 * l + 0.05 * (l % 2))
    return m.to(dtype)


def build_model(meta, state=None, device="cuda", engine="simt", dropout_ps=None):
    """The product model (drop-in module path) with the same weights as ``build_oracle``."""
    from cgr_mpnn_3D.models.GNN import GNN
    torch.manual_seed(meta["wseed"])
    m = GNN(meta["fa"], meta["fb"], depth=meta["depth"], hidden_sizes=[meta["hidden"]] * meta["depth"],
            dropout_ps=dropout_ps or [0.0] * meta["depth"], activation_fn=ACTS[meta["act"]],
            use_learnable_skip=meta["skip"])
    if state is not None:
        m.load_state_dict(state)
    elif meta["skip"]:
        with torch.no_grad():
            for l, p in enumerate(m.skip_weights):
                p.fill_(1.0 - 0.15 * l + 0.05 * (l % 2))
    m.engine = engine
    return m.to(device)


# ---------------------------------------------------------------------------------------------
# The reference's training loop, restated for the tests (test infrastructure, like the oracle): what
# ``RxnGraphTrainer.train()`` does (reference training/trainer.py:185-217 with _train_epoch :124-155 and _val_epoch
# :157-183), over the test-only ``torch_geometric.loader.DataLoader`` stand-in (tests/_pyg_shim).  Pinned against the
# UNMODIFIED trainer by tests/golden/trainer_loop.npz (tests/golden/make_trainer_golden.py).
def run_reference_training_loop(model, optimizer, loss_fn, lr_scheduler, train_data, val_data, device, num_epochs,
                                batch_size, save_path, val_frequency=5):
    import sys
    shim = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_pyg_shim")
    if shim not in sys.path:
        sys.path.insert(0, shim)
    import torch_geometric as tg
    train_loader = tg.loader.DataLoader(dataset=train_data, batch_size=batch_size, shuffle=True)     # trainer.py:105-111
    val_loader = tg.loader.DataLoader(dataset=val_data, batch_size=batch_size, shuffle=False)        # trainer.py:112-118
    hist = {"train_losses": [], "val_losses": []}
    best = np.inf
    for epoch in range(num_epochs):                                                                  # trainer.py:195
        model.train()                                                                                # trainer.py:135
        total = 0.0
        for data in train_loader:                                                                    # trainer.py:138-147
            data = data.to(device)
            optimizer.zero_grad()
            pred = model(data)
            loss = loss_fn(pred, data.y)
            loss.backward()
            optimizer.step()
            total += loss_fn(pred, data.y).item()
        hist["train_losses"].append(float(np.sqrt(total / len(train_loader.dataset))))               # trainer.py:149
        if epoch % val_frequency == 0 or epoch == num_epochs - 1:                                    # trainer.py:200
            model.eval()                                                                             # trainer.py:167
            total = 0.0
            with torch.no_grad():
                for data in val_loader:                                                              # trainer.py:170-175
                    data = data.to(device)
                    total += loss_fn(model(data), data.y).item()
            val = float(np.sqrt(total / len(val_loader.dataset)))                                    # trainer.py:177
            hist["val_losses"].append(val)
            if val < best:                                                                           # trainer.py:205-208
                best = val
                torch.save(model, save_path)
        lr_scheduler.step()                                                                          # trainer.py:212
    return hist
