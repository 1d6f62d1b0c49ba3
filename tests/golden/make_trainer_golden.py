"""Generate the training-loop golden fixture by running the UNMODIFIED reference trainer.

Build-container only (needs ``/root/reference``; the GPU box never runs this):

    python tests/golden/make_trainer_golden.py

Imports ``/root/reference/cgr_mpnn_3D/training/trainer.py`` and ``models/GNN.py`` verbatim (``tests/_pyg_shim`` stands
in for torch_geometric; the RDKit-dependent ``ChemDataset`` and the W&B logger, which the trainer only names in type
hints, are stubbed) and runs ``RxnGraphTrainer.train()`` (trainer.py:185-217) for 3 epochs on 256 synthetic reactions
with the optimizer / loss / scheduler of ``train.py:117-121``: Adam(lr, weight_decay, amsgrad=True), MSELoss(sum),
ExponentialLR(gamma).  Stored: the per-epoch train / validation RMSE the trainer returns, the validation predictions of
the best model it pickled, and the seeds -- ``tests/test_gpu_parity.py::test_training_loop_matches_reference_trainer``
replays the same loop on the B200 drop-in and compares.
"""
import importlib.util
import os
import sys
import tempfile
import types

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests", "_pyg_shim"))
sys.path.insert(0, ROOT)

from cgr_mpnn_3d_b200.data import make_reactions  # noqa: E402

REF_ROOT = "/root/reference"
CFG = dict(fa=110, fb=14, depth=3, hidden=64, n_train=256, n_val=64, batch=64, epochs=3, lr=1e-3, wd=1e-5, gamma=0.9,
           data_seed=77, weight_seed=5, loop_seed=123)


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference_trainer():
    # modules the trainer imports at its top but only uses as type hints (trainer.py:8,10)
    for name in ("cgr_mpnn_3D", "cgr_mpnn_3D.data", "cgr_mpnn_3D.data.ChemDataset", "wandb_logger"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["cgr_mpnn_3D.data.ChemDataset"].ChemDataset = object
    sys.modules["wandb_logger"].WandBLogger = object
    gnn = _load(os.path.join(REF_ROOT, "cgr_mpnn_3D", "models", "GNN.py"), "_reference_gnn")
    trainer = _load(os.path.join(REF_ROOT, "cgr_mpnn_3D", "training", "trainer.py"), "_reference_trainer")
    return gnn, trainer


def datasets(cfg):
    rx = make_reactions(cfg["n_train"] + cfg["n_val"], seed=cfg["data_seed"], kind="t1x", fa=cfg["fa"])
    return rx[: cfg["n_train"]], rx[cfg["n_train"]:]


def main():
    cfg = CFG
    gnn, trainer_mod = load_reference_trainer()
    torch.set_num_threads(1)
    train_data, val_data = datasets(cfg)
    torch.manual_seed(cfg["weight_seed"])
    model = gnn.GNN(cfg["fa"], cfg["fb"], depth=cfg["depth"], hidden_sizes=[cfg["hidden"]] * cfg["depth"],
                    dropout_ps=[0.0] * cfg["depth"], activation_fn=F.relu, use_learnable_skip=True)
    init_state = {k: v.detach().clone().numpy() for k, v in model.state_dict().items()}
    opt = torch.optim.Adam(model.parameters(), lr=cfg["lr"], weight_decay=cfg["wd"], amsgrad=True)     # train.py:117-119
    loss_fn = torch.nn.MSELoss(reduction="sum")                                                       # train.py:120
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=cfg["gamma"])                            # train.py:121
    with tempfile.TemporaryDirectory() as tmp:
        t = trainer_mod.RxnGraphTrainer("golden", model, opt, loss_fn, sched, train_data, val_data, torch.device("cpu"),
                                        cfg["epochs"], model_save_dir=tmp, batch_size=cfg["batch"], num_workers=1,
                                        val_frequency=5, logger=None)
        torch.manual_seed(cfg["loop_seed"])        # fixes the shuffled order of every epoch
        hist = t.train()
        best = torch.load(os.path.join(tmp, "golden.pth"), map_location="cpu", weights_only=False).eval()
        import torch_geometric as tg
        with torch.no_grad():
            val_pred = torch.cat([best(b) for b in tg.loader.DataLoader(val_data, batch_size=cfg["batch"])]).numpy()
    out = {"train_losses": np.array(hist["train_losses"], dtype=np.float64),
           "val_losses": np.array(hist["val_losses"], dtype=np.float64), "val_pred": val_pred,
           "cfg_keys": np.array(sorted(cfg)), "cfg_vals": np.array([float(cfg[k]) for k in sorted(cfg)])}
    out.update({"w/" + k: v for k, v in init_state.items()})
    np.savez_compressed(os.path.join(HERE, "trainer_loop.npz"), **out)
    print("train RMSE per epoch:", hist["train_losses"], "val:", hist["val_losses"])


if __name__ == "__main__":
    main()
