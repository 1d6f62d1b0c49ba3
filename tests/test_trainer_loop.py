"""The reference's training loop against the drop-in module.

``tests/golden/trainer_loop.npz`` holds what the UNMODIFIED reference produced: ``RxnGraphTrainer.train()``
(training/trainer.py:185-217) driving the reference ``GNN`` for 3 epochs on 256 synthetic reactions with
Adam(amsgrad) / MSELoss(sum) / ExponentialLR (train.py:117-121).  ``/root/reference`` does not exist on the GPU box, so
the loop is restated in ``tests/util.py`` (each line cited) and pinned here on the CPU against that fixture with the
oracle model; the GPU test then runs the same loop on the B200 drop-in -- same shuffled batches, same optimizer -- and
compares the loss curve, the pickled best model (torch.save of the whole module, trainer.py:208) and its predictions
after the ``torch.load`` the reference's test.py:93 performs.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from cgr_mpnn_3d_b200.data import make_reactions
from tests.util import GOLDEN, run_reference_training_loop


def _fixture():
    z = np.load(os.path.join(GOLDEN, "trainer_loop.npz"))
    cfg = {str(k): float(v) for k, v in zip(z["cfg_keys"], z["cfg_vals"])}
    for k in ("fa", "fb", "depth", "hidden", "n_train", "n_val", "batch", "epochs", "data_seed", "weight_seed", "loop_seed"):
        cfg[k] = int(cfg[k])
    rx = make_reactions(cfg["n_train"] + cfg["n_val"], seed=cfg["data_seed"], kind="t1x", fa=cfg["fa"])
    state = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w/")}
    return z, cfg, rx[: cfg["n_train"]], rx[cfg["n_train"]:], state


def _run(model, cfg, train_data, val_data, device, save_path, optimizer_cls=torch.optim.Adam):
    opt = optimizer_cls(model.parameters(), lr=cfg["lr"], weight_decay=cfg["wd"], amsgrad=True)       # train.py:117-119
    loss_fn = torch.nn.MSELoss(reduction="sum")                                                      # train.py:120
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=cfg["gamma"])                           # train.py:121
    torch.manual_seed(cfg["loop_seed"])
    return run_reference_training_loop(model, opt, loss_fn, sched, train_data, val_data, device, cfg["epochs"],
                                       cfg["batch"], save_path)


def test_restated_loop_reproduces_the_reference_trainer(tmp_path):
    """CPU: the restated loop + the oracle model give exactly what the unmodified trainer + reference model gave."""
    from oracle.gnn_oracle import OracleGNN
    z, cfg, train_data, val_data, state = _fixture()
    torch.set_num_threads(1)
    model = OracleGNN(cfg["fa"], cfg["fb"], depth=cfg["depth"], hidden_sizes=[cfg["hidden"]] * cfg["depth"],
                      dropout_ps=[0.0] * cfg["depth"], activation_fn=F.relu, use_learnable_skip=True)
    model.load_state_dict(state)
    hist = _run(model, cfg, train_data, val_data, torch.device("cpu"), str(tmp_path / "best.pth"))
    np.testing.assert_allclose(hist["train_losses"], z["train_losses"], rtol=1e-12, atol=0)
    np.testing.assert_allclose(hist["val_losses"], z["val_losses"], rtol=1e-12, atol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("optimizer", ["torch_adam", "fused_adam"])
def test_training_loop_matches_reference_trainer(tmp_path, optimizer):
    """B200: same loop, drop-in ``cgr_mpnn_3D.models.GNN.GNN`` on cuda: loss curve within 1e-4 per epoch."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import sys
    shim = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_pyg_shim")
    if shim not in sys.path:
        sys.path.insert(0, shim)
    import torch_geometric as tg
    from cgr_mpnn_3D.models.GNN import GNN
    from cgr_mpnn_3d_b200.optim import FusedAdam
    from oracle.gnn_oracle import scale_normalised_error
    z, cfg, train_data, val_data, state = _fixture()
    model = GNN(cfg["fa"], cfg["fb"], depth=cfg["depth"], hidden_sizes=[cfg["hidden"]] * cfg["depth"],
                dropout_ps=[0.0] * cfg["depth"], activation_fn=F.relu, use_learnable_skip=True)
    model.load_state_dict(state)
    model = model.to("cuda")                                                                         # train.py:114
    path = str(tmp_path / "best.pth")
    hist = _run(model, cfg, train_data, val_data, torch.device("cuda"), path,
                torch.optim.Adam if optimizer == "torch_adam" else FusedAdam)
    np.testing.assert_allclose(hist["train_losses"], z["train_losses"], rtol=1e-4, atol=0)
    np.testing.assert_allclose(hist["val_losses"], z["val_losses"], rtol=1e-4, atol=0)
    # the pickle the trainer wrote (torch.save(self.model), trainer.py:208) loads the way test.py:93 loads it and
    # predicts what the reference's best model predicted, on the device (test.py:96-113) and on the CPU (CLI :61-76)
    best = torch.load(path, map_location="cuda", weights_only=False).eval()
    assert type(best).__module__ == "cgr_mpnn_3D.models.GNN"
    with torch.no_grad():
        pred = torch.cat([best(b.to("cuda")) for b in tg.loader.DataLoader(val_data, batch_size=cfg["batch"])])
        assert scale_normalised_error(pred, torch.from_numpy(z["val_pred"])) < 1e-4
        best_cpu = torch.load(path, map_location="cpu", weights_only=False).eval()
        one = next(iter(tg.loader.DataLoader(val_data, batch_size=1)))                # B = 1, as test.py:85-90 builds it
        assert scale_normalised_error(best_cpu(one), torch.from_numpy(z["val_pred"][:1])) < 1e-4
