"""What switching from the reference looks like, end to end (synthetic data: no RDKit / data set in this image).

    reference (train.py / trainer.py / test.py)                     here
    ------------------------------------------------------------    -----------------------------------------------
    from cgr_mpnn_3D.models.GNN import GNN                           same import (drop-in module)
    ChemDataset + tg.loader.DataLoader(shuffle=True)                 ReactionStore(...).loader(batch, shuffle=True)
    torch.optim.Adam(..., amsgrad=True) + ExponentialLR              FusedAdam(..., amsgrad=True) + ExponentialLR
    loss = MSELoss(reduction="sum"); loss.backward(); opt.step()     unchanged
    torch.save(model, path) / torch.load(path)                       unchanged (load_reference_checkpoint for old files)
    for batch in test_loader: model(batch)                           store.predict(model, batch_size)
"""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F
from cgr_mpnn_3D.models.GNN import GNN                       # the reference's import path
from cgr_mpnn_3d_b200.checkpoint import load_reference_checkpoint
from cgr_mpnn_3d_b200.data import make_reactions
from cgr_mpnn_3d_b200.optim import FusedAdam
from cgr_mpnn_3d_b200.store import ReactionStore

torch.manual_seed(0)
fa = 78 + 96
graphs = make_reactions(2048, seed=0, kind="t1x", fa=fa)
for g in graphs:                                              # a learnable synthetic target
    g.y[:] = 0.05 * g.x[:, :8].sum() + 0.01 * g.num_edges
store = ReactionStore.from_graphs(graphs[:1792], device="cuda")
held_out = ReactionStore.from_graphs(graphs[1792:], device="cuda")

model = GNN(fa, 14, depth=3, hidden_sizes=[128] * 3, dropout_ps=[0.02] * 3, use_learnable_skip=True).to("cuda")
opt = FusedAdam(model.parameters(), lr=1e-3, weight_decay=1e-6, amsgrad=True)          # train.py:117-119
sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.9)                          # train.py:121
loss_fn = torch.nn.MSELoss(reduction="sum")                                             # train.py:120
for epoch in range(5):
    model.train()
    t0, tot = time.perf_counter(), 0.0
    for batch in store.loader(64, shuffle=True, seed=epoch):
        opt.zero_grad()
        loss = loss_fn(model(batch), batch.y)                                           # trainer.py:141-142
        loss.backward()
        opt.step()
        tot += float(loss.detach())
    sched.step()
    torch.cuda.synchronize()
    model.eval()
    with torch.no_grad():
        pred = held_out.predict(model, batch_size=64)
        rmse = float(((pred - held_out.y_all) ** 2).mean().sqrt())
    print(f"epoch {epoch}: train RMSE {np.sqrt(tot / len(store)):.4f}  held-out RMSE {rmse:.4f}  "
          f"({len(store) / (time.perf_counter() - t0):.0f} reactions/s incl. evaluation)")
with tempfile.TemporaryDirectory() as tmp:
    path = os.path.join(tmp, "model.pth")
    torch.save(model, path)                                                             # trainer.py:208
    again = load_reference_checkpoint(path, map_location="cuda").eval()                 # test.py:93
    with torch.no_grad():
        assert torch.allclose(held_out.predict(again, batch_size=64), pred, atol=1e-5)
print("checkpoint round trip ok")
