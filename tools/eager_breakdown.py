"""Where the host time of an eager training step goes: wall time of the two C calls vs everything around them."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from cgr_mpnn_3d_b200 import _lib
from cgr_mpnn_3d_b200.model import GNN
from cgr_mpnn_3d_b200.data import make_batch

lib = _lib.load()
acc = {"fwd_c": 0.0, "bwd_c": 0.0}
_f, _b = lib.cgr_gnn_forward, lib.cgr_gnn_backward

def fwd(*a):
    t = time.perf_counter(); r = _f(*a); acc["fwd_c"] += time.perf_counter() - t; return r
def bwd(*a):
    t = time.perf_counter(); r = _b(*a); acc["bwd_c"] += time.perf_counter() - t; return r
lib.cgr_gnn_forward, lib.cgr_gnn_backward = fwd, bwd

torch.manual_seed(0)
m = GNN(846, 14, depth=4, hidden_sizes=[400] * 4, dropout_ps=[0.0] * 4, activation_fn=F.relu,
        use_learnable_skip=True).to("cuda").train()
d = make_batch(64, seed=0, kind="t1x", fa=846).to("cuda")
def step(parts=None):
    t0 = time.perf_counter()
    m.zero_grad(set_to_none=True)
    t1 = time.perf_counter()
    out = m(d)
    t2 = time.perf_counter()
    loss = F.mse_loss(out, d.y, reduction="sum")
    t3 = time.perf_counter()
    loss.backward()
    t4 = time.perf_counter()
    if parts is not None:
        for k, v in zip(("zero_grad", "forward", "loss", "backward"), (t1 - t0, t2 - t1, t3 - t2, t4 - t3)):
            parts[k] = parts.get(k, 0.0) + v
for _ in range(10):
    step()
torch.cuda.synchronize()
acc["fwd_c"] = acc["bwd_c"] = 0.0
parts = {}
n = 200
t0 = time.perf_counter()
for _ in range(n):
    step(parts)
torch.cuda.synchronize()
tot = time.perf_counter() - t0
print(f"eager step {tot / n * 1e6:.0f} us:", {k: round(v / n * 1e6) for k, v in parts.items()},
      "of which C calls:", {k: round(v / n * 1e6) for k, v in acc.items()})
