"""Tiny driver for profiling: N eager forward passes of the headline model at a given batch size.

    python tools/run_forward.py --batch 64 --iters 6 [--engine tc|simt] [--train]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from cgr_mpnn_3D.models.GNN import GNN
from cgr_mpnn_3d_b200.data import make_batch

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--iters", type=int, default=6)
ap.add_argument("--engine", default="auto")
ap.add_argument("--train", action="store_true")
ap.add_argument("--hidden", type=int, default=400)
ap.add_argument("--depth", type=int, default=4)
ap.add_argument("--policy", default="latency")
ap.add_argument("--precision", default="fp32")
a = ap.parse_args()
torch.manual_seed(0)
m = GNN(846, 14, depth=a.depth, hidden_sizes=[a.hidden] * a.depth, dropout_ps=[0.0] * a.depth, activation_fn=F.relu,
        use_learnable_skip=True).cuda()
m.engine = a.engine
m.tile_policy = a.policy
m.precision = a.precision
d = make_batch(a.batch, seed=0, fa=846).to("cuda")
if a.train:
    m.train()
    for _ in range(a.iters):
        m.zero_grad()
        ((m(d) - d.y) ** 2).sum().backward()
else:
    m.eval()
    with torch.no_grad():
        for _ in range(a.iters):
            out = m(d)
torch.cuda.synchronize()
print("ok", float(out.sum()) if not a.train else "train")
