"""Tiny driver for profiling the fused forward: N eager inference forwards (tile policy selectable)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from cgr_mpnn_3D.models.GNN import GNN
from cgr_mpnn_3d_b200.data import make_batch
ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=64); ap.add_argument("--iters", type=int, default=6)
ap.add_argument("--policy", default="throughput"); a = ap.parse_args()
torch.manual_seed(0)
m = GNN(846, 14, depth=4, hidden_sizes=[400] * 4, dropout_ps=[0.0] * 4, activation_fn=F.relu, use_learnable_skip=True).cuda().eval()
m.tile_policy = a.policy
d = make_batch(a.batch, seed=0, fa=846).to("cuda")
with torch.no_grad():
    for _ in range(a.iters):
        out = m(d)
torch.cuda.synchronize()
print("ok", float(out.sum()))
