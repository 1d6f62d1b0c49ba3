"""Tiny driver for profiling: N group forwards (GNN.forward_group) of G batches of the headline model.

    python tools/run_group.py --batch 64 --group 20 --iters 3
"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from cgr_mpnn_3D.models.GNN import GNN
from cgr_mpnn_3d_b200.data import make_batch

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64); ap.add_argument("--group", type=int, default=20)
ap.add_argument("--iters", type=int, default=3); ap.add_argument("--precision", default="fp32")
a = ap.parse_args()
torch.manual_seed(0)
m = GNN(846, 14, depth=4, hidden_sizes=[400] * 4, dropout_ps=[0.0] * 4, activation_fn=F.relu, use_learnable_skip=True).cuda().eval()
m.tile_policy = "throughput"; m.precision = a.precision
bs = [make_batch(a.batch, seed=i, fa=846).to("cuda") for i in range(a.group)]
with torch.no_grad():
    for _ in range(a.iters):
        outs = m.forward_group(bs)
torch.cuda.synchronize()
print("ok", float(sum(o.sum() for o in outs)))
