"""One training step (cfg-3 shape) between cudaProfilerStart/Stop, for an ncu launch list:
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv python tools/train_launches.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from cgr_mpnn_3d_b200.model import GNN
from cgr_mpnn_3d_b200.data import make_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
m = GNN(846, 14, depth=4, hidden_sizes=[400] * 4, dropout_ps=[0.0] * 4, activation_fn=F.relu,
        use_learnable_skip=True).to("cuda").train()
m.engine = os.environ.get("CGR_ENGINE", "auto")
d = make_batch(B, seed=0, kind="t1x", fa=846).to("cuda")
for _ in range(3):
    m.zero_grad(set_to_none=True)
    ((m(d) - d.y) ** 2).sum().backward()
torch.cuda.synchronize()
torch.cuda.profiler.start()
m.zero_grad(set_to_none=True)
((m(d) - d.y) ** 2).sum().backward()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
