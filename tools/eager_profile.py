"""cProfile of the eager training step's host side (cfg-3 shape): where the Python time around the two C calls goes."""
import cProfile, os, pstats, sys, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from cgr_mpnn_3d_b200.model import GNN
from cgr_mpnn_3d_b200.data import make_batch

torch.manual_seed(0)
m = GNN(846, 14, depth=4, hidden_sizes=[400] * 4, dropout_ps=[0.0] * 4, activation_fn=F.relu,
        use_learnable_skip=True).to("cuda").train()
d = make_batch(64, seed=0, kind="t1x", fa=846).to("cuda")


def step():
    m.zero_grad(set_to_none=True)
    loss = F.mse_loss(m(d), d.y, reduction="sum")
    loss.backward()


for _ in range(20):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(300):
    step()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])
