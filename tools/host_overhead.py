"""cProfile of the host side of eager training steps (where the CPU time between launches goes)."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from cgr_mpnn_3d_b200.model import GNN
from cgr_mpnn_3d_b200.data import make_batch

torch.manual_seed(0)
m = GNN(846, 14, depth=4, hidden_sizes=[400] * 4, dropout_ps=[0.0] * 4, activation_fn=F.relu,
        use_learnable_skip=True).to("cuda").train()
d = make_batch(64, seed=0, kind="t1x", fa=846).to("cuda")
mode = sys.argv[1] if len(sys.argv) > 1 else "train"

def step():
    if mode == "train":
        m.zero_grad(set_to_none=True)
        F.mse_loss(m(d), d.y, reduction="sum").backward()
    else:
        with torch.no_grad():
            m(d)
if mode != "train":
    m.eval()
for _ in range(5):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    step()
torch.cuda.synchronize()
print(f"{mode}: {(time.perf_counter() - t0) / 50 * 1e6:.1f} us per eager step")
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats(sys.argv[2] if len(sys.argv) > 2 else "cumulative").print_stats(40)
