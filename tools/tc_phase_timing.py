"""Debug: per-phase clock64 breakdown of the bond-layer kernel (last layer of one forward)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from cgr_mpnn_3D.models.GNN import GNN
from cgr_mpnn_3d_b200 import _lib
from cgr_mpnn_3d_b200.data import make_batch

ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=64); ap.add_argument("--policy", default="latency"); a = ap.parse_args()
torch.manual_seed(0)
m = GNN(846, 14, depth=4, hidden_sizes=[400] * 4, dropout_ps=[0.0] * 4, activation_fn=F.relu, use_learnable_skip=True).cuda().eval()
m.tile_policy = a.policy
d = make_batch(a.batch, seed=0, fa=846).to("cuda")
lib = _lib.load()
with torch.no_grad():
    m(d); m(d)
    torch.cuda.synchronize()
    dbg = torch.zeros(200000 * 8, dtype=torch.int64, device="cuda")
    lib.cgr_tc_debug_buffer(dbg.data_ptr())
    m(d)
    torch.cuda.synchronize()
    lib.cgr_tc_debug_buffer(None)
t = dbg.view(-1, 8).cpu()
t = t[t[:, 0] != 0].double()
names = ["setup(1-0)", "first_data(6-1)", "stream(7-6)", "mma_tail(2-7)", "tmem2smem+R(3-2)", "rows(5-3)", "total(5-0)"]
cols = [t[:, 1] - t[:, 0], t[:, 6] - t[:, 1], t[:, 7] - t[:, 6], t[:, 2] - t[:, 7], t[:, 3] - t[:, 2], t[:, 5] - t[:, 3], t[:, 5] - t[:, 0]]
print("ctas", t.shape[0])
for n, c in zip(names, cols):
    print(f"{n:18s} mean {c.mean():9.0f}  min {c.min():9.0f}  max {c.max():9.0f} cycles")
