"""Debug: clock64 stamps of the fused forward kernel (tc_fwd.cuh), per CTA and item (tile, layer).

stamp 0: accumulator ready seen by the epilogue; 1: gather of the last chunk done; 2: item tail (fences, cluster arrive /
readout reduction) done; 3: producer passed the dependency wait of this item (layers >= 1).
"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from cgr_mpnn_3D.models.GNN import GNN
from cgr_mpnn_3d_b200 import _lib
from cgr_mpnn_3d_b200.data import make_batch

ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=64); ap.add_argument("--policy", default="latency"); ap.add_argument("--precision", default="fp32"); a = ap.parse_args()
torch.manual_seed(0)
m = GNN(846, 14, depth=4, hidden_sizes=[400] * 4, dropout_ps=[0.0] * 4, activation_fn=F.relu, use_learnable_skip=True).cuda().eval()
m.tile_policy = a.policy
m.precision = a.precision
d = make_batch(a.batch, seed=0, fa=846).to("cuda")
lib = _lib.load()
PER = 2 * 16 * 4
with torch.no_grad():
    m(d); m(d)
    torch.cuda.synchronize()
    dbg = torch.zeros(40000 * PER, dtype=torch.int64, device="cuda")
    lib.cgr_tc_debug_buffer(dbg.data_ptr())
    m(d)
    torch.cuda.synchronize()
    lib.cgr_tc_debug_buffer(None)
t = dbg.view(-1, 32, 4).cpu()
t = t[t[:, 0, 0] != 0]
n_items = int((t[0, :, 0] != 0).sum())
print("ctas", t.shape[0], "items per cta", n_items)
t = t[:, :n_items].double()
start = t[:, 0, 0:1]
print("item: acc_ready(rel. to item 0)  gather(1-0)  tail(2-1)  gap to next acc_ready  dep_wait_passed(3, rel)")
for i in range(n_items):
    rel = (t[:, i, 0] - start[:, 0]).mean()
    gather = (t[:, i, 1] - t[:, i, 0]).mean()
    tail = (t[:, i, 2] - t[:, i, 1]).mean()
    gap = (t[:, i + 1, 0] - t[:, i, 2]).mean() if i + 1 < n_items else float("nan")
    dep = (t[:, i, 3] - start[:, 0]).mean() if i >= 1 and float(t[:, i, 3].min()) > 0 else float("nan")
    print(f"{i:3d}  {rel:10.0f}  {gather:8.0f}  {tail:8.0f}  {gap:8.0f}  {dep:10.0f}")
print("total (last tail - first acc_ready): mean %.0f cycles" % (t[:, n_items - 1, 2] - t[:, 0, 0]).mean())
