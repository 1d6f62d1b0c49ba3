"""Debug: clock64 stamps of the fused forward kernel (tc_fwd.cuh), per CTA and item (tile, layer).

stamp 0: gather warps see the item's first staged chunk; 1: gather of the last chunk done; 2: item tail (publish /
readout reduction) done; 3: producer passed the dependency wait of this item (layers >= 1); 4: MMA warp owns the
accumulator slot; 7: first operand stage of the item landed; 5: staging warps see the accumulator complete;
6: last chunk staged (accumulator released).

Needs a library built with the stamps compiled in:  CGR_FWD_STAMPS=1 python -m cgr_mpnn_3d_b200.build --force
(the shipped kernel has none), and a plain rebuild afterwards.
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
PER = 2 * 16 * 8
with torch.no_grad():
    m(d); m(d)
    torch.cuda.synchronize()
    dbg = torch.zeros(40000 * PER, dtype=torch.int64, device="cuda")
    lib.cgr_tc_debug_buffer(dbg.data_ptr())
    m(d)
    torch.cuda.synchronize()
    lib.cgr_tc_debug_buffer(None)
t = dbg.view(-1, 32, 8).cpu()
t = t[t[:, 0, 0] != 0]
n_items = int((t[0, :31, 0] != 0).sum())
print("ctas", t.shape[0], "items per cta", n_items)
t = t[:, :n_items].double()
start = t[:, 0:1, 0:1]
r = t - start
r[t == 0] = float("nan")
def col(i, k):
    return float(torch.nanmean(r[:, i, k]))
print("item:  mma_slot(4)  stage0_landed(7)  acc_done(5)  staged(6) | gather_first(0)  gather_done(1)  tail_done(2) | dep_passed(3)")
for i in range(n_items):
    print(f"{i:3d}  {col(i,4):10.0f} {col(i,7):10.0f} {col(i,5):10.0f} {col(i,6):10.0f} | {col(i,0):10.0f} {col(i,1):10.0f} {col(i,2):10.0f} | {col(i,3):10.0f}")
print("edge-init phase (gather warps, before the first item): %.0f cycles; it ends %.0f cycles before item 0's first chunk" % (float((t[:, 31, 1] - t[:, 31, 0]).mean()) if False else float((dbg.view(-1, 32, 8).cpu()[:t.shape[0], 31, 1] - dbg.view(-1, 32, 8).cpu()[:t.shape[0], 31, 0]).double().mean()), float((t[:, 0, 0] - dbg.view(-1, 32, 8).cpu()[:t.shape[0], 31, 1].double()).mean())))
print("per item means: mma (5-7) %.0f   fill wait (7-4) %.0f   gather (1-0) %.0f   tail (2-1) %.0f   period %.0f" % (
    sum(col(i, 5) - col(i, 7) for i in range(n_items)) / n_items, sum(col(i, 7) - col(i, 4) for i in range(n_items)) / n_items,
    sum(col(i, 1) - col(i, 0) for i in range(n_items)) / n_items, sum(col(i, 2) - col(i, 1) for i in range(n_items)) / n_items,
    (col(n_items - 1, 2) - col(0, 0)) / max(n_items - 1, 1)))
e = dbg.view(-1, 32, 8).cpu()[:t.shape[0], 31].double()
print("edge-init detail (cycles after its start): staged %.0f | tile0 computed %.0f published %.0f | tile1 computed %.0f published %.0f | end %.0f" % tuple(
    float((e[:, k] - e[:, 0]).mean()) for k in (2, 3, 4, 5, 6, 1)))
