"""Stress of the fused forward's layer hand-over (cluster-scope release / acquire + proxy fences): many repetitions of
group forwards and of a large batch must return bit-identical energies every time (a stale read of a peer's slice would
show as a run-to-run difference), and match the exact-fp32 engine."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from cgr_mpnn_3D.models.GNN import GNN
from cgr_mpnn_3d_b200.data import make_batch

torch.manual_seed(0)
m = GNN(846, 14, depth=4, hidden_sizes=[400] * 4, dropout_ps=[0.0] * 4, activation_fn=F.relu, use_learnable_skip=True).cuda().eval()
m.tile_policy = "throughput"
bs = [make_batch(64, seed=i, fa=846).to("cuda") for i in range(20)]
big = make_batch(8192, seed=99, fa=846).to("cuda")
bad = 0
with torch.no_grad():
    ref = [o.clone() for o in m.forward_group(bs)]
    refb = m(big).clone()
    for it in range(300):
        outs = m.forward_group(bs)
        if any(not torch.equal(a, b) for a, b in zip(outs, ref)):
            bad += 1
    for it in range(30):
        if not torch.equal(m(big), refb):
            bad += 1
    m.engine = "simt"
    exact = torch.cat([m(b) for b in bs])
torch.cuda.synchronize()
err = float((torch.cat(ref) - exact).abs().max() / exact.abs().mean())
print("mismatching repetitions:", bad, " max error vs exact-fp32 engine: %.2e" % err)
assert bad == 0 and err < 1e-4
