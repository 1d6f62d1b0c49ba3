"""cfg-5 stress shape (BASELINE.json configs[4]): drug-like reactions, depth 6, hidden 1024, batch 1024.
Times the inference forward for the engines that apply and checks them against each other."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from cgr_mpnn_3d_b200.model import GNN
from cgr_mpnn_3d_b200.data import make_batch
from cgr_mpnn_3d_b200 import _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
H = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
D = int(sys.argv[3]) if len(sys.argv) > 3 else 6
data = make_batch(B, seed=0, kind="drug", fa=78).to("cuda")
print(f"B={B} atoms={data.x.shape[0]} bonds={data.edge_index.shape[1]} H={H} depth={D}")
torch.manual_seed(0)
outs = {}
for engine in ("simt", "tc"):
    torch.manual_seed(0)
    m = GNN(78, 14, depth=D, hidden_sizes=[H] * D, dropout_ps=[0.0] * D, activation_fn=F.relu,
            use_learnable_skip=False).to("cuda").eval()
    m.engine = engine
    with torch.no_grad():
        for _ in range(3):
            out = m(data)
        torch.cuda.synchronize()
        l0 = _lib.load().cgr_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            out = m(data)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
    outs[engine] = out
    print(f"{engine}: {ms:.3f} ms/forward = {B / ms * 1e3:.0f} reactions/s, launches/forward "
          f"{(_lib.load().cgr_launch_count() - l0) // 5}")
d = (outs["simt"] - outs["tc"]).abs().max().item() / outs["simt"].abs().max().item()
print(f"simt vs tc scale-normalised diff {d:.2e}")
