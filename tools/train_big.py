"""Fused training step at a large batch (cfg-4 sized batch of 8192 reactions): time and peak memory."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from cgr_mpnn_3d_b200.model import GNN
from cgr_mpnn_3d_b200.data import make_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
torch.manual_seed(0)
m = GNN(846, 14, depth=4, hidden_sizes=[400] * 4, dropout_ps=[0.0] * 4, activation_fn=F.relu,
        use_learnable_skip=True).to("cuda").train()
d = make_batch(B, seed=0, kind="t1x", fa=846).to("cuda")
for engine in ("auto", "tc_layerwise"):
    m.engine = engine
    for _ in range(2):
        m.zero_grad(set_to_none=True)
        F.mse_loss(m(d), d.y, reduction="sum").backward()
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        m.zero_grad(set_to_none=True)
        F.mse_loss(m(d), d.y, reduction="sum").backward()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"B={B} engine={engine} fused={m.__dict__['_last_fused_train']}: {ms:.2f} ms/step = {B / ms:.0f}k reactions/s, "
          f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB, grad norm "
          f"{float(sum((p.grad.double() ** 2).sum() for p in m.parameters()) ** 0.5):.6e}")
