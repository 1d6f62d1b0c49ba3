"""predict_stream throughput for a few (depth, workers, coalesce) settings on pinned cfg-2 batches."""
import itertools, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from cgr_mpnn_3d_b200.model import GNN
from cgr_mpnn_3d_b200.data import make_batch

B, steps = 64, 1500
torch.manual_seed(0)
m = GNN(846, 14, depth=4, hidden_sizes=[400] * 4, dropout_ps=[0.0] * 4, activation_fn=F.relu,
        use_learnable_skip=True).to("cuda").eval()
pool = [make_batch(B, seed=i, kind="t1x", fa=846).pin_memory() for i in range(48)]
h2d = sum(t.numel() * t.element_size() for t in (pool[0].x, pool[0].edge_attr, pool[0].edge_index, pool[0].batch, pool[0].ptr))
for depth, workers, coalesce in [(4, 2, 8), (8, 2, 8), (4, 4, 8), (8, 4, 8), (4, 2, 16), (8, 4, 16), (3, 1, 8), (6, 3, 4), (8, 4, 2)]:
    list(m.predict_stream((pool[i % 48] for i in range(64)), depth=depth, workers=workers, coalesce=coalesce))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = sum(1 for _ in m.predict_stream((pool[i % 48] for i in range(steps)), depth=depth, workers=workers, coalesce=coalesce))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"depth={depth} workers={workers} coalesce={coalesce}: {B * n / dt / 1e3:.0f}k reactions/s, H2D {h2d * n / dt / 1e9:.1f} GB/s")
