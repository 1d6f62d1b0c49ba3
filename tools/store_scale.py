"""Resident-store screening at scale: N reactions packed in HBM, energies of all of them in one call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F
from cgr_mpnn_3d_b200.data import make_reactions
from cgr_mpnn_3d_b200.model import GNN
from cgr_mpnn_3d_b200.store import ReactionStore

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
t0 = time.perf_counter()
graphs = make_reactions(n, seed=1, kind="t1x", fa=846)
t1 = time.perf_counter()
store = ReactionStore.from_graphs(graphs, device="cuda")
del graphs
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"{n} reactions: synthetic generation {t1 - t0:.1f} s, packing + upload {t2 - t1:.1f} s, store {store.nbytes() / 2**30:.2f} GiB")
torch.manual_seed(0)
m = GNN(846, 14, depth=4, hidden_sizes=[400] * 4, dropout_ps=[0.0] * 4, activation_fn=F.relu,
        use_learnable_skip=True).to("cuda").eval()
order = np.random.default_rng(0).permutation(n)
for bs in (64, 1024, 8192):
    store.predict(m, batch_size=bs, order=order)          # warm-up: workspaces are sized for this job and cached
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = store.predict(m, batch_size=bs, order=order)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"batch {bs}: {n / dt / 1e6:.2f} M reactions/s ({dt * 1e3:.0f} ms for all {n}), finite={bool(torch.isfinite(out).all())}")
