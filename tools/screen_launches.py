"""The kernels either side of the forward at batch 8192 (cfg-4 shapes), for an ncu launch list with DRAM bytes:
(1) the screening loop `ReactionStore.predict` (store_gather_split -> csr_by_reaction -> atom projection -> fused forward),
(2) the chain a FRESH device batch pays before its first forward: `store.batch` (cgr_store_gather, fp32 copy), the general
    CSR builder (count / scan / fill / sort), tile plan build + check, feature split -- then one forward.
Prints CUDA-event times of both when run plainly; under
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv
the launch list gives each kernel's duration and DRAM traffic (tools/launch_bytes_table.py turns it into a table)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F
from cgr_mpnn_3d_b200.collate import build_plan, plan_for, split_features_for
from cgr_mpnn_3d_b200.data import Batch, make_reactions
from cgr_mpnn_3d_b200.model import GNN
from cgr_mpnn_3d_b200.store import ReactionStore

n, bs = 16384, 8192
store = ReactionStore.from_graphs(make_reactions(n, seed=1, kind="t1x", fa=846), device="cuda")
torch.manual_seed(0)
m = GNN(846, 14, depth=4, hidden_sizes=[400] * 4, dropout_ps=[0.0] * 4, activation_fn=F.relu,
        use_learnable_skip=True).to("cuda").eval()
order = np.random.default_rng(0).permutation(n)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
with torch.no_grad():
    store.predict(m, batch_size=bs, order=order, slots=2)                    # warm-up: workspaces
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = store.predict(m, batch_size=bs, order=order, slots=2)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"(1) store.predict, 2 batches of {bs}: {dt * 1e3:.2f} ms = {n / dt / 1e6:.2f} M reactions/s")
    for rep in range(2):                                                    # second pass is the measured one
        ev[0].record()
        b = store.batch(order[:bs], with_plan=False)
        ev[1].record()
        fresh = Batch(b.x, b.edge_index.clone(), b.edge_attr, b.batch, b.ptr, None)   # no cached plan: the general builders
        plan = plan_for(fresh)
        plan.ensure_tiles()                                                  # tile plan build + check (one host sync)
        split_features_for(fresh, plan)
        ev[2].record()
        e = m(fresh)
        ev[3].record()
        torch.cuda.synchronize()
    N, E = int(b.x.shape[0]), int(b.edge_index.shape[1])
    print(f"(2) fresh batch of {bs} (N={N}, E={E}): store.batch {ev[0].elapsed_time(ev[1]):.3f} ms, "
          f"CSR + tile plan + feature split {ev[1].elapsed_time(ev[2]):.3f} ms, forward {ev[2].elapsed_time(ev[3]):.3f} ms")
    print("algorithmic bytes: gather+split", N * 846 * 8 + E * 14 * 8 + E * 2 * 12, " store.batch (fp32 copy)",
          N * 846 * 8 + E * 14 * 8 + E * 2 * 12 + N * 8, " feature split", N * 846 * 8)
    print("finite:", bool(torch.isfinite(out).all() and torch.isfinite(e).all()),
          " same energies either way:", float((out[:bs] - e).abs().max()))
