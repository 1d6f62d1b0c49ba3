"""Data-parallel check on real GPUs (torchrun, NCCL): gradients SUM-reduced over ranks == gradients of one process on the
concatenated batch (the reference's loss is MSELoss(reduction="sum"), train.py:120).  Prints one line on rank 0."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch.nn.functional as F
from cgr_mpnn_3d_b200.data import collate_host, make_reactions
from cgr_mpnn_3d_b200.model import GNN
from cgr_mpnn_3d_b200.parallel import _shared_flat_view, allreduce_gradients_, broadcast_parameters_, shard_reactions

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl")
rx = make_reactions(64 * world, seed=5, kind="t1x", fa=846)
torch.manual_seed(1 + rank)
m = GNN(846, 14, depth=4, hidden_sizes=[400] * 4, dropout_ps=[0.0] * 4, activation_fn=F.relu,
        use_learnable_skip=True).cuda().train()
broadcast_parameters_(m.parameters(), src=0)
mine = collate_host(shard_reactions(rx, rank, world)).to("cuda")
F.mse_loss(m(mine), mine.y, reduction="sum").backward()
fused = m.__dict__["_last_fused_train"]
in_place = _shared_flat_view([p for p in m.parameters()]) is not None
allreduce_gradients_(m.parameters())
torch.cuda.synchronize()
if rank == 0:
    ref = GNN(846, 14, depth=4, hidden_sizes=[400] * 4, dropout_ps=[0.0] * 4, activation_fn=F.relu,
              use_learnable_skip=True).cuda().train()
    ref.load_state_dict(m.state_dict())
    full = collate_host(rx).to("cuda")
    F.mse_loss(ref(full), full.y, reduction="sum").backward()
    worst = 0.0
    for a, b in zip(m.parameters(), ref.parameters()):
        err = ((a.grad.double() - b.grad.double()).abs() / b.grad.double().abs().max().clamp_min(1e-30)).flatten()
        q = float(torch.quantile(err[: 2 ** 24].float(), 0.995)) if err.numel() >= 1000 else float(err.max())
        worst = max(worst, q)
    import json
    ok = bool(fused and in_place and worst < 1e-4)
    print(f"world={world} fused_train={fused} in_place_allreduce={in_place} worst 99.5% grad error vs single process: {worst:.2e}",
          flush=True)
    print("CHECK_DP " + json.dumps({"world": world, "fused_train": bool(fused), "in_place_allreduce": bool(in_place),
                                    "worst_q995_grad_error": worst, "ok": ok}), flush=True)
dist.barrier()
dist.destroy_process_group()
