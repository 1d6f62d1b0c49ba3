"""Multi-GPU helpers: one process per GPU, reactions sharded across ranks.

* Inference: reactions are independent, so ranks take contiguous shards and never communicate
  (:func:`shard_range`, :func:`shard_reactions`).
* Data-parallel training: the reference's loss is ``MSELoss(reduction="sum")`` (``train.py:120``), so G
  replicas on batches of 64 match one process on the concatenated 64*G batch iff gradients are
  **summed** (not averaged).  :func:`allreduce_gradients_` packs every ``.grad`` into one flat fp32
  buffer (1,485,205 elements = 5.94 MB for d4/h400) and issues ONE all-reduce (NCCL over NVLink on
  GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced shard ``[lo, hi)`` of ``n_items`` for ``rank`` (first ``n % world`` ranks get one extra)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_reactions(reactions: Sequence, rank: int, world: int) -> Sequence:
    lo, hi = shard_range(len(reactions), rank, world)
    return reactions[lo:hi]


def shard_by_bonds(n_bonds: Sequence[int], world: int) -> List[Tuple[int, int]]:
    """Contiguous shards balanced by total directed bonds rather than reaction count (drug-like shapes)."""
    total = float(sum(n_bonds))
    bounds, acc, r = [0], 0.0, 1
    for i, e in enumerate(n_bonds):
        acc += e
        while r < world and acc >= total * r / world:
            bounds.append(i + 1)
            r += 1
    while len(bounds) < world:
        bounds.append(len(n_bonds))
    bounds.append(len(n_bonds))
    return [(bounds[i], bounds[i + 1]) for i in range(world)]


def flat_gradients(params: Iterable[torch.nn.Parameter]) -> Tuple[torch.Tensor, List[torch.nn.Parameter]]:
    ps = [p for p in params if p.requires_grad]
    if not ps:
        raise ValueError("no trainable parameters")
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float() for p in ps])
    return flat, ps


def _shared_flat_view(ps: List[torch.nn.Parameter]):
    """The gradients of the CGR backward are views of one flat buffer (ops.gnn_backward_impl); when every ``.grad`` still
    is (nothing accumulated or replaced them), return a 1-D view spanning all of them, else None."""
    gs = [p.grad for p in ps]
    if any(g is None or g.dtype != torch.float32 or not g.is_contiguous() for g in gs):
        return None
    base = gs[0].untyped_storage().data_ptr()
    end = gs[0].storage_offset()
    for g in gs:
        if g.untyped_storage().data_ptr() != base or g.storage_offset() < end or g.storage_offset() - end > 3:
            return None
        end = g.storage_offset() + g.numel()
    first = gs[0]
    if end > first.untyped_storage().nbytes() // 4:
        return None
    return first.as_strided((end - first.storage_offset(),), (1,), first.storage_offset())


def allreduce_gradients_(params: Iterable[torch.nn.Parameter], group=None) -> torch.Tensor:
    """In-place SUM all-reduce of all gradients through one flat buffer; returns the reduced buffer.  Gradients that
    already live in one flat buffer (the CGR backward's do) are reduced in place: one collective, no copies."""
    ps = [p for p in params if p.requires_grad]
    view = _shared_flat_view(ps) if ps else None
    if view is not None:
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(view, op=dist.ReduceOp.SUM, group=group)
        return view
    flat, ps = flat_gradients(ps)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for p in ps:
        n = p.numel()
        g = flat[off:off + n].view_as(p).to(p.dtype)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n
    return flat


def broadcast_parameters_(params: Iterable[torch.nn.Parameter], src: int = 0, group=None) -> None:
    """Replicas start from rank ``src``'s weights (one flat broadcast)."""
    ps = list(params)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    flat = torch.cat([p.detach().reshape(-1).float() for p in ps])
    dist.broadcast(flat, src=src, group=group)
    off = 0
    with torch.no_grad():
        for p in ps:
            n = p.numel()
            p.copy_(flat[off:off + n].view_as(p))
            off += n
