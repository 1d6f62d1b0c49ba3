"""CPU, world_size 2 over gloo: sharding and the SUM gradient all-reduce of data-parallel training."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cgr_mpnn_3d_b200.parallel import shard_by_bonds, shard_range


def test_shard_range_partitions():
    for n in (0, 1, 7, 64, 1000003):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_shard_by_bonds_balances():
    nb = [10] * 50 + [200] * 5
    spans = shard_by_bonds(nb, 4)
    assert spans[0][0] == 0 and spans[-1][1] == len(nb)
    assert all(spans[i][1] == spans[i + 1][0] for i in range(3))
    loads = [sum(nb[a:b]) for a, b in spans]
    assert max(loads) <= 2 * (sum(nb) / 4) + 200


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch.nn.functional as F
    from cgr_mpnn_3d_b200.data import collate_host, make_reactions
    from cgr_mpnn_3d_b200.parallel import allreduce_gradients_, broadcast_parameters_, shard_reactions
    from oracle.gnn_oracle import OracleGNN, mse_sum_loss
    torch.set_num_threads(1)
    rx = make_reactions(12, seed=3, kind="t1x", fa=78)
    torch.manual_seed(100 + rank)                      # deliberately different initial replicas
    model = OracleGNN(78, 14, depth=2, hidden_sizes=[32] * 2, dropout_ps=[0.0] * 2, activation_fn=F.relu,
                      use_learnable_skip=True)
    broadcast_parameters_(model.parameters(), src=0)
    mine = collate_host(shard_reactions(rx, rank, world))
    mse_sum_loss(model(mine), mine.y).backward()
    local = [p.grad.clone() for p in model.parameters()]
    allreduce_gradients_(model.parameters())           # SUM, not mean: loss is MSE(reduction="sum")
    # the same reduction when the gradients are views of ONE flat buffer (what the CGR backward returns): in place, one
    # collective, and the same numbers
    from cgr_mpnn_3d_b200.parallel import _shared_flat_view
    packed = [p.grad.clone() for p in model.parameters()]
    ps = list(model.parameters())
    sizes = [(p.numel() + 3) // 4 * 4 for p in ps]
    flat = torch.full((sum(sizes),), float("nan"))          # padding between pieces is never read back
    flat.zero_()
    off = 0
    for p, g, n in zip(ps, local, sizes):
        flat[off:off + p.numel()] = g.reshape(-1)
        p.grad = flat[off:off + p.numel()].view(p.shape)
        off += n
    assert _shared_flat_view(ps) is not None
    out = allreduce_gradients_(ps)
    assert out.data_ptr() == flat.data_ptr()
    assert all(torch.equal(p.grad, g) for p, g in zip(ps, packed))
    if rank == 0:
        ref = OracleGNN(78, 14, depth=2, hidden_sizes=[32] * 2, dropout_ps=[0.0] * 2, activation_fn=F.relu,
                        use_learnable_skip=True)
        ref.load_state_dict(model.state_dict())
        full = collate_host(rx)
        mse_sum_loss(ref(full), full.y).backward()
        err = max(float((a.grad - b.grad).abs().max() / b.grad.abs().max().clamp_min(1e-30))
                  for a, b in zip(model.parameters(), ref.parameters()))
        q.put(err)
    dist.barrier()
    dist.destroy_process_group()


def test_sum_allreduce_matches_single_process_on_concatenated_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert q.get(timeout=5) < 1e-5


def test_allreduce_uses_shared_flat_gradient_buffer_in_place():
    """Gradients that are views of one flat buffer (what the CGR backward returns) are reduced in place."""
    import torch
    from cgr_mpnn_3d_b200.parallel import _shared_flat_view, allreduce_gradients_
    shapes = [(5, 7), (5,), (1,), (3, 3), (1,)]
    ps = [torch.nn.Parameter(torch.zeros(s)) for s in shapes]
    sizes = [(p.numel() + 3) // 4 * 4 for p in ps]
    flat = torch.arange(sum(sizes), dtype=torch.float32)
    off = 0
    for p, n in zip(ps, sizes):
        p.grad = flat[off:off + p.numel()].view(p.shape)
        off += n
    view = _shared_flat_view(ps)
    assert view is not None and view.data_ptr() == flat.data_ptr() and view.numel() == off - sizes[-1] + ps[-1].numel()
    out = allreduce_gradients_(ps)               # no process group: nothing to reduce, same buffer returned
    assert out.data_ptr() == flat.data_ptr()
    # separately allocated gradients take the packing path
    for p in ps:
        p.grad = torch.ones_like(p)
    assert _shared_flat_view(ps) is None
    out = allreduce_gradients_(ps)
    assert out.numel() == sum(p.numel() for p in ps) and bool((out == 1).all())
