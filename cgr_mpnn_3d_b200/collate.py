"""Device-side batch collation and CSR index construction (north-star subsystem 1).

Host-side mirror of what the reference obtains from PyG's ``DataLoader`` /
``Batch.from_data_list`` (``cgr_mpnn_3D/training/trainer.py:105-118``, ``test.py:85-90``) and of the
index arithmetic inside ``DMPNNConv.forward`` (``cgr_mpnn_3D/models/GNN.py:131-141``):

* :func:`collate` — list of per-reaction graphs -> one :class:`~cgr_mpnn_3d_b200.data.Batch`
  (``x``, ``edge_attr``, ``y`` concatenated; ``edge_index`` offset; ``batch``; ``ptr``), the integer
  fields produced by ``cgr_collate_indices`` on the GPU.
* :class:`GraphPlan` — the derived int32 arrays ``src`` (b2a), ``dst``, ``in_ptr``/``in_idx`` (a2b CSR,
  ascending bond id per atom) and ``atom_ptr``, produced by ``cgr_csr_build``; ``b2revb`` is ``e ^ 1``.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .data import Batch, Graph


def _stream() -> int:
    return _lib.current_stream_handle()


class GraphPlan:
    """Index arrays the kernels consume, built once per batch and cached on the batch object."""

    __slots__ = ("n_atoms", "n_bonds", "n_rxn", "src", "dst", "in_ptr", "in_idx", "atom_ptr", "status",
                 "tile_info", "n_tiles", "tc_ok", "tc_status")

    def __init__(self):
        self.tile_info = None    # tile plan of the tcgen05 engine (built lazily by ensure_tiles)
        self.n_tiles = 0
        self.tc_ok = None
        self.tc_status = None

    def ensure_tiles(self) -> bool:
        """Pack whole reactions into 128-bond row tiles (tcgen05 engine).  One host sync, cached."""
        if self.tc_ok is not None:
            return self.tc_ok
        lib = _lib.load()
        dev = self.src.device
        self.tile_info = torch.zeros((max(1, self.n_rxn), 8), dtype=torch.int32, device=dev)
        st = torch.zeros(2, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.cgr_tc_plan_build(self.in_ptr.data_ptr(), self.atom_ptr.data_ptr(), self.n_rxn,
                                             self.tile_info.data_ptr(), st.data_ptr(), _stream()), "cgr_tc_plan_build")
            _lib.check(lib.cgr_tc_plan_check(self.tile_info.data_ptr(), self.n_rxn, self.src.data_ptr(),
                                             self.dst.data_ptr(), st.data_ptr(), _stream()), "cgr_tc_plan_check")
        n_tiles, ok = (int(v) for v in st.tolist())
        self.n_tiles = n_tiles
        self.tc_ok = bool(ok) and n_tiles > 0
        # [0] sticky fp16-range flag of the tcgen05 engine, [1..] self-resetting readout arrival counters
        self.tc_status = torch.zeros(1 + max(1, n_tiles), dtype=torch.int32, device=dev)
        return self.tc_ok

    def check(self) -> None:
        """Synchronising validity check of the reference's silent preconditions (GNN.py:106,136-138)."""
        s = int(self.status.item())
        if s & 2:
            raise RuntimeError("edge_index holds an atom id outside [0, num_nodes)")
        if s & 1:
            raise RuntimeError("directed bonds are not adjacent (e, e^1) reverse pairs "
                               "(reference GNN.py:136-138 assumes graph_features.py:193-195 ordering)")
        if s & 4:
            raise RuntimeError("an atom has no incoming bond (reference GNN.py:106 raises on this input)")


def build_plan(edge_index: torch.Tensor, num_nodes: int, batch: Optional[torch.Tensor],
               ptr: Optional[torch.Tensor] = None, num_graphs: Optional[int] = None) -> GraphPlan:
    """CSR a2b / b2a arrays for a batched ``edge_index`` [2,E] int64 on the GPU."""
    if not edge_index.is_cuda:
        raise RuntimeError("build_plan needs CUDA tensors: the CGR hot path has no CPU implementation")
    lib = _lib.load()
    dev = edge_index.device
    ei = edge_index.contiguous()
    if ei.dtype != torch.int64:
        ei = ei.to(torch.int64)
    e = int(ei.shape[1])
    n = int(num_nodes)
    p = GraphPlan()
    p.n_atoms, p.n_bonds = n, e
    i32 = dict(dtype=torch.int32, device=dev)
    # the kernels launch on `dev` whatever the caller's current device is (a model / batch on cuda:1 in a process whose
    # current device is cuda:0)
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        p.src = torch.empty(e, **i32)
        p.dst = torch.empty(e, **i32)
        p.in_ptr = torch.empty(n + 1, **i32)
        p.in_idx = torch.empty(e, **i32)
        p.status = torch.empty(1, **i32)
        ws_bytes = lib.cgr_csr_workspace(n, e)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _lib.check(lib.cgr_csr_build(ei.data_ptr(), e, n, p.src.data_ptr(), p.dst.data_ptr(), p.in_ptr.data_ptr(),
                                     p.in_idx.data_ptr(), p.status.data_ptr(), ws.data_ptr(), ws_bytes, st),
                   "cgr_csr_build")
        if batch is None:
            p.n_rxn = 1
            p.atom_ptr = torch.tensor([0, n], **i32)
        else:
            if ptr is not None:
                p.n_rxn = int(ptr.numel()) - 1
                p.atom_ptr = ptr.to(device=dev, dtype=torch.int32)
            else:
                # reference semantics: global_add_pool sizes its output as batch.max()+1 (one host sync,
                # like PyG's own); pass `ptr` or `num_graphs` to avoid it
                p.n_rxn = int(num_graphs) if num_graphs is not None else int(batch.max()) + 1
                b = batch.contiguous()
                if b.dtype != torch.int64:
                    b = b.to(torch.int64)
                p.atom_ptr = torch.empty(p.n_rxn + 1, **i32)
                _lib.check(lib.cgr_atom_ptr_from_batch(b.data_ptr(), n, p.n_rxn, p.atom_ptr.data_ptr(), st),
                           "cgr_atom_ptr_from_batch")
    return p


def split_features_for(data, plan: GraphPlan):
    """FP16 (hi, lo) form of ``data.x`` for the tcgen05 engine, cached on the batch object (batch
    preparation, like the CSR arrays): ``cgr_tc_split_features``."""
    x = data.x
    key = (x.data_ptr(), x._version, tuple(x.shape))
    cached = getattr(data, "_cgr_xsplit", None)
    if cached is not None and cached[0] == key:
        return cached[1], cached[2]
    lib = _lib.load()
    xc = x.contiguous() if x.dtype == torch.float32 else x.float().contiguous()
    n, fa = int(xc.shape[0]), int(xc.shape[1])
    ld = int(lib.cgr_tc_features_ld(fa))
    x_hi = torch.empty((n, ld), dtype=torch.float16, device=xc.device)
    x_lo = torch.empty((n, ld), dtype=torch.float16, device=xc.device)
    with torch.cuda.device(xc.device):
        _lib.check(lib.cgr_tc_split_features(xc.data_ptr(), n, fa, x_hi.data_ptr(), x_lo.data_ptr(),
                                             plan.tc_status.data_ptr(), _stream()), "cgr_tc_split_features")
    try:
        data._cgr_xsplit = (key, x_hi, x_lo)
    except Exception:
        pass
    return x_hi, x_lo


def plan_for(data) -> GraphPlan:
    """Plan cached on the batch object (a batch is reused by forward, backward and every layer)."""
    plan = getattr(data, "_cgr_plan", None)
    ei = data.edge_index
    if plan is not None and plan.n_bonds == int(ei.shape[1]) and plan.src.device == ei.device \
            and getattr(data, "_cgr_plan_key", None) == (ei.data_ptr(), ei._version):
        return plan
    plan = build_plan(ei, int(data.x.shape[0]), getattr(data, "batch", None), getattr(data, "ptr", None))
    try:
        data._cgr_plan = plan
        data._cgr_plan_key = (ei.data_ptr(), ei._version)
    except Exception:
        pass
    return plan


def collate(graphs: Sequence[Graph], device="cuda", pinned: bool = True) -> Batch:
    """Collate per-reaction graphs into one device batch; integer fields are computed on the GPU.

    Float payloads (``x``, ``edge_attr``, ``y``) are concatenated on the host into (pinned) staging
    buffers and copied once; ``edge_index`` offsets, ``batch`` and ``ptr`` come from
    ``cgr_collate_indices`` and are bit-identical to PyG's collate.
    """
    lib = _lib.load()
    dev = torch.device(device)
    b = len(graphs)
    n_nodes = np.fromiter((g.x.shape[0] for g in graphs), dtype=np.int64, count=b)
    n_edges = np.fromiter((g.edge_index.shape[1] for g in graphs), dtype=np.int64, count=b)
    n, e = int(n_nodes.sum()), int(n_edges.sum())

    def stage(arr: np.ndarray) -> torch.Tensor:
        t = torch.from_numpy(np.ascontiguousarray(arr))
        if pinned:
            t = t.pin_memory()
        return t.to(dev, non_blocking=True)

    x = stage(np.concatenate([g.x for g in graphs], axis=0))
    ea = stage(np.concatenate([g.edge_attr for g in graphs], axis=0))
    y = stage(np.concatenate([g.y for g in graphs], axis=0))
    local = np.empty((2, e), dtype=np.int64)
    np.concatenate([g.edge_index[0] for g in graphs], out=local[0])
    np.concatenate([g.edge_index[1] for g in graphs], out=local[1])
    d_local, d_nn, d_ne = stage(local), stage(n_nodes), stage(n_edges)

    i64 = dict(dtype=torch.int64, device=dev)
    edge_index = torch.empty((2, e), **i64)
    batch = torch.empty(n, **i64)
    ptr = torch.empty(b + 1, **i64)
    edge_ptr = torch.empty(b + 1, **i64)
    ws_bytes = lib.cgr_collate_workspace(b)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.cgr_collate_indices(d_nn.data_ptr(), d_ne.data_ptr(), d_local.data_ptr(), b, e, n,
                                           edge_index.data_ptr(), batch.data_ptr(), ptr.data_ptr(),
                                           edge_ptr.data_ptr(), ws.data_ptr(), ws_bytes, _stream()),
                   "cgr_collate_indices")
    out = Batch(x=x, edge_index=edge_index, edge_attr=ea, batch=batch, ptr=ptr, y=y)
    out.edge_ptr = edge_ptr
    return out
