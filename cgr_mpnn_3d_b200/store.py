"""Device-resident reaction store: batch assembly on the GPU.

The reference builds every batch on the host, item by item: ``ChemDataset.__getitem__`` →
``molgraph2data`` (``data/ChemDataset.py:69-94``: CGR atom features, float32 concatenation with the MACE block
``arr_{i}`` of the ``.npz`` written by ``download_preprocess_datasets.py:142``, ``edge_index`` transposed to
``[2, e]``, label from the ``.csv``) and PyG's collate (``training/trainer.py:105-118``).  A B200 holds the
whole featurised data set in HBM (1 M T1x-sized reactions with Fa = 846: 58 GB of 180 GB), so here the data set
is packed ONCE into contiguous device arrays and a batch is assembled by one kernel (``cgr_store_gather``) from
a list of reaction ids: no per-item Python, no per-step host→device copy of features.  The batch also carries
the index arrays of the kernels (one-launch CSR, tile plan computed on the host from the known offsets), so a
forward on it starts without any synchronisation.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterator, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .collate import GraphPlan
from .data import Batch


def _stream() -> int:
    return _lib.current_stream_handle()


class ReactionStore:
    """Packed reactions in device memory + device-side batch assembly."""

    def __init__(self, x_all, ea_all, ei_all, node_ptr, edge_ptr, y_all, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("ReactionStore lives in GPU memory: pass a CUDA device (there is no CPU path)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        if x_all.dtype != np.float32 or ea_all.dtype != np.float32:
            raise TypeError("features must be float32 (ChemDataset.py:83-86 builds float tensors)")
        self.fa, self.fb = int(x_all.shape[1]), int(ea_all.shape[1])
        self.node_ptr_host = np.ascontiguousarray(node_ptr, dtype=np.int64)
        self.edge_ptr_host = np.ascontiguousarray(edge_ptr, dtype=np.int64)
        self.n_rxn = int(self.node_ptr_host.shape[0]) - 1
        self.e_all = int(ea_all.shape[0])
        dev = self.device
        self.x_all = torch.from_numpy(np.ascontiguousarray(x_all)).to(dev)
        self.ea_all = torch.from_numpy(np.ascontiguousarray(ea_all)).to(dev)
        self.ei_all = torch.from_numpy(np.ascontiguousarray(ei_all, dtype=np.int32)).to(dev)      # [2, E_all] local ids
        self.node_ptr = torch.from_numpy(self.node_ptr_host).to(dev)
        self.edge_ptr = torch.from_numpy(self.edge_ptr_host).to(dev)
        self.y_all = torch.from_numpy(np.ascontiguousarray(y_all, dtype=np.float32)).to(dev)

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_graphs(cls, graphs: Sequence, device="cuda", mace_npz: Optional[str] = None,
                    labels: Optional[np.ndarray] = None) -> "ReactionStore":
        """Pack per-reaction graphs (fields of the reference's per-item ``Data``: ``x [n, F]``, ``edge_index [2, e]``
        local ids, ``edge_attr [e, Fb]``, ``y [1]``).  ``mace_npz``: the reference's ``.npz`` of per-reaction 3D
        descriptors (``arr_{i}`` → ``[n_atoms, F3D]``), concatenated to ``x`` in float32 as ``ChemDataset.py:83-86``
        does.  ``labels`` override ``y`` (second column of the reference's ``.csv``)."""
        xs = [np.asarray(g.x, dtype=np.float32) for g in graphs]
        if mace_npz is not None:
            with np.load(mace_npz) as z:
                for i in range(len(xs)):
                    m = np.asarray(z[f"arr_{i}"], dtype=np.float32)      # float64 descriptors would promote x
                    if m.shape[0] != xs[i].shape[0]:
                        raise ValueError(f"arr_{i}: {m.shape[0]} atoms, graph {i} has {xs[i].shape[0]}")
                    xs[i] = np.concatenate([xs[i], m], axis=1)
        n = np.fromiter((x.shape[0] for x in xs), dtype=np.int64, count=len(xs))
        e = np.fromiter((g.edge_index.shape[1] for g in graphs), dtype=np.int64, count=len(xs))
        node_ptr = np.zeros(len(xs) + 1, dtype=np.int64)
        edge_ptr = np.zeros(len(xs) + 1, dtype=np.int64)
        np.cumsum(n, out=node_ptr[1:])
        np.cumsum(e, out=edge_ptr[1:])
        x_all = np.concatenate(xs, axis=0)
        ea_all = np.concatenate([np.asarray(g.edge_attr, dtype=np.float32) for g in graphs], axis=0)
        ei_all = np.concatenate([np.asarray(g.edge_index, dtype=np.int64) for g in graphs], axis=1)
        if ei_all.size and (ei_all.min() < 0 or ei_all.max() >= 2 ** 31):
            raise ValueError("edge_index out of range")
        y_all = (np.asarray(labels, dtype=np.float32) if labels is not None
                 else np.concatenate([np.asarray(g.y, dtype=np.float32).reshape(-1)[:1] for g in graphs]))
        if y_all.shape[0] != len(xs):
            raise ValueError("one label per reaction expected")
        return cls(x_all, ea_all, ei_all.astype(np.int32), node_ptr, edge_ptr, y_all, device)

    @classmethod
    def from_reference_files(cls, graphs2d: Sequence, csv_path: str, npz_path: Optional[str] = None,
                             device="cuda") -> "ReactionStore":
        """Labels from the reference's ``.csv`` (``smiles,ea``: second column as float32, ``ChemDataset.py:27-32``) and
        the optional MACE ``.npz``; ``graphs2d`` are the CGR graphs of the same rows (the SMILES → CGR featurisation
        needs RDKit and stays outside this package)."""
        import pandas as pd
        labels = pd.read_csv(csv_path).iloc[:, 1].values.astype(np.float32)
        if labels.shape[0] != len(graphs2d):
            raise ValueError(f"{csv_path} has {labels.shape[0]} rows, {len(graphs2d)} graphs given")
        return cls.from_graphs(graphs2d, device=device, mace_npz=npz_path, labels=labels)

    def __len__(self) -> int:
        return self.n_rxn

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.x_all, self.ea_all, self.ei_all, self.node_ptr,
                                                          self.edge_ptr, self.y_all))

    # ------------------------------------------------------------------ batch assembly
    def batch(self, indices, with_plan: bool = True) -> Batch:
        """Assemble the batch of reactions ``indices`` (host int array / list / CPU tensor, any order) on the device.
        Fields equal the host collate of the same reactions bit for bit; ``with_plan`` also attaches the kernels' index
        arrays (built without a host synchronisation), so ``model(batch)`` launches straight away."""
        lib = _lib.load()
        sel = np.ascontiguousarray(np.asarray(indices, dtype=np.int64).reshape(-1))
        b = int(sel.shape[0])
        if b == 0:
            raise ValueError("empty batch")
        if sel.min() < 0 or sel.max() >= self.n_rxn:
            raise IndexError("reaction index out of range")
        n_sel = self.node_ptr_host[sel + 1] - self.node_ptr_host[sel]
        e_sel = self.edge_ptr_host[sel + 1] - self.edge_ptr_host[sel]
        meta = np.empty(3 * b + 2, dtype=np.int64)              # [sel | out_node_ptr | out_edge_ptr] in one upload
        meta[:b] = sel
        meta[b] = 0
        np.cumsum(n_sel, out=meta[b + 1:2 * b + 1])
        meta[2 * b + 1] = 0
        np.cumsum(e_sel, out=meta[2 * b + 2:])
        n_out, e_out = int(meta[2 * b]), int(meta[3 * b + 1])
        dev = self.device
        with torch.cuda.device(dev):
            meta_d = torch.from_numpy(meta).to(dev, non_blocking=True)
            sel_d, optr, oeptr = meta_d[:b], meta_d[b:2 * b + 1], meta_d[2 * b + 1:]
            x = torch.empty((n_out, self.fa), dtype=torch.float32, device=dev)
            ea = torch.empty((e_out, self.fb), dtype=torch.float32, device=dev)
            ei = torch.empty((2, e_out), dtype=torch.int64, device=dev)
            bvec = torch.empty(n_out, dtype=torch.int64, device=dev)
            y = torch.empty(b, dtype=torch.float32, device=dev)
            _lib.check(lib.cgr_store_gather(self.x_all.data_ptr(), self.ea_all.data_ptr(), self.ei_all.data_ptr(),
                                            self.node_ptr.data_ptr(), self.edge_ptr.data_ptr(), self.y_all.data_ptr(),
                                            self.e_all, sel_d.data_ptr(), optr.data_ptr(), oeptr.data_ptr(), b, self.fa,
                                            self.fb, e_out, x.data_ptr(), ea.data_ptr(), ei.data_ptr(), bvec.data_ptr(),
                                            y.data_ptr(), _stream()), "cgr_store_gather")
            data = Batch(x, ei, ea, bvec, optr, y)
            if with_plan:
                self._attach_plan(data, meta[b:2 * b + 1], meta[2 * b + 1:], optr, oeptr, n_out, e_out, b)
        return data

    def _attach_plan(self, data, node_ptr_h, edge_ptr_h, optr, oeptr, n_out, e_out, b) -> None:
        """Kernel index arrays from the offsets the host already knows: one-launch CSR (``cgr_csr_build_by_reaction``)
        and the tile plan of the tcgen05 engine computed on the host (``cgr_tc_plan_host``): no synchronisation."""
        lib = _lib.load()
        dev = self.device
        p = GraphPlan()
        p.n_atoms, p.n_bonds, p.n_rxn = n_out, e_out, b
        i32 = dict(dtype=torch.int32, device=dev)
        p.src = torch.empty(e_out, **i32)
        p.dst = torch.empty(e_out, **i32)
        p.in_ptr = torch.empty(n_out + 1, **i32)
        p.in_idx = torch.empty(e_out, **i32)
        p.status = torch.zeros(1, **i32)
        p.atom_ptr = optr.to(torch.int32)
        eptr32 = oeptr.to(torch.int32)
        if int((edge_ptr_h[1:] - edge_ptr_h[:-1]).max()) <= 256 and int((node_ptr_h[1:] - node_ptr_h[:-1]).max()) <= 256:
            _lib.check(lib.cgr_csr_build_by_reaction(data.edge_index.data_ptr(), eptr32.data_ptr(), p.atom_ptr.data_ptr(),
                                                     b, e_out, n_out, p.src.data_ptr(), p.dst.data_ptr(),
                                                     p.in_ptr.data_ptr(), p.in_idx.data_ptr(), p.status.data_ptr(),
                                                     _stream()), "cgr_csr_build_by_reaction")
        else:                                  # very large reactions: the general CSR builder
            ws_bytes = lib.cgr_csr_workspace(n_out, e_out)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            _lib.check(lib.cgr_csr_build(data.edge_index.data_ptr(), e_out, n_out, p.src.data_ptr(), p.dst.data_ptr(),
                                         p.in_ptr.data_ptr(), p.in_idx.data_ptr(), p.status.data_ptr(), ws.data_ptr(),
                                         ws_bytes, _stream()), "cgr_csr_build")
        tiles = np.zeros((b, 8), dtype=np.int32)
        n_tiles = C.c_int64(0)
        rc = lib.cgr_tc_plan_host(node_ptr_h.ctypes.data, edge_ptr_h.ctypes.data, b, tiles.ctypes.data, C.byref(n_tiles))
        if rc == 0 and n_tiles.value > 0:
            p.n_tiles = int(n_tiles.value)
            p.tile_info = torch.from_numpy(tiles[: max(1, p.n_tiles)]).to(dev, non_blocking=True)
            p.tc_ok = True
            p.tc_status = torch.zeros(1 + p.n_tiles, **i32)
        elif rc == -3:
            p.tc_ok = False                    # a reaction exceeds a 128-row tile: layer-wise kernels
            p.tc_status = torch.zeros(2, **i32)
        else:
            _lib.check(rc, "cgr_tc_plan_host")
        data._cgr_plan = p
        data._cgr_plan_key = (data.edge_index.data_ptr(), data.edge_index._version)

    def predict(self, model, batch_size: int = 64, order=None, slots: int = 8) -> torch.Tensor:
        """Energies of the reactions ``order`` (default: all, in store order) as one DEVICE tensor, computed by
        ``cgr_store_infer``: the whole per-batch loop (assemble on the device, index arrays, tcgen05 forward) runs in C,
        pipelined over ``slots`` streams; nothing but a few hundred bytes of offsets crosses the host link per batch.
        Batches are consecutive ``batch_size``-runs of ``order``.  Needs the tcgen05 engine (ReLU/SiLU/GELU all fine,
        hidden % 4 == 0, reactions of at most 128 directed bonds); otherwise use ``loader`` + ``model``."""
        lib = _lib.load()
        if not model._host_supported():
            raise RuntimeError("ReactionStore.predict needs a model the tcgen05 engine supports; iterate loader() instead")
        order = np.arange(self.n_rxn, dtype=np.int64) if order is None else \
            np.ascontiguousarray(np.asarray(order, dtype=np.int64).reshape(-1))
        n_total = int(order.shape[0])
        if n_total == 0:
            return torch.empty(0, dtype=torch.float32, device=self.device)
        ctx, dev = model._host_ctx(self.fa, self.fb)
        if dev != self.device:
            raise RuntimeError(f"model parameters live on {dev}, the store on {self.device}")
        cs = _lib.CgrStore(x_all=self.x_all.data_ptr(), ea_all=self.ea_all.data_ptr(), ei_all=self.ei_all.data_ptr(),
                           node_ptr=self.node_ptr.data_ptr(), edge_ptr=self.edge_ptr.data_ptr(),
                           node_ptr_host=self.node_ptr_host.ctypes.data, edge_ptr_host=self.edge_ptr_host.ctypes.data,
                           n_rxn=self.n_rxn, e_all=self.e_all, fa=self.fa, fb=self.fb)
        slots = max(1, min(int(slots), (n_total + batch_size - 1) // batch_size))
        dev_b, host_b = C.c_size_t(), C.c_size_t()
        _lib.check(lib.cgr_store_infer_workspace(C.byref(ctx.params), C.byref(cs), order.ctypes.data, n_total, batch_size,
                                                 C.byref(dev_b), C.byref(host_b)), "cgr_store_infer_workspace")
        # workspaces are kept across calls and only ever grow: a call that needs no more than an earlier one (fewer
        # slots, smaller batches) allocates nothing -- cudaMalloc / pinning of a batch-8192 workspace costs tens of ms
        need_d, need_h = dev_b.value * slots + 2048, host_b.value * slots
        cache = self.__dict__.get("_predict_ws")
        if cache is None or cache[0].numel() < need_d or cache[1].numel() < need_h or len(cache[2]) < slots:
            old_streams = cache[2] if cache is not None else []
            cache = None                                       # free the smaller workspace before taking the larger one
            self.__dict__["_predict_ws"] = None
            with torch.cuda.device(dev):
                # 2 % headroom: the next shard's largest batch is rarely exactly this one's
                cache = (torch.empty(need_d + need_d // 50, dtype=torch.uint8, device=dev),
                         torch.empty(need_h + need_h // 50, dtype=torch.uint8).pin_memory(),
                         old_streams + [torch.cuda.Stream(device=dev) for _ in range(slots - len(old_streams))])
            self.__dict__["_predict_ws"] = cache
        dws, hws, streams = cache[0], cache[1], cache[2][:slots]
        out = torch.empty(n_total, dtype=torch.float32, device=dev)
        cur = torch.cuda.current_stream(dev)
        for st in streams:
            st.wait_stream(cur)                    # weights / store contents written on the caller's stream
        sarr = (C.c_void_p * slots)(*[st.cuda_stream for st in streams])
        with torch.cuda.device(dev):
            rc = lib.cgr_store_infer(C.byref(ctx.params), C.byref(cs), order.ctypes.data, n_total, batch_size,
                                     out.data_ptr(), dws.data_ptr(), dev_b.value, hws.data_ptr(), host_b.value, slots, sarr)
        _lib.check(rc, "cgr_store_infer")
        return out

    def loader(self, batch_size: int, shuffle: bool = False, seed: int = 0, drop_last: bool = False,
               with_plan: bool = True) -> Iterator[Batch]:
        """One pass over the store in batches assembled on the device (what ``tg.loader.DataLoader(dataset, batch_size,
        shuffle)`` yields at ``trainer.py:105-118``, without the per-item host work)."""
        order = np.arange(self.n_rxn, dtype=np.int64)
        if shuffle:
            np.random.default_rng(seed).shuffle(order)
        for lo in range(0, self.n_rxn, batch_size):
            idx = order[lo:lo + batch_size]
            if drop_last and idx.shape[0] < batch_size:
                break
            yield self.batch(idx, with_plan=with_plan)
