"""Drop-in ``GNN`` / ``DMPNNConv`` with the reference's module API and state_dict layout.

Mirrors ``cgr_mpnn_3D/models/GNN.py`` of the reference: constructor arguments (``:14-25``), the
parameter tree and its registration order (``:53-74``: ``edge_init``, ``convs.{l}.lin``,
``edge_to_node``, ``ffn``, ``skip_weights.{l}``), ``forward(data)`` on a batched CGR graph whose
``data.x`` already carries the MACE 3D fingerprint columns (``:76-110``).  The arithmetic runs in
``libcgr_b200.so`` through the ``cgr_b200::gnn_forward`` custom op; nothing is computed in Python
and there is no CPU path.  Host (CPU) inputs are staged to the GPU inside ``forward`` — that is the
end-to-end entry the reference's CPU-only CLI (``cli_tool/activation_energy_predictor.py:61-76``)
exercises — and the result is returned on the caller's device.
"""
from __future__ import annotations

import itertools
from typing import Callable, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .collate import plan_for, split_features_for

_ACT_BY_FN = {F.relu: 0, F.silu: 1, F.gelu: 2, torch.relu: 0}
_ACT_BY_NAME = {"relu": 0, "silu": 1, "gelu": 2}
_seed_counter = itertools.count(1)


def global_add_pool(x: torch.Tensor, batch: Optional[torch.Tensor], size: Optional[int] = None) -> torch.Tensor:
    """Stand-in for ``torch_geometric.nn.global_add_pool`` (the reference's default ``pooling_fn``,
    GNN.py:23).  Only stored as an attribute for API compatibility; pooling itself is fused in the
    readout kernel."""
    if batch is None:
        return x.sum(dim=-2, keepdim=True)
    n = int(batch.max()) + 1 if size is None else size
    return x.new_zeros((n, x.shape[-1])).index_add_(0, batch, x)


def _act_id(fn) -> int:
    if isinstance(fn, str):
        return _ACT_BY_NAME[fn.lower()]
    if fn in _ACT_BY_FN:
        return _ACT_BY_FN[fn]
    name = getattr(fn, "__name__", "")
    if name in _ACT_BY_NAME:
        return _ACT_BY_NAME[name]
    raise ValueError(f"activation_fn {fn!r} is not supported by the CUDA path (F.relu, F.silu, F.gelu)")


class DMPNNConv(nn.Module):
    """Directed message-passing layer, reference GNN.py:113-145.

    Holds ``lin = Linear(H, H)`` exactly like the reference so the state_dict keys are
    ``convs.{l}.lin.{weight,bias}``.  Called on its own it performs one aggregation +
    projection through the C ABI: ``forward(edge_index, edge_attr) -> (a_message, lin(a[row] - rev))``.
    """

    def __init__(self, hidden_size: int, aggr: str = "add"):
        super().__init__()
        if aggr != "add":
            raise ValueError("only aggr='add' is supported (the reference's default and only used value)")
        self.aggr = aggr
        self.lin = nn.Linear(hidden_size, hidden_size)

    def forward(self, edge_index: torch.Tensor, edge_attr: torch.Tensor):
        from .stage_ops import conv_forward
        return conv_forward(self, edge_index, edge_attr)


class GNN(nn.Module):
    """Reference-compatible CGR D-MPNN (+3D fingerprint) model running on hand-written sm_100a kernels."""

    #: synchronising validation of the reference's silent input preconditions (debug aid)
    validate_inputs: bool = False

    def __init__(
        self,
        num_node_features: int,
        num_edge_features: int,
        depth: int = 3,
        hidden_sizes: list = None,
        dropout_ps: list = None,
        activation_fn: Callable = F.relu,
        aggr: str = "add",
        pooling_fn: Callable = global_add_pool,
        use_learnable_skip: bool = False,
    ):
        super().__init__()
        self.depth = depth
        self.hidden_sizes = hidden_sizes or [300] * depth
        self.dropout_ps = dropout_ps or [0.02] * depth
        self.activation_fn = activation_fn
        self.pooling_fn = pooling_fn
        self.use_learnable_skip = use_learnable_skip
        self.engine = "auto"
        #: "fp32": FP16x3-split tensor-core GEMMs, 1e-4 parity with the reference (2e-6 of fp64).  "fast": single-pass
        #: fp16 operands in the inference kernels (one MMA per k-step instead of three), ~1e-3; training is unaffected.
        self.precision = "fp32"
        #: "latency": a lone forward gets the one-CTA-per-SM kernels; "throughput": the caller pipelines several
        #: forwards over CUDA streams (bench, predict_stream) and wants the two-CTAs-per-SM configuration
        self.tile_policy = "latency"

        self.edge_init = nn.Linear(num_node_features + num_edge_features, self.hidden_sizes[0])
        self.convs = nn.ModuleList()
        for i in range(self.depth):
            self.convs.append(DMPNNConv(self.hidden_sizes[i], aggr=aggr))   # IndexError like GNN.py:59-60
        self.edge_to_node = nn.Linear(num_node_features + self.hidden_sizes[-1], self.hidden_sizes[-1])
        self.ffn = nn.Linear(self.hidden_sizes[-1], 1)
        if self.use_learnable_skip:
            self.skip_weights = nn.ParameterList(
                [nn.Parameter(torch.tensor(1.0)) for _ in range(self.depth)])
        self.num_node_features = num_node_features
        self.num_edge_features = num_edge_features

    # ------------------------------------------------------------------------------------------
    def _param_list(self) -> List[torch.Tensor]:
        # module attribute lookups cost ~1 us each: keep the list, re-derive it if a parameter object was replaced
        cached = self.__dict__.get("_plist")
        if cached is not None:
            i_fb = 2 + 2 * self.depth + 3                              # ffn.bias
            if cached[0] is self._modules["edge_init"]._parameters["weight"] and \
                    cached[i_fb] is self._modules["ffn"]._parameters["bias"]:
                return list(cached)
        ps = self._param_list_uncached()
        self.__dict__["_plist"] = tuple(ps)
        return ps

    def _param_list_uncached(self) -> List[torch.Tensor]:
        ps = [self.edge_init.weight, self.edge_init.bias]
        for c in self.convs:
            ps += [c.lin.weight, c.lin.bias]
        ps += [self.edge_to_node.weight, self.edge_to_node.bias, self.ffn.weight, self.ffn.bias]
        if self.use_learnable_skip:
            ps += list(self.skip_weights)
        return ps

    def _engine_id(self, plan, needs_saved: bool) -> int:
        """simt = exact-fp32 layer-wise kernels.  tc = tcgen05 FP16x3 tensor-core kernels: fused tile kernels for
        an inference forward (needs a hidden size divisible by 4 and reactions of at most 128 directed bonds),
        layer-wise path with tensor-core GEMMs when activations must be saved for a backward.
        auto picks tc whenever it applies -- unless the current weights already drove an activation out of the fp16
        range of the FP16x3 split (see _settle_overflow): then the exact-fp32 engine is used until the weights change."""
        e = getattr(self, "engine", "auto")
        if e in ("simt", 0):
            return _lib.ENGINE_SIMT
        can_tc = self.depth <= 13
        if e in ("tc", "tc_layerwise", 1) and not can_tc:
            raise RuntimeError("engine='tc' supports depth <= 13")
        if can_tc and self._tc_demoted():
            if e in ("tc", "tc_layerwise", 1):
                raise RuntimeError("tcgen05 engine: an activation exceeded the fp16 range of the FP16x3 split with these "
                                   "weights (the energies of that forward were NaN); use engine='auto' or 'simt'")
            return _lib.ENGINE_SIMT
        return _lib.ENGINE_TC if can_tc else _lib.ENGINE_SIMT

    # ------------------------------------------------------------------ fp16-range guard of the tcgen05 engine ----
    # The FP16x3 operand split represents |v| < 65504.  The kernels flag anything larger in tc_status[0] and write NaN
    # energies (nothing wrong can pass silently).  The host reads the flag back WITHOUT stalling the stream:
    #   * CPU callers and eager training synchronise anyway (result copy / loss.item()), so the flag is read right
    #     there and the step is re-run on the exact-fp32 engine: correct energies, loss and gradients;
    #   * a device-tensor inference forward only queues a 4-byte copy; when a later call finds it set, this weight
    #     version is demoted to the exact-fp32 engine (engine='auto') -- the flagged forward itself returned NaN;
    #   * under CUDA-graph capture nothing can be read back: NaN poisoning is the signal.
    def _weights_key(self):
        return tuple((p.data_ptr(), p._version) for p in self._param_list())

    def _tc_demoted(self) -> bool:
        self._poll_overflow()
        key = self.__dict__.get("_tc_demoted_key")
        if key is None:
            return False
        if key != self._weights_key():
            self.__dict__["_tc_demoted_key"] = None      # weights changed: give the tensor-core engine another go
            return False
        return True

    def _queue_overflow_check(self, tc_status: torch.Tensor):
        """Queue a 4-byte copy of the flag word into a pinned ring slot; returns (event, ring, slot)."""
        ring = self.__dict__.get("_ovf_ring")
        if ring is None:
            ring = torch.zeros(64, dtype=torch.int32).pin_memory()
            self.__dict__["_ovf_ring"] = ring
            self.__dict__["_ovf_next"] = 0
        pend = self.__dict__.setdefault("_ovf_pending", [])
        if len(pend) >= 60:                      # never hand out a slot an unfinished copy still targets
            self._poll_overflow(wait=True)
        slot = self.__dict__["_ovf_next"]
        self.__dict__["_ovf_next"] = (slot + 1) % 64
        ring[slot:slot + 1].copy_(tc_status[:1], non_blocking=True)
        evs = self.__dict__.setdefault("_ovf_events", {})        # one reusable event per ring slot and device
        key = (slot, tc_status.device.index)
        ev = evs.get(key)
        if ev is None:
            ev = evs[key] = torch.cuda.Event()
        ev.record()
        return ev, ring, slot

    def _poll_overflow(self, wait: bool = False) -> None:
        pend = self.__dict__.get("_ovf_pending")
        if not pend or torch.cuda.is_current_stream_capturing():      # event queries are illegal during graph capture
            return
        keep = []
        for ev, ring, slot in pend:
            if wait:
                ev.synchronize()
            if not ev.query():
                keep.append((ev, ring, slot))
            elif int(ring[slot]) & 3:
                self._demote(self._weights_key(), int(ring[slot]))
        self.__dict__["_ovf_pending"] = keep

    def _demote(self, key, bits: int) -> None:
        import warnings
        self.__dict__["_tc_demoted_key"] = key
        warnings.warn("cgr_mpnn_3d_b200: %s left the fp16 range of the tcgen05 FP16x3 split (|v| >= 65504); that forward "
                      "returned NaN energies unless it was re-run, and this weight version now runs on the exact-fp32 "
                      "engine" % ("an input feature" if bits & 2 else "an activation"), RuntimeWarning, stacklevel=3)

    def _fast(self) -> bool:
        prec = getattr(self, "precision", "fp32")
        if prec not in ("fp32", "fast"):
            raise ValueError("model.precision must be 'fp32' (parity mode) or 'fast' (single-pass fp16 inference)")
        return prec == "fast"

    def _fused_ok(self, plan) -> bool:
        """The fused tile kernels need a hidden size divisible by 4, <= 32 bond features and reactions that fit a
        128-bond tile; other graphs (drug-like stress shape) run layer-wise with tensor-core GEMMs."""
        return (self.hidden_sizes[0] % 4 == 0 and self.num_edge_features <= 32 and plan.ensure_tiles())

    def _tc_weights(self, params, fa: int, fb: int) -> torch.Tensor:
        from . import ops
        key = tuple((p.data_ptr(), p._version) for p in params)
        cache = self.__dict__.get("_tc_cache")
        if cache is None or cache[0] != key:
            buf = ops.prepare_tc_weights(params, self.depth, _act_id(self.activation_fn),
                                         bool(self.use_learnable_skip), fa, fb)
            cache = (key, buf)
            self.__dict__["_tc_cache"] = cache
        return cache[1]

    def _check_arch(self) -> None:
        hs = list(self.hidden_sizes[: self.depth])
        if len(hs) < self.depth or any(h != hs[0] for h in hs) or self.hidden_sizes[-1] != hs[0]:
            raise RuntimeError("all hidden_sizes must be equal (the reference adds h_0 to every layer, GNN.py:94-97)")
        if len(self.dropout_ps) < self.depth:
            raise IndexError("dropout_ps shorter than depth (reference GNN.py:101 would raise)")

    def forward(self, data) -> torch.Tensor:
        x, edge_index, edge_attr = data.x, data.edge_index, data.edge_attr
        if edge_attr is None:
            raise RuntimeError("data.edge_attr is None (the reference fails at GNN.py:86 on this input too)")
        self._check_arch()
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device: the CGR hot path only exists as sm_100a kernels (no CPU fallback)")
        caller_device = x.device
        if caller_device.type == "cpu" and not (torch.is_grad_enabled() and any(
                p.requires_grad for p in self.parameters())) and not (self.training and any(
                    float(p) > 0 for p in self.dropout_ps[: self.depth])) and getattr(self, "engine", "auto") != "simt" \
                and not self._tc_demoted():
            out = self._infer_host_chunked(data)    # large host batch: pipelined over slices of whole reactions
            if out is None:
                out = self._infer_host(data)        # one C call on the host buffers (end-to-end entry)
            if out is not None:
                return out
        return self._forward_device(data, caller_device)

    def _forward_device(self, data, caller_device, force_simt: bool = False) -> torch.Tensor:
        x, edge_index, edge_attr = data.x, data.edge_index, data.edge_attr
        params = self._param_list()
        pdev = params[0].device
        dev = pdev if pdev.type == "cuda" else (caller_device if caller_device.type == "cuda"
                                                else torch.device("cuda", torch.cuda.current_device()))
        if pdev != dev:
            # CPU-resident module (reference CLI maps the checkpoint to CPU): use a device mirror
            params = self._device_mirror(dev)
        if caller_device != dev:
            data = _stage_to_device(data, dev)
            x, edge_index, edge_attr = data.x, data.edge_index, data.edge_attr
        plan = plan_for(data)
        if self.validate_inputs:
            plan.check()
        from . import ops
        needs_grad = bool(torch.is_grad_enabled() and any(p.requires_grad for p in params))
        seed = (torch.initial_seed() * 0x9E3779B97F4A7C15 + next(_seed_counter)) & 0x7FFFFFFFFFFFFFFF
        # dropout only in train mode (GNN.py:100-102); activations are saved whenever a backward may follow
        dps = [float(p) if self.training else 0.0 for p in self.dropout_ps[: self.depth]]
        train_flag = needs_grad or any(p > 0 for p in dps)
        engine = _lib.ENGINE_SIMT if force_simt else self._engine_id(plan, train_flag)
        fast = self._fast() and engine == _lib.ENGINE_TC and not train_flag
        empty_i = torch.empty(0, dtype=torch.int32, device=dev)
        fused_train = False
        if engine == _lib.ENGINE_TC:
            # training with the tile-local fused kernels: ReLU networks on tileable batches
            fused_train = bool(train_flag and getattr(self, "engine", "auto") != "tc_layerwise"
                               and _act_id(self.activation_fn) == 0 and self.hidden_sizes[0] <= 1024
                               and self.depth <= 12 and self._fused_ok(plan))
            if (train_flag and not fused_train) or not self._fused_ok(plan):
                # layer-wise path with tensor-core GEMMs: no tile plan needed
                tile_info, n_tiles = empty_i, 0
                if plan.tc_status is None:
                    plan.tc_status = torch.zeros(2, dtype=torch.int32, device=dev)
                tc_status = plan.tc_status
            else:
                tile_info, n_tiles, tc_status = plan.tile_info, plan.n_tiles, plan.tc_status
            if fused_train:
                # weights change every optimizer step: the fused training forward prepares them itself, next to the
                # saved activations, and the backward reads them from there (CUDA-graph safe: nothing cached here)
                tc_w = torch.empty(0, dtype=torch.uint8, device=dev)
            elif train_flag:
                tc_w = ops.prepare_tc_weights([p.detach() for p in params], self.depth, _act_id(self.activation_fn),
                                              bool(self.use_learnable_skip), int(x.shape[1]), int(edge_attr.shape[1]))
            else:
                tc_w = self._tc_weights([p.detach() for p in params], int(x.shape[1]), int(edge_attr.shape[1]))
            x_hi, x_lo = split_features_for(data, plan)
        else:
            tile_info, n_tiles, tc_status = empty_i, 0, empty_i
            tc_w = torch.empty(0, dtype=torch.uint8, device=dev)
            x_hi = x_lo = torch.empty(0, dtype=torch.float16, device=dev)
        throughput = getattr(self, "tile_policy", "latency") == "throughput"
        if needs_grad:
            call = dict(src=plan.src, dst=plan.dst, in_ptr=plan.in_ptr, in_idx=plan.in_idx, atom_ptr=plan.atom_ptr,
                        depth=self.depth, act=_act_id(self.activation_fn), use_skip=bool(self.use_learnable_skip),
                        dropout_ps=dps, seed=seed, engine=engine, tile_info=tile_info, n_tiles=n_tiles,
                        tc_status=tc_status, tc_weights=tc_w, x_hi=x_hi, x_lo=x_lo, tc_throughput=throughput,
                        fused_train=fused_train)
            out = ops.GnnFunction.apply(call, x, edge_attr, *params)
        else:
            out = ops.gnn_forward_impl(x, edge_attr, plan.src, plan.dst, plan.in_ptr, plan.in_idx, plan.atom_ptr,
                                       [p.detach() for p in params], self.depth, _act_id(self.activation_fn),
                                       bool(self.use_learnable_skip), dps, train_flag, seed,
                                       _lib.ENGINE_TC_FAST if (fast and n_tiles > 0) else engine, tile_info,
                                       n_tiles, tc_status, tc_w, x_hi, x_lo, throughput, fused_train)[0]
        self.__dict__["_last_plan"] = plan if (engine == _lib.ENGINE_TC and n_tiles > 0) else None
        self.__dict__["_last_engine"] = engine
        self.__dict__["_last_fused_train"] = fused_train
        if engine == _lib.ENGINE_TC and not torch.cuda.is_current_stream_capturing():
            with torch.cuda.device(dev):
                pending = self._queue_overflow_check(tc_status)
            if needs_grad or caller_device != dev:
                # a synchronisation follows anyway (loss.item() at trainer.py:145-147 / the copy to the caller's device):
                # read the flag here and, if it is set, repeat this step on the exact-fp32 engine
                pending[0].synchronize()
                bits = int(pending[1][pending[2]])
                if bits & 3:
                    self._demote(self._weights_key(), bits)
                    return self._forward_device(data, caller_device, force_simt=True)
            else:
                self.__dict__.setdefault("_ovf_pending", []).append(pending)
        if caller_device != dev:
            out = out.to(caller_device)
        return out

    # ------------------------------------------------------------------------------------------
    def _device_mirror(self, dev) -> List[torch.Tensor]:
        ps = self._param_list()
        key = tuple((p.data_ptr(), p._version) for p in ps) + (str(dev),)
        cache = self.__dict__.get("_mirror_cache")
        if cache is None or cache[0] != key:
            cache = (key, [p.detach().to(dev) for p in ps])
            self.__dict__["_mirror_cache"] = cache
        return cache[1]

    # ------------------------------------------------------------------ host-buffer inference ----
    def _host_ctx(self, fa: int, fb: int):
        """ctypes parameter block + prepared tcgen05 weights, cached per parameter version."""
        from . import ops
        params = self._param_list()
        key = tuple((p.data_ptr(), p._version) for p in params) + (fa, fb, self._fast())
        cache = self.__dict__.get("_host_ctx_cache")
        if cache is not None and cache[0] == key:
            return cache[1], cache[2]
        pdev = params[0].device
        dev = pdev if pdev.type == "cuda" else torch.device("cuda", torch.cuda.current_device())
        dparams = params if pdev == dev else self._device_mirror(dev)
        dparams = [ops._f32c(p.detach()) for p in dparams]
        tc_w = self._tc_weights(dparams, fa, fb)
        ctx = ops._Ctx(dparams, self.depth, _act_id(self.activation_fn), bool(self.use_learnable_skip), fa, fb,
                       [0.0] * self.depth)
        ctx.params.tc_weights = tc_w.data_ptr()
        ctx._keep = (dparams, tc_w)
        ctx.params.tc_throughput = 0
        ctx.params.tc_fast = int(self._fast())
        self.__dict__["_host_ctx_cache"] = (key, ctx, dev)
        return ctx, dev

    def _host_slot(self, slot: int, ctx, dev, n: int, e: int, b: int):
        import ctypes as C
        lib = _lib.load()
        slots = self.__dict__.setdefault("_host_slots", {})
        cur = slots.get(slot)
        dev_b, host_b = C.c_size_t(), C.c_size_t()
        _lib.check(lib.cgr_infer_host_workspace(C.byref(ctx.params), n, e, b, C.byref(dev_b), C.byref(host_b)),
                   "cgr_infer_host_workspace")
        if cur is None or cur[0].numel() < dev_b.value or cur[1].numel() < host_b.value or cur[2].numel() < b \
                or cur[0].device != dev:
            cur = (torch.empty(int(dev_b.value * 1.25) + 4096, dtype=torch.uint8, device=dev),
                   torch.empty(int(host_b.value * 1.25) + 4096, dtype=torch.uint8).pin_memory(),
                   torch.empty(max(b, 64) * 2, dtype=torch.float32).pin_memory(),
                   torch.cuda.Stream(device=dev))
            slots[slot] = cur
        return cur

    @staticmethod
    def _host_fields(data):
        x, ei, ea = data.x, data.edge_index, data.edge_attr
        batch, ptr = getattr(data, "batch", None), getattr(data, "ptr", None)
        if x.dtype != torch.float32 or ea.dtype != torch.float32 or ei.dtype != torch.int64:
            return None
        x, ei, ea = x.contiguous(), ei.contiguous(), ea.contiguous()
        if ptr is not None:
            ptr = ptr.contiguous()
            b = int(ptr.numel()) - 1
        elif batch is not None:
            batch = batch.contiguous()
            b = int(batch[-1]) + 1          # sorted ascending (PyG collate)
        else:
            b = 1
        return x, ei, ea, batch, ptr, int(x.shape[0]), int(ei.shape[1]), b

    def _host_supported(self) -> bool:
        return not (self.hidden_sizes[0] % 4 or self.depth > 13 or self.num_edge_features > 32)

    def _infer_host(self, data):
        """Inference on HOST tensors through ``cgr_gnn_infer_host``: H2D staging, index arrays, tcgen05
        forward and D2H of the energies in one call.  Returns None when the batch is not tileable (the
        generic path then handles it)."""
        import ctypes as C
        if not self._host_supported():
            return None
        f = self._host_fields(data)
        if f is None:
            return None
        x, ei, ea, batch, ptr, n, e, b = f
        lib = _lib.load()
        ctx, dev = self._host_ctx(int(x.shape[1]), int(ea.shape[1]))
        with torch.cuda.device(dev):
            dws, hws, hout, _ = self._host_slot(0, ctx, dev, n, e, b)
            ctx.params.tc_throughput = 0
            rc = lib.cgr_gnn_infer_host(C.byref(ctx.params), x.data_ptr(), ea.data_ptr(), ei.data_ptr(),
                                        _lib.ptr(ptr), _lib.ptr(batch), n, e, b, hout.data_ptr(), dws.data_ptr(),
                                        dws.numel(), hws.data_ptr(), hws.numel(),
                                        _lib.current_stream_handle())
        if rc == -3:            # not tileable: generic path (layer-wise kernels); fp16 range: exact-fp32 engine
            msg = lib.cgr_last_error_string() or b""
            if b"fp16 range" in msg:
                self._demote(self._weights_key(), 1)
            return None
        _lib.check(rc, "cgr_gnn_infer_host")
        return hout[:b].clone()

    #: a host batch of at least this many reactions is split into slices of HOST_CHUNK whole reactions that travel through
    #: ``predict_stream`` (copies of one slice overlap the kernels of another: 0.72 -> ~0.9 M reactions/s at 8192
    #: reactions with Fa = 846, where the one-call entry first copies 450 MB and only then computes)
    HOST_CHUNK_MIN = 4096
    HOST_CHUNK = 1024

    def _infer_host_chunked(self, data):
        """Reactions are independent, so a big collated host batch is a list of smaller ones: views of ``x`` /
        ``edge_attr`` rows, ``edge_index`` / ``ptr`` rebased to the slice.  Returns None when it does not apply."""
        if not self._host_supported():
            return None
        ptr, batch = getattr(data, "ptr", None), getattr(data, "batch", None)
        if ptr is None and batch is None:
            return None
        b = int(ptr.numel()) - 1 if ptr is not None else int(batch[-1]) + 1
        if b < self.HOST_CHUNK_MIN:
            return None
        f = self._host_fields(data)
        if f is None:
            return None
        x, ei, ea, batch, ptr, n, e, b = f
        if ptr is None:
            ptr = torch.zeros(b + 1, dtype=torch.int64)
            ptr[1:] = torch.bincount(batch, minlength=b).cumsum(0)
        src = ei[0].numpy()
        chunks, e0 = [], 0
        for b0 in range(0, b, self.HOST_CHUNK):
            b1 = min(b, b0 + self.HOST_CHUNK)
            a0, a1 = int(ptr[b0]), int(ptr[b1])
            # bonds of a collated batch are grouped by reaction: the slice ends at the first bond of a later reaction
            # (source atom >= a1) -- a bisection on that monotone predicate, 18 probes instead of a pass over every bond
            e1 = e
            if b1 < b:
                lo_, hi_ = e0, e
                while lo_ < hi_:
                    mid = (lo_ + hi_) >> 1
                    if src[mid] < a1:
                        lo_ = mid + 1
                    else:
                        hi_ = mid
                e1 = lo_
            if e1 <= e0 or a1 <= a0:
                return None
            lp = ptr[b0:b1 + 1] - a0
            lb = batch[a0:a1] - b0 if batch is not None else torch.repeat_interleave(torch.arange(b1 - b0), lp[1:] - lp[:-1])
            chunks.append(_HostSlice(x[a0:a1], ei[:, e0:e1] - a0, ea[e0:e1], lp, lb))
            e0 = e1
        outs = list(self.predict_stream(chunks, depth=4, workers=2, coalesce=1))
        return torch.cat([o.reshape(-1) for o in outs])

    def forward_group(self, batches) -> list:
        """Energies of several independent device-resident batches (a screening job is a stream of them), inference
        only: ``[self(b) for b in batches]`` in TWO launches per group of up to 24 batches instead of two per batch --
        one atom projection over every batch's atom tiles, one fused cluster kernel over every batch's tile groups
        (``cgr_gnn_forward_group``).  Same kernels and arithmetic as ``forward`` with ``tile_policy="throughput"``: the
        energies are bit-identical to it.  Batches the tcgen05 fused kernels cannot take (reactions of more than 128
        directed bonds, engine "simt", a weight version demoted by the fp16-range guard) fall back to per-batch
        ``forward`` calls -- inside CUDA, never on the CPU."""
        from . import ops
        batches = list(batches)
        if not batches:
            return []
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise RuntimeError("forward_group is an inference entry: call it under torch.no_grad()")
        self._check_arch()
        params = self._param_list()
        dev = params[0].device
        if dev.type != "cuda" or any(b.x.device != dev for b in batches):
            raise RuntimeError("forward_group needs the module and every batch on the same CUDA device")
        plans = [plan_for(b) for b in batches]
        if self.validate_inputs:
            for pl in plans:
                pl.check()
        if (getattr(self, "engine", "auto") in ("simt", 0, "tc_layerwise") or self._tc_demoted() or self.depth > 13
                or not all(self._fused_ok(pl) for pl in plans)):
            return [self._forward_device(b, dev) for b in batches]
        fa, fb = int(batches[0].x.shape[1]), int(batches[0].edge_attr.shape[1])
        tc_w = self._tc_weights([p.detach() for p in params], fa, fb)
        outs = []
        with torch.cuda.device(dev):
            lo = 0
            while lo < len(batches):
                # a call takes up to GROUP_MAX batches; a batch that appears twice goes to the next call (its status
                # words -- readout tickets, range flag -- belong to one forward at a time)
                graphs, seen = [], set()
                while lo < len(batches) and len(graphs) < ops.GROUP_MAX and id(plans[lo]) not in seen:
                    b, pl = batches[lo], plans[lo]
                    seen.add(id(pl))
                    x_hi, x_lo = split_features_for(b, pl)
                    graphs.append((b.x, b.edge_attr, pl, x_hi, x_lo))
                    lo += 1
                outs += ops.gnn_forward_group(graphs, [p.detach() for p in params], self.depth,
                                              _act_id(self.activation_fn), bool(self.use_learnable_skip), tc_w,
                                              fast=self._fast())
            if not torch.cuda.is_current_stream_capturing():
                for pl in plans:
                    self.__dict__.setdefault("_ovf_pending", []).append(self._queue_overflow_check(pl.tc_status))
        return outs

    def predict_stream(self, batches, depth: int = 4, workers: int = 2, coalesce: int = 8,
                       coalesce_reactions: int = 1024):
        """Pipelined inference over an iterable of HOST batches (the screening workload): yields one CPU
        tensor of energies per batch, in order.  Up to ``coalesce`` consecutive batches (closed early once they
        hold ``coalesce_reactions`` reactions, which bounds the workspace) travel in ONE submission (``cgr_gnn_infer_host_multi_async``: each is staged straight from its own host buffers
        into a device-side super-batch, so copies and launches are amortised; per-reaction results equal
        separate submissions up to fp32 rounding of the final column sum).  Up to ``depth`` submissions are in flight, each on its own
        stream with its own staging buffers, so the H2D copies of one overlap the kernels of another;
        ``workers`` host threads issue the calls (the C entry releases the GIL).  Batches that cannot use
        the tcgen05 engine go through ``forward``."""
        import ctypes as C
        from collections import deque
        from concurrent.futures import ThreadPoolExecutor
        lib = _lib.load()
        depth, coalesce = max(1, depth), max(1, coalesce)
        pending = deque()
        pool = ThreadPoolExecutor(max_workers=max(1, workers))

        def submit(slot_id, group):
            fields = [self._host_fields(d) for d in group]
            if any(f is None for f in fields):
                return None
            k = len(fields)
            arr = (_lib.CgrHostBatch * k)()
            n = e = b = 0
            for hb, (x, ei, ea, batch, ptr, nj, ej, bj) in zip(arr, fields):
                hb.x, hb.edge_attr, hb.edge_index = x.data_ptr(), ea.data_ptr(), ei.data_ptr()
                hb.ptr, hb.batch = _lib.ptr(ptr), _lib.ptr(batch)
                hb.n_atoms, hb.n_bonds, hb.n_rxn = nj, ej, bj
                n += nj; e += ej; b += bj
            x0, ea0 = fields[0][0], fields[0][2]
            ctx, dev = self._host_ctx(int(x0.shape[1]), int(ea0.shape[1]))
            with torch.cuda.device(dev):
                slot = self._host_slot(slot_id, ctx, dev, n, e, b)
                dws, hws, hout, st = slot
                ctx.params.tc_throughput = 1          # several submissions in flight
                rc = lib.cgr_gnn_infer_host_multi_async(C.byref(ctx.params), arr, k, hout.data_ptr(), dws.data_ptr(),
                                                        dws.numel(), hws.data_ptr(), hws.numel(), st.cuda_stream)
            if rc == -3:
                return None
            _lib.check(rc, "cgr_gnn_infer_host_multi_async")
            return slot, ctx, n, e, b, [f[7] for f in fields], (arr, fields)

        def finish():
            fut, group = pending.popleft()
            res = fut.result()
            if res is None:                         # not tileable: generic path
                return [self.forward(d) for d in group]
            slot, ctx, n, e, b, counts, _keep = res
            dws, hws, hout, st = slot
            st.synchronize()
            rc = lib.cgr_infer_host_check(C.byref(ctx.params), n, e, b, hws.data_ptr())
            if rc == -3:                            # an operand left the fp16 range: exact-fp32 engine for this group
                self._demote(self._weights_key(), 1)
                return [self.forward(d) for d in group]
            _lib.check(rc, "cgr_infer_host_check")
            return list(hout[:b].clone().split(counts))

        try:
            with torch.no_grad():
                if self._host_supported() and not self._tc_demoted():
                    first = True
                    i = 0
                    group, group_rxn = [], 0

                    def flush():
                        nonlocal i, group, group_rxn
                        if group:
                            pending.append((pool.submit(submit, 1 + i % depth, group), group))
                            i += 1
                            group, group_rxn = [], 0

                    for data in batches:
                        if data.x.device.type != "cpu" or data.edge_attr is None:
                            flush()
                            while pending:
                                yield from finish()
                            yield self.forward(data)
                            continue
                        if first:                    # build the cached parameter block outside the workers
                            self._host_ctx(int(data.x.shape[1]), int(data.edge_attr.shape[1]))
                            first = False
                        group.append(data)
                        ptr = getattr(data, "ptr", None)
                        group_rxn += int(ptr.numel()) - 1 if ptr is not None else 64
                        if len(group) >= coalesce or group_rxn >= coalesce_reactions:
                            if len(pending) >= depth:
                                yield from finish()
                            flush()
                    if len(pending) >= depth:
                        yield from finish()
                    flush()
                    while pending:
                        yield from finish()
                else:
                    for data in batches:
                        yield self.forward(data)
        finally:
            pool.shutdown(wait=True)

    def check_numerics(self) -> None:
        """Synchronising check of the tcgen05 forwards issued so far: raises if an operand left the fp16 range of the
        FP16x3 split (those forwards returned NaN energies; ``engine='auto'`` then continues on the exact-fp32 engine).
        The flag word is cleared by the next forward of the same batch."""
        self._poll_overflow(wait=True)
        plan = self.__dict__.get("_last_plan")
        bad = plan is not None and plan.tc_status is not None and (int(plan.tc_status[0].item()) & 3) != 0
        if bad or self.__dict__.get("_tc_demoted_key") == self._weights_key():
            raise RuntimeError("tcgen05 engine: an activation exceeded the fp16 range of the FP16x3 split; "
                               "set model.engine = 'simt' (engine='auto' has switched already)")

    def __getstate__(self):
        state = self.__dict__.copy()
        state.pop("_mirror_cache", None)     # torch.save(model) must not pickle device mirrors
        state.pop("_tc_cache", None)
        state.pop("_host_slots", None)
        state.pop("_ovf_events", None)
        state.pop("_host_ctx_cache", None)
        state.pop("_last_plan", None)
        state.pop("_plist", None)
        state.pop("_last_fused_train", None)
        state.pop("_ovf_pending", None)
        state.pop("_ovf_ring", None)
        state.pop("_ovf_next", None)
        state.pop("_tc_demoted_key", None)
        return state


class _HostSlice:
    """A run of whole reactions of a collated host batch, with the fields ``GNN.forward`` reads."""

    def __init__(self, x, edge_index, edge_attr, ptr, batch):
        self.x, self.edge_index, self.edge_attr, self.ptr, self.batch = x, edge_index, edge_attr, ptr, batch


def _stage_to_device(data, dev):
    """Host -> device staging of one batch (pinned when possible); keeps the duck-typed interface."""
    from .data import Batch

    def mv(t):
        if t is None:
            return None
        if not t.is_cuda and not t.is_pinned():
            try:
                t = t.pin_memory()
            except RuntimeError:
                pass
        return t.to(dev, non_blocking=True)

    return Batch(mv(data.x), mv(data.edge_index), mv(data.edge_attr), mv(getattr(data, "batch", None)),
                 mv(getattr(data, "ptr", None)), mv(getattr(data, "y", None)))
