"""Fused optimizer step for the hot path's parameters.

``FusedAdam`` is a drop-in for the optimizer the reference builds at train.py:117-119 --
``torch.optim.Adam(model.parameters(), lr=lr, weight_decay=weight_decay, amsgrad=True)`` -- and steps at
training/trainer.py:144: same constructor arguments, same ``param_groups`` (so ``ExponentialLR``, train.py:121, drives
it unchanged), same per-parameter state names (``step``, ``exp_avg``, ``exp_avg_sq``, ``max_exp_avg_sq``), so a
``state_dict`` moves between the two.  ``step()`` updates every parameter tensor in ONE kernel launch
(``cgr_adam_step``) instead of torch's dozen foreach launches.  CUDA parameters only: there is no CPU path.
"""
import ctypes as C

import torch

from . import _lib


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False,
                 grad_scale=1.0):
        if lr < 0.0:
            raise ValueError(f"Invalid learning rate: {lr}")
        if eps < 0.0:
            raise ValueError(f"Invalid epsilon value: {eps}")
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError(f"Invalid beta parameter at index 0: {betas[0]}")
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid beta parameter at index 1: {betas[1]}")
        if weight_decay < 0.0:
            raise ValueError(f"Invalid weight_decay value: {weight_decay}")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad)
        super().__init__(params, defaults)
        self.grad_scale = float(grad_scale)      # multiplies every gradient first (e.g. 1/world for a replica mean)
        self._fast = {}          # per group: [signature, ctypes table, items, step count] of the steady-state launch

    # The per-parameter "step" tensors torch keeps are refreshed lazily: in steady state the step count lives in
    # self._fast (one Python int per group) and is written back whenever the state is inspected or the set of
    # tensors changes.
    def _sync_steps(self):
        for fast in self._fast.values():
            for _, st in fast[2]:
                st["step"].fill_(float(fast[3]))

    def state_dict(self):
        self._sync_steps()
        return super().state_dict()

    def load_state_dict(self, state_dict):
        self._fast = {}
        return super().load_state_dict(state_dict)

    def _state_for(self, p, amsgrad):
        st = self.state[p]
        if len(st) == 0:
            st["step"] = torch.tensor(0.0, dtype=torch.float32)         # host scalar, like torch's default path
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            if amsgrad:
                st["max_exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        elif amsgrad and "max_exp_avg_sq" not in st:
            st["max_exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    def _launch(self, lib, group, arr, n, step, amsgrad, dev, params):
        beta1, beta2 = group["betas"]
        with torch.cuda.device(dev):
            rc = lib.cgr_adam_step(arr, n, float(group["lr"]), float(beta1), float(beta2), float(group["eps"]),
                                   float(group["weight_decay"]), step, int(amsgrad), self.grad_scale,
                                   torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "cgr_adam_step")
        # the kernel wrote the parameters through raw pointers: tell autograd (and every cache keyed on tensor versions,
        # e.g. the prepared tcgen05 weights of GNN) that they changed
        torch.autograd.graph.increment_version(params)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for gi, group in enumerate(self.param_groups):
            amsgrad = bool(group["amsgrad"])
            sig = [amsgrad]
            for p in group["params"]:
                g = p.grad
                if g is not None:
                    sig.append(p.data_ptr())
                    sig.append(g.data_ptr())
            fast = self._fast.get(gi)
            if fast is not None and fast[0] == sig:          # steady state: same tensors as the previous step
                fast[3] += 1
                self._launch(lib, group, fast[1], len(fast[2]), fast[3], amsgrad, fast[4], fast[5])
                continue
            if fast is not None:
                for _, st in fast[2]:
                    st["step"].fill_(float(fast[3]))
                del self._fast[gi]
            by_step = {}
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda:
                    raise RuntimeError("FusedAdam: parameters must live on a CUDA device (no CPU path); "
                                       "use torch.optim.Adam for CPU tensors")
                if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or p.grad.is_sparse:
                    raise RuntimeError("FusedAdam: dense fp32 parameters and gradients only")
                if not p.is_contiguous():
                    raise RuntimeError("FusedAdam: parameters must be contiguous")
                st = self._state_for(p, amsgrad)
                by_step.setdefault(int(st["step"].item()) + 1, []).append((p, st))
            for step, items in by_step.items():
                grads = [p.grad if p.grad.is_contiguous() else p.grad.contiguous() for p, _ in items]
                arr = (_lib.CgrAdamTensor * len(items))()
                for a, (p, st), g in zip(arr, items, grads):
                    a.param, a.grad = p.data_ptr(), g.data_ptr()
                    a.exp_avg, a.exp_avg_sq = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
                    a.max_exp_avg_sq = st["max_exp_avg_sq"].data_ptr() if amsgrad else None
                    a.numel = p.numel()
                dev = items[0][0].device
                self._launch(lib, group, arr, len(items), step, amsgrad, dev, [p for p, _ in items])
                for _, st in items:
                    st["step"] += 1
                if len(by_step) == 1 and all(g is p.grad for (p, _), g in zip(items, grads)):
                    self._fast[gi] = [sig, arr, items, step, dev, [p for p, _ in items]]
        return loss


class PeerFusedAdam(torch.optim.Optimizer):
    """Data-parallel Adam without a collective library: ``step()`` is ONE kernel (``cgr_peer_allreduce_adam``) that waits
    for all replicas, SUMS their gradients by reading every replica's gradient arena over NVLink (CUDA-IPC-mapped peer
    memory, fixed rank order: replicas stay bit-identical) and applies the Adam / amsgrad update of train.py:117-119.

    One process per GPU on one node (``torch.distributed`` initialised, any backend: it is only used once, to exchange the
    IPC handles).  The CGR backward writes its flat gradient buffer straight into the arena (``ops.set_grad_arena``), so
    there is no packing copy either.  Gradients are SUMMED over replicas (the reference's loss is
    ``MSELoss(reduction="sum")``); pass ``grad_scale=1/world`` for a mean.  All parameters must get a gradient every step
    (true for the CGR model); all replicas must call ``step()`` the same number of times.
    """

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False, grad_scale=1.0,
                 group=None, two_shot=None):
        import torch.distributed as dist
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad)
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise ValueError("PeerFusedAdam supports one parameter group")
        ps = self.param_groups[0]["params"]
        if not ps or any((not p.is_cuda) or p.dtype != torch.float32 or not p.is_contiguous() for p in ps):
            raise RuntimeError("PeerFusedAdam: contiguous fp32 CUDA parameters only (no CPU path)")
        if len(ps) > 48:
            raise ValueError("PeerFusedAdam supports at most 48 parameter tensors")
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerFusedAdam needs torch.distributed (one process per GPU of one node)")
        self.grad_scale = float(grad_scale)
        self._group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 16:
            raise ValueError("PeerFusedAdam supports up to 16 replicas")
        self.device = ps[0].device
        self._sizes = [(p.numel() + 3) // 4 * 4 for p in ps]
        self._total = sum(self._sizes)
        lib = _lib.load()
        with torch.cuda.device(self.device):
            # [arena 0 | arena 1 | reduced slices | flag pad]: one block, exported to the peers through CUDA IPC
            self._shared = torch.zeros(3 * self._total + 64, dtype=torch.float32, device=self.device)
            torch.cuda.synchronize(self.device)
            handle = C.create_string_buffer(64)
            offset = C.c_int64(0)
            _lib.check(lib.cgr_ipc_export(self._shared.data_ptr(), handle, C.byref(offset)), "cgr_ipc_export")
            gathered = [None] * self.world
            dist.all_gather_object(gathered, (self.device.index, handle.raw, int(offset.value)), group=group)
            self._peer_ptrs = []                  # address of every rank's shared block as seen from this device
            for r, (peer_dev, h, off) in enumerate(gathered):
                if r == self.rank:
                    self._peer_ptrs.append(self._shared.data_ptr())
                else:
                    _lib.check(lib.cgr_enable_peer_access(int(peer_dev)), "cgr_enable_peer_access")
                    ptr = C.c_void_p()
                    _lib.check(lib.cgr_ipc_open(C.create_string_buffer(h, 64), off, C.byref(ptr)), "cgr_ipc_open")
                    self._peer_ptrs.append(int(ptr.value))
            dist.barrier(group=group)
        self._arena_ptrs = []
        for buf in (0, 1):
            arr = (C.c_void_p * self.world)(*[q + buf * self._total * 4 for q in self._peer_ptrs])
            self._arena_ptrs.append(arr)
        self._reduced_ptrs = (C.c_void_p * self.world)(*[q + 2 * self._total * 4 for q in self._peer_ptrs])
        self._flag_ptrs = (C.c_void_p * self.world)(*[q + 3 * self._total * 4 for q in self._peer_ptrs])
        # one-shot: every rank reads all W arenas (W-1 gradient sizes per GPU); two-shot: reduce one slice each, then
        # gather (2 (W-1)/W sizes) -- the better trade from about 4 replicas up
        import os
        env = os.environ.get("CGR_PEER_TWO_SHOT")
        self.two_shot = bool(two_shot) if two_shot is not None else (env == "1" if env is not None else self.world > 4)
        self._cur = 0                             # arena the next backward writes into
        self._sync_step = 0
        self._adam_step = 0
        self._handed_out = 0                      # sync step whose arena a backward already received
        self._param_ptrs = [p.data_ptr() for p in ps]
        for p in ps:
            st = self.state[p]
            st["step"] = torch.tensor(0.0, dtype=torch.float32)
            st["exp_avg"] = torch.zeros_like(p)
            st["exp_avg_sq"] = torch.zeros_like(p)
            if amsgrad:
                st["max_exp_avg_sq"] = torch.zeros_like(p)
        self._table = None
        from . import ops
        ops.set_grad_arena(self._provide)

    def _provide(self, n_floats, device, params=None):
        """Arena of the current step for ONE backward: a second backward before ``step()`` (gradient accumulation, two
        models of the same shape) gets a private buffer instead of overwriting the first one's gradients in place
        (``step()`` then reports the misplaced gradient)."""
        if n_floats != self._total or device != self.device or self._handed_out == self._sync_step + 1:
            return None
        if params is not None and [q.data_ptr() for q in params] != self._param_ptrs:
            return None                            # another model of the same shape: not this optimizer's gradients
        if not torch.cuda.is_current_stream_capturing():
            self._handed_out = self._sync_step + 1
        return self._shared[self._cur * self._total:(self._cur + 1) * self._total]

    def state_dict(self):
        for p in self.param_groups[0]["params"]:
            self.state[p]["step"].fill_(float(self._adam_step))
        return super().state_dict()

    def load_state_dict(self, state_dict):
        """Resume: torch installs NEW exp_avg / exp_avg_sq / max_exp_avg_sq tensors, so the cached pointer table is
        rebuilt, and the bias-correction step continues from the saved ``step`` (identical on every replica: they
        all load the same state).  The flag protocol's own counter keeps running -- it only has to agree between
        replicas of this process group, which never stopped."""
        super().load_state_dict(state_dict)
        self._table = None
        steps = {int(float(st["step"])) for st in self.state.values() if "step" in st}
        if len(steps) > 1:
            raise RuntimeError("PeerFusedAdam: parameters carry different step counts; one step per parameter set only")
        self._adam_step = steps.pop() if steps else 0
        for p in self.param_groups[0]["params"]:
            st = self.state[p]
            st["step"] = torch.tensor(float(self._adam_step), dtype=torch.float32)
            for k in ("exp_avg", "exp_avg_sq", "max_exp_avg_sq"):
                if k in st and (st[k].device != p.device or not st[k].is_contiguous() or st[k].dtype != torch.float32):
                    st[k] = st[k].to(device=p.device, dtype=torch.float32).contiguous()
            if self.param_groups[0]["amsgrad"] and "max_exp_avg_sq" not in st:
                st["max_exp_avg_sq"] = torch.zeros_like(p)

    @torch.no_grad()
    def step(self, closure=None, arena=None):
        """``arena`` (0 / 1): only for CUDA-graph replay, where ``p.grad`` objects are not refreshed -- names the arena the
        replayed backward wrote (the one that was current when that graph was captured); must agree on all replicas."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        group = self.param_groups[0]
        ps = group["params"]
        amsgrad = bool(group["amsgrad"])
        if arena is not None:
            self._cur = int(arena) & 1
        else:
            base = self._shared.data_ptr() + self._cur * self._total * 4
            off = 0
            for p, n in zip(ps, self._sizes):      # every gradient must sit at its slot of the current arena
                g = p.grad
                if g is None or g.data_ptr() != base + off * 4:
                    raise RuntimeError("PeerFusedAdam: a gradient is missing or was not written into the shared arena "
                                       "(use zero_grad(set_to_none=True) and one backward per step)")
                off += n
        if self._table is None:
            arr = (_lib.CgrAdamTensor * len(ps))()
            off = 0
            for a, p, n in zip(arr, ps, self._sizes):
                st = self.state[p]
                a.param, a.grad = p.data_ptr(), off                 # grad = offset inside the arenas, in floats
                a.exp_avg, a.exp_avg_sq = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
                a.max_exp_avg_sq = st["max_exp_avg_sq"].data_ptr() if amsgrad else None
                a.numel = p.numel()
                off += n
            self._table = arr
        self._sync_step += 1
        self._adam_step += 1
        beta1, beta2 = group["betas"]
        with torch.cuda.device(self.device):
            rc = lib.cgr_peer_allreduce_adam(self._table, len(ps), self._arena_ptrs[self._cur], self._flag_ptrs,
                                             self._reduced_ptrs if self.two_shot else None, self._total, self.world,
                                             self.rank, self._sync_step, float(group["lr"]), float(beta1), float(beta2),
                                             float(group["eps"]), float(group["weight_decay"]), self._adam_step,
                                             int(amsgrad), self.grad_scale,
                                             torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(rc, "cgr_peer_allreduce_adam")
        torch.autograd.graph.increment_version(ps)
        self._cur ^= 1                               # the next backward writes the other arena
        return loss
