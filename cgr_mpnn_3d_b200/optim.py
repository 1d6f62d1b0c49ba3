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
