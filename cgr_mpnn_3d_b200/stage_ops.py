"""Thin Python wrappers of the stage-level C entry points (one per north-star subsystem).

Used by the parity tests to check each stage against the oracle in isolation, and by
``DMPNNConv.forward`` when a conv layer is called on its own.  Inference-only (no autograd); the
differentiable path is the fused ``cgr_b200::gnn_forward`` op.
"""
from __future__ import annotations

import torch

from . import _lib
from .collate import GraphPlan, build_plan


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _f32(t):
    return t.detach().float().contiguous()


def edge_init_fwd(x, edge_attr, plan: GraphPlan, weight, bias, act: int, want_z: bool = False):
    """reference GNN.py:85-86."""
    lib = _lib.load()
    x, edge_attr, weight, bias = _f32(x), _f32(edge_attr), _f32(weight), _f32(bias)
    n, fa = x.shape
    e, fb = edge_attr.shape
    H = weight.shape[0]
    h0 = torch.empty((e, H), dtype=torch.float32, device=x.device)
    z0 = torch.empty_like(h0) if want_z else None
    ws = torch.empty(n * H + 64, dtype=torch.float32, device=x.device)
    _lib.check(lib.cgr_edge_init_fwd(x.data_ptr(), edge_attr.data_ptr(), plan.src.data_ptr(), weight.data_ptr(),
                                     bias.data_ptr(), n, e, fa, fb, H, act, h0.data_ptr(), _lib.ptr(z0),
                                     ws.data_ptr(), ws.numel() * 4, _stream()), "cgr_edge_init_fwd")
    return (h0, z0) if want_z else h0


def bond_update_fwd(h_in, h0, plan: GraphPlan, weight, bias, skip, act: int, dropout_p: float = 0.0,
                    seed: int = 0, layer: int = 0, training: bool = False):
    """reference GNN.py:91-102.  Returns (h_out, m, z)."""
    lib = _lib.load()
    h_in, h0, weight, bias = _f32(h_in), _f32(h0), _f32(weight), _f32(bias)
    skip = None if skip is None else _f32(skip)
    e, H = h_in.shape
    h_out, m, z = torch.empty_like(h_in), torch.empty_like(h_in), torch.empty_like(h_in)
    _lib.check(lib.cgr_bond_update_fwd(h_in.data_ptr(), h0.data_ptr(), plan.in_ptr.data_ptr(),
                                       plan.in_idx.data_ptr(), plan.src.data_ptr(), weight.data_ptr(),
                                       bias.data_ptr(), _lib.ptr(skip), act, float(dropout_p), seed, layer,
                                       int(training), h_out.data_ptr(), m.data_ptr(), z.data_ptr(), e,
                                       plan.n_atoms, H, _stream()), "cgr_bond_update_fwd")
    return h_out, m, z


def readout_fwd(h, x, plan: GraphPlan, w_e2n, b_e2n, w_ffn, b_ffn, act: int):
    """reference GNN.py:105-110.  Returns (out, s, hv, pooled)."""
    lib = _lib.load()
    h, x, w_e2n, b_e2n, w_ffn, b_ffn = (_f32(t) for t in (h, x, w_e2n, b_e2n, w_ffn, b_ffn))
    n, fa = x.shape
    e, H = h.shape
    b = plan.n_rxn
    f32 = dict(dtype=torch.float32, device=x.device)
    out, s, hv, pooled = torch.empty(b, **f32), torch.empty((n, H), **f32), torch.empty((n, H), **f32), \
        torch.empty((b, H), **f32)
    _lib.check(lib.cgr_readout_fwd(h.data_ptr(), x.data_ptr(), plan.in_ptr.data_ptr(), plan.in_idx.data_ptr(),
                                   plan.atom_ptr.data_ptr(), w_e2n.data_ptr(), b_e2n.data_ptr(), w_ffn.data_ptr(),
                                   b_ffn.data_ptr(), act, out.data_ptr(), s.data_ptr(), hv.data_ptr(), None,
                                   pooled.data_ptr(), n, e, b, fa, H, _stream()), "cgr_readout_fwd")
    return out, s, hv, pooled


def conv_forward(conv, edge_index, edge_attr):
    """``DMPNNConv.forward(edge_index, edge_attr) -> (a_message, lin(a[row] - rev))`` (GNN.py:131-141)."""
    if not edge_attr.is_cuda:
        raise RuntimeError("DMPNNConv.forward needs CUDA tensors: no CPU implementation")
    lib = _lib.load()
    h = _f32(edge_attr)
    e, H = h.shape
    n = int(edge_index[1].max()) + 1      # PyG propagate sizing (x=None, size=None), one sync as in PyG
    plan = build_plan(edge_index, n, None)
    w, b = _f32(conv.lin.weight), _f32(conv.lin.bias)
    a = torch.empty((n, H), dtype=torch.float32, device=h.device)
    y, m = torch.empty_like(h), torch.empty_like(h)
    _lib.check(lib.cgr_conv_fwd(h.data_ptr(), plan.in_ptr.data_ptr(), plan.in_idx.data_ptr(), plan.src.data_ptr(),
                                w.data_ptr(), b.data_ptr(), a.data_ptr(), y.data_ptr(), m.data_ptr(), e, n, H,
                                _stream()), "cgr_conv_fwd")
    return a, y


def dropout_mask(seed: int, layer: int, p: float, n_bonds: int, hidden: int, device) -> torch.Tensor:
    lib = _lib.load()
    mask = torch.empty((n_bonds, hidden), dtype=torch.uint8, device=device)
    _lib.check(lib.cgr_dropout_mask(seed, layer, float(p), n_bonds, hidden, mask.data_ptr(), _stream()),
               "cgr_dropout_mask")
    return mask.bool()


def mse_sum(pred, y):
    """reference train.py:120 ``MSELoss(reduction='sum')``: returns (loss, dL/dpred)."""
    lib = _lib.load()
    pred, y = _f32(pred), _f32(y)
    loss = torch.empty(1, dtype=torch.float32, device=pred.device)
    grad = torch.empty_like(pred)
    _lib.check(lib.cgr_mse_sum_fwd_bwd(pred.data_ptr(), y.data_ptr(), pred.numel(), loss.data_ptr(),
                                       grad.data_ptr(), _stream()), "cgr_mse_sum_fwd_bwd")
    return loss, grad


class _MSESumLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, y):
        loss, grad = mse_sum(pred.detach(), y.detach())
        ctx.save_for_backward(grad)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None


def mse_sum_loss(pred, y):
    """Drop-in for ``torch.nn.MSELoss(reduction="sum")(pred, y)`` (train.py:120, trainer.py:142) as ONE launch that
    produces the loss and dL/dpred together (``cgr_mse_sum_fwd_bwd``); the target gets no gradient."""
    return _MSESumLoss.apply(pred, y)
