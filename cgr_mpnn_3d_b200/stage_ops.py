"""Thin Python wrappers of the stage-level C entry points (one per north-star subsystem).

Used by the parity tests to check each stage (forward and backward) against the oracle in isolation, and by
``DMPNNConv.forward`` when a conv layer is called on its own.  No autograd is attached here; the
differentiable path is the fused ``cgr_b200::gnn_forward`` op.
"""
from __future__ import annotations

import torch

from . import _lib
from .collate import GraphPlan, build_plan


def _stream() -> int:
    return _lib.current_stream_handle()


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


def readout_fwd(h, x, plan: GraphPlan, w_e2n, b_e2n, w_ffn, b_ffn, act: int, want_z: bool = False):
    """reference GNN.py:105-110.  Returns (out, s, hv, pooled[, zv])."""
    lib = _lib.load()
    h, x, w_e2n, b_e2n, w_ffn, b_ffn = (_f32(t) for t in (h, x, w_e2n, b_e2n, w_ffn, b_ffn))
    n, fa = x.shape
    e, H = h.shape
    b = plan.n_rxn
    f32 = dict(dtype=torch.float32, device=x.device)
    out, s, hv, pooled = torch.empty(b, **f32), torch.empty((n, H), **f32), torch.empty((n, H), **f32), \
        torch.empty((b, H), **f32)
    zv = torch.empty((n, H), **f32) if want_z else None
    _lib.check(lib.cgr_readout_fwd(h.data_ptr(), x.data_ptr(), plan.in_ptr.data_ptr(), plan.in_idx.data_ptr(),
                                   plan.atom_ptr.data_ptr(), w_e2n.data_ptr(), b_e2n.data_ptr(), w_ffn.data_ptr(),
                                   b_ffn.data_ptr(), act, out.data_ptr(), s.data_ptr(), hv.data_ptr(), _lib.ptr(zv),
                                   pooled.data_ptr(), n, e, b, fa, H, _stream()), "cgr_readout_fwd")
    return (out, s, hv, pooled, zv) if want_z else (out, s, hv, pooled)


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


# ---------------------------------------------------------------------------------------------
# stage-level backward (explicit mirror of autograd, one call per stage)
def _bwd_ws(lib, n, e, fa, fb, H, device):
    nbytes = int(lib.cgr_stage_bwd_workspace(n, e, fa, fb, H))
    return torch.empty(nbytes, dtype=torch.uint8, device=device), nbytes


def readout_bwd(grad_out, x, plan: GraphPlan, w_e2n, w_ffn, act: int, s, hv, pooled, zv=None):
    """Backward of :func:`readout_fwd`: returns (gw_e2n, gb_e2n, gw_ffn, gb_ffn, dh) with ``dh`` [E, H] the gradient
    w.r.t. the bond states the readout consumed."""
    lib = _lib.load()
    grad_out, x, w_e2n, w_ffn, s, hv, pooled = (_f32(t) for t in (grad_out, x, w_e2n, w_ffn, s, hv, pooled))
    zv = None if zv is None else _f32(zv)
    n, fa = x.shape
    H = hv.shape[1]
    e, b = plan.n_bonds, plan.n_rxn
    f32 = dict(dtype=torch.float32, device=x.device)
    gw, gb, gwf, gbf = torch.empty((H, fa + H), **f32), torch.empty(H, **f32), torch.empty((1, H), **f32), torch.empty(1, **f32)
    dh = torch.empty((e, H), **f32)
    ws, nbytes = _bwd_ws(lib, n, e, fa, 0, H, x.device)
    _lib.check(lib.cgr_readout_bwd(grad_out.data_ptr(), x.data_ptr(), plan.in_ptr.data_ptr(), plan.in_idx.data_ptr(),
                                   plan.atom_ptr.data_ptr(), plan.dst.data_ptr(), w_e2n.data_ptr(), w_ffn.data_ptr(), act,
                                   s.data_ptr(), hv.data_ptr(), _lib.ptr(zv), pooled.data_ptr(), gw.data_ptr(),
                                   gb.data_ptr(), gwf.data_ptr(), gbf.data_ptr(), dh.data_ptr(), n, e, b, fa, H,
                                   ws.data_ptr(), nbytes, _stream()), "cgr_readout_bwd")
    return gw, gb, gwf, gbf, dh


def bond_update_bwd(dh_out, h_out, m, h0, plan: GraphPlan, weight, skip, act: int, z=None, dropout_p: float = 0.0,
                    seed: int = 0, layer: int = 0, training: bool = False, dh0_acc=None):
    """Backward of :func:`bond_update_fwd`: returns (gw, gb, gskip, dh_in, dh0_acc); ``dh0_acc`` accumulates
    ``skip * dz`` over the layers (pass the previous layer's result to add to it)."""
    lib = _lib.load()
    dh_out, h_out, m, h0, weight = (_f32(t) for t in (dh_out, h_out, m, h0, weight))
    z = None if z is None else _f32(z)
    skip = None if skip is None else _f32(skip)
    e, H = dh_out.shape
    f32 = dict(dtype=torch.float32, device=dh_out.device)
    gw, gb = torch.empty((H, H), **f32), torch.empty(H, **f32)
    gskip = torch.empty((), **f32) if skip is not None else None
    dh_in = torch.empty((e, H), **f32)
    first = dh0_acc is None
    if first:
        dh0_acc = torch.empty((e, H), **f32)
    ws, nbytes = _bwd_ws(lib, plan.n_atoms, e, 1, 0, H, dh_out.device)
    _lib.check(lib.cgr_bond_update_bwd(dh_out.data_ptr(), h_out.data_ptr(), _lib.ptr(z), m.data_ptr(), h0.data_ptr(),
                                       plan.in_ptr.data_ptr(), plan.in_idx.data_ptr(), plan.dst.data_ptr(),
                                       weight.data_ptr(), _lib.ptr(skip), act, float(dropout_p), seed, layer, int(training),
                                       gw.data_ptr(), gb.data_ptr(), _lib.ptr(gskip), dh_in.data_ptr(), dh0_acc.data_ptr(),
                                       int(first), e, plan.n_atoms, H, ws.data_ptr(), nbytes, _stream()),
               "cgr_bond_update_bwd")
    return gw, gb, gskip, dh_in, dh0_acc


def edge_init_bwd(dh0, h0, x, edge_attr, plan: GraphPlan, act: int, z0=None):
    """Backward of :func:`edge_init_fwd`: returns (gw_init [H, fa+fb], gb_init [H])."""
    lib = _lib.load()
    dh0, h0, x, edge_attr = (_f32(t) for t in (dh0, h0, x, edge_attr))
    z0 = None if z0 is None else _f32(z0)
    n, fa = x.shape
    e, fb = edge_attr.shape
    H = h0.shape[1]
    f32 = dict(dtype=torch.float32, device=x.device)
    gw, gb = torch.empty((H, fa + fb), **f32), torch.empty(H, **f32)
    ws, nbytes = _bwd_ws(lib, n, e, fa, fb, H, x.device)
    _lib.check(lib.cgr_edge_init_bwd(dh0.data_ptr(), h0.data_ptr(), _lib.ptr(z0), x.data_ptr(), edge_attr.data_ptr(),
                                     plan.in_ptr.data_ptr(), plan.in_idx.data_ptr(), act, gw.data_ptr(), gb.data_ptr(), n,
                                     e, fa, fb, H, ws.data_ptr(), nbytes, _stream()), "cgr_edge_init_bwd")
    return gw, gb
