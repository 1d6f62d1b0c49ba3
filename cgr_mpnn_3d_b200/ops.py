"""torch custom ops ``cgr_b200::gnn_forward`` / ``cgr_b200::gnn_backward`` over the C ABI.

The ops take the reference model's parameter tensors in state_dict order
(``cgr_mpnn_3D/models/GNN.py:53-74``) and run the whole forward
(``GNN.forward``, ``GNN.py:76-110``) or its explicit backward in ``libcgr_b200.so``.
Autograd is registered explicitly (``register_autograd``); there is no composite fallback and no
CPU kernel: calling them with CPU tensors raises.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import torch
from torch import Tensor

from . import _lib

# order of the flat parameter list passed to the ops
#   [edge_init.weight, edge_init.bias, convs.0.lin.weight, convs.0.lin.bias, ..., edge_to_node.weight,
#    edge_to_node.bias, ffn.weight, ffn.bias, (skip_weights.0, ...)]


def _stream() -> int:
    return _lib.current_stream_handle()


def _unpack(params: Sequence[Tensor], depth: int, use_skip: bool):
    w_init, b_init = params[0], params[1]
    w_conv = [params[2 + 2 * l] for l in range(depth)]
    b_conv = [params[3 + 2 * l] for l in range(depth)]
    o = 2 + 2 * depth
    w_e2n, b_e2n, w_ffn, b_ffn = params[o], params[o + 1], params[o + 2], params[o + 3]
    skip = [params[o + 4 + l] for l in range(depth)] if use_skip else []
    return w_init, b_init, w_conv, b_conv, w_e2n, b_e2n, w_ffn, b_ffn, skip


class _Ctx:
    """Keeps ctypes arrays alive for the duration of one C call."""

    def __init__(self, params: Sequence[Tensor], depth: int, act: int, use_skip: bool, fa: int, fb: int,
                 dropout_ps: Sequence[float]):
        w_init, b_init, w_conv, b_conv, w_e2n, b_e2n, w_ffn, b_ffn, skip = _unpack(params, depth, use_skip)
        hidden = int(w_init.shape[0])
        self.hidden = hidden
        self._wc = _lib.ptr_array(w_conv)
        self._bc = _lib.ptr_array(b_conv)
        self._sk = _lib.ptr_array(skip) if use_skip else None
        self._dp = (C.c_float * depth)(*[float(p) for p in dropout_ps[:depth]])
        self.params = _lib.CgrParams(
            fa=fa, fb=fb, hidden=hidden, depth=depth, act=act, use_skip=int(use_skip),
            w_init=w_init.data_ptr(), b_init=b_init.data_ptr(),
            w_conv=C.cast(self._wc, _lib.c_void_pp), b_conv=C.cast(self._bc, _lib.c_void_pp),
            skip=C.cast(self._sk, _lib.c_void_pp) if use_skip else None,
            w_e2n=w_e2n.data_ptr(), b_e2n=b_e2n.data_ptr(), w_ffn=w_ffn.data_ptr(), b_ffn=b_ffn.data_ptr(),
            host_dropout_p=C.cast(self._dp, _lib.c_float_p),
        )


_CTX_CACHE: dict = {}

# Optional provider of the flat gradient buffer the backward writes into: callable (n_floats, device, params) -> 1-D fp32
# tensor or None (optim.PeerFusedAdam installs one that hands out its CUDA-IPC-shared arenas to ITS parameters' backward).
_GRAD_ARENA = None


def set_grad_arena(provider) -> None:
    global _GRAD_ARENA
    _GRAD_ARENA = provider


def _ctx_cached(role: str, params: Sequence[Tensor], depth: int, act: int, use_skip: bool, fa: int, fb: int,
                dropout_ps: Sequence[float]) -> "_Ctx":
    """ctypes parameter block keyed by the parameter storage (optimizers update in place, so the pointers of a
    training run never change); one block per role because forward and backward may run on different threads."""
    key = (role, depth, act, use_skip, fa, fb, tuple(float(p) for p in dropout_ps[:depth]),
           tuple(p.data_ptr() for p in params), int(params[0].shape[0]))
    ctx = _CTX_CACHE.get(key)
    if ctx is None:
        if len(_CTX_CACHE) >= 16:
            _CTX_CACHE.clear()
        ctx = _Ctx(params, depth, act, use_skip, fa, fb, dropout_ps)
        _CTX_CACHE[key] = ctx
    ctx.params.tc_weights = None
    ctx.params.tc_throughput = 0
    ctx.params.tc_fast = 0
    return ctx


def _graph_struct(x, edge_attr, src, dst, in_ptr, in_idx, atom_ptr, tile_info=None, n_tiles=0,
                  tc_status=None, x_hi=None, x_lo=None) -> _lib.CgrGraph:
    has_tiles = tile_info is not None and tile_info.numel() > 0 and n_tiles > 0
    has_split = x_hi is not None and x_lo is not None and x_hi.numel() > 0 and x_lo.numel() > 0
    return _lib.CgrGraph(
        n_atoms=int(x.shape[0]), n_bonds=int(src.shape[0]), n_rxn=int(atom_ptr.shape[0]) - 1,
        x=x.data_ptr(), edge_attr=edge_attr.data_ptr(), src=src.data_ptr(), dst=dst.data_ptr(),
        in_ptr=in_ptr.data_ptr(), in_idx=in_idx.data_ptr(), atom_ptr=atom_ptr.data_ptr(),
        tile_info=tile_info.data_ptr() if has_tiles else None, n_tiles=int(n_tiles) if has_tiles else 0,
        # the fp16-range flag word is observable on both tcgen05 paths (fused tile kernels and layer-wise GEMMs)
        tc_status=tc_status.data_ptr() if (tc_status is not None and tc_status.numel() > 0) else None,
        x_hi=x_hi.data_ptr() if has_split else None, x_lo=x_lo.data_ptr() if has_split else None,
    )


def _require_cuda(*tensors: Tensor) -> None:
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError("cgr_b200 ops need CUDA tensors: the CGR hot path has no CPU implementation")


def _f32c(t: Tensor) -> Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


GROUP_MAX = 24      # batches one cgr_gnn_forward_group call takes (tcf::MAX_GROUP)


def gnn_forward_group(graphs: Sequence[tuple], params: Sequence[Tensor], depth: int, act: int, use_skip: bool,
                      tc_weights: Tensor, fast: bool = False) -> List[Tensor]:
    """Inference forward of several independent batches in TWO launches (``cgr_gnn_forward_group``): one atom
    projection over every batch's atom tiles, one fused cluster kernel over every batch's tile groups.

    ``graphs``: per batch ``(x, edge_attr, plan, x_hi, x_lo)`` with a tile plan (``plan.tile_info``); returns one
    energy tensor per batch (views of one allocation).  tcgen05 engine only, no autograd."""
    lib = _lib.load()
    if not 1 <= len(graphs) <= GROUP_MAX:
        raise ValueError(f"a group holds 1..{GROUP_MAX} batches")
    x0 = graphs[0][0]
    _require_cuda(x0, *params)
    params = [_f32c(p) for p in params]
    fa, fb = int(x0.shape[1]), int(graphs[0][1].shape[1])
    ctx = _ctx_cached("fwd", params, depth, act, use_skip, fa, fb, [0.0] * depth)
    if tc_weights.numel() > 0:
        ctx.params.tc_weights = tc_weights.data_ptr()
    ctx.params.tc_throughput = 1
    ctx.params.tc_fast = 1 if fast else 0
    n = len(graphs)
    keep = []
    garr = (_lib.CgrGraph * n)()
    sizes = []
    for i, (x, edge_attr, plan, x_hi, x_lo) in enumerate(graphs):
        x, edge_attr = _f32c(x), _f32c(edge_attr)
        keep.append((x, edge_attr))
        if plan.tile_info is None or plan.n_tiles <= 0:
            raise RuntimeError("group forward needs tileable batches (reactions of <= 128 bonds)")
        garr[i] = _graph_struct(x, edge_attr, plan.src, plan.dst, plan.in_ptr, plan.in_idx, plan.atom_ptr,
                                plan.tile_info, plan.n_tiles, plan.tc_status, x_hi, x_lo)
        sizes.append(int(garr[i].n_rxn))
    out_all = torch.empty(sum(sizes), dtype=torch.float32, device=x0.device)
    outs = list(out_all.split(sizes))
    optr = (C.c_void_p * n)(*[o.data_ptr() for o in outs])
    with torch.cuda.device(x0.device):
        ws_bytes = int(lib.cgr_forward_group_workspace(C.byref(ctx.params), garr, n))
        if ws_bytes <= 0:
            raise RuntimeError("cgr_forward_group_workspace: unsupported group")
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x0.device)
        _lib.check(lib.cgr_gnn_forward_group(C.byref(ctx.params), garr, n, optr, ws.data_ptr(), ws_bytes, _stream()),
                   "cgr_gnn_forward_group")
    return outs


@torch.library.custom_op("cgr_b200::gnn_forward", mutates_args=())
def gnn_forward(x: Tensor, edge_attr: Tensor, src: Tensor, dst: Tensor, in_ptr: Tensor, in_idx: Tensor,
                atom_ptr: Tensor, params: Sequence[Tensor], depth: int, act: int, use_skip: bool,
                dropout_ps: Sequence[float], training: bool, seed: int, engine: int, tile_info: Tensor,
                n_tiles: int, tc_status: Tensor, tc_weights: Tensor, x_hi: Tensor, x_lo: Tensor,
                tc_throughput: bool, fused_train: bool) -> List[Tensor]:
    """Dispatcher-registered form of :func:`gnn_forward_impl` (same arguments, same results)."""
    return gnn_forward_impl(x, edge_attr, src, dst, in_ptr, in_idx, atom_ptr, params, depth, act, use_skip, dropout_ps,
                            training, seed, engine, tile_info, n_tiles, tc_status, tc_weights, x_hi, x_lo,
                            tc_throughput, fused_train)


def gnn_forward_impl(x: Tensor, edge_attr: Tensor, src: Tensor, dst: Tensor, in_ptr: Tensor, in_idx: Tensor,
                     atom_ptr: Tensor, params: Sequence[Tensor], depth: int, act: int, use_skip: bool,
                     dropout_ps: Sequence[float], training: bool, seed: int, engine: int, tile_info: Tensor,
                     n_tiles: int, tc_status: Tensor, tc_weights: Tensor, x_hi: Tensor, x_lo: Tensor,
                     tc_throughput: bool, fused_train: bool) -> List[Tensor]:
    """Returns ``[out, h_all, m_all, z_all, s, hv, zv, pooled, tc_blob]`` (saved tensors are empty in eval).
    ``fused_train``: training on the tcgen05 engine with the tile-local fused kernels -- everything the backward
    needs lives in ``tc_blob`` and the layer-wise buffers stay empty.

    ``tile_info`` / ``n_tiles`` / ``tc_status`` / ``tc_weights`` feed the tcgen05 engine (empty tensors
    when unused)."""
    _require_cuda(x, edge_attr, src, *params)
    lib = _lib.load()
    x, edge_attr = _f32c(x), _f32c(edge_attr)
    params = [_f32c(p) for p in params]
    fa, fb = int(x.shape[1]), int(edge_attr.shape[1])
    ctx = _ctx_cached("fwd", params, depth, act, use_skip, fa, fb, dropout_ps)
    if tc_weights.numel() > 0:
        ctx.params.tc_weights = tc_weights.data_ptr()
    ctx.params.tc_throughput = int(tc_throughput)
    if engine == _lib.ENGINE_TC_FAST:        # "fast" precision mode of the tcgen05 engine: inference only
        if training:
            raise RuntimeError("precision='fast' is an inference mode; training runs in the fp32-parity mode")
        ctx.params.tc_fast = 1
        engine = _lib.ENGINE_TC
    H = ctx.hidden
    g = _graph_struct(x, edge_attr, src, dst, in_ptr, in_idx, atom_ptr, tile_info, n_tiles, tc_status, x_hi, x_lo)
    n, e, b = g.n_atoms, g.n_bonds, g.n_rxn
    f32 = dict(dtype=torch.float32, device=x.device)
    out = torch.empty(b, **f32)
    tc_blob = torch.empty(0, dtype=torch.uint8, device=x.device)
    if training and fused_train:
        with torch.cuda.device(x.device):
            blob_bytes = int(lib.cgr_tc_saved_bytes(C.byref(ctx.params), C.byref(g)))
        if blob_bytes <= 0:
            raise RuntimeError("fused tcgen05 training is not available for this configuration")
        tc_blob = torch.empty(blob_bytes, dtype=torch.uint8, device=x.device)
        h_all, m_all, z_all, s, hv, zv, pooled = (torch.empty(0, **f32) for _ in range(7))
        saved = _lib.CgrSaved(tc_blob=tc_blob.data_ptr(), tc_blob_bytes=blob_bytes)
        saved_p = C.byref(saved)
    elif training:
        h_all = torch.empty((depth + 1, e, H), **f32)
        m_all = torch.empty((depth, e, H), **f32)
        need_z = act != 0
        z_all = torch.empty((depth + 1, e, H), **f32) if need_z else torch.empty(0, **f32)
        s = torch.empty((n, H), **f32)
        hv = torch.empty((n, H), **f32)
        zv = torch.empty((n, H), **f32) if need_z else torch.empty(0, **f32)
        pooled = torch.empty((b, H), **f32)
        saved = _lib.CgrSaved(h_all=h_all.data_ptr(), m_all=m_all.data_ptr(),
                              z_all=z_all.data_ptr() if need_z else None, s=s.data_ptr(), hv=hv.data_ptr(),
                              zv=zv.data_ptr() if need_z else None, pooled=pooled.data_ptr())
        saved_p = C.byref(saved)
    else:
        h_all, m_all, z_all, s, hv, zv, pooled = (torch.empty(0, **f32) for _ in range(7))
        saved_p = None
    with torch.cuda.device(x.device):
        ws_bytes = lib.cgr_forward_workspace(C.byref(ctx.params), C.byref(g), int(training), engine)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        _lib.check(lib.cgr_gnn_forward(C.byref(ctx.params), C.byref(g), out.data_ptr(), saved_p, int(training),
                                       seed & 0xFFFFFFFFFFFFFFFF, engine, ws.data_ptr(), ws_bytes, _stream()),
                   "cgr_gnn_forward")
    return [out, h_all, m_all, z_all, s, hv, zv, pooled, tc_blob]


@gnn_forward.register_fake
def _(x, edge_attr, src, dst, in_ptr, in_idx, atom_ptr, params, depth, act, use_skip, dropout_ps, training, seed,
      engine, tile_info, n_tiles, tc_status, tc_weights, x_hi, x_lo, tc_throughput, fused_train):
    H = params[0].shape[0]
    n, e, b = x.shape[0], src.shape[0], atom_ptr.shape[0] - 1
    mk = lambda *s: x.new_empty(s, dtype=torch.float32)
    blob = x.new_empty((0,), dtype=torch.uint8)
    if not training:
        return [mk(b)] + [mk(0) for _ in range(7)] + [blob]
    if fused_train:
        kp = (H + 63) // 64 * 64
        rows = int(n_tiles) * 128
        return [mk(b)] + [mk(0) for _ in range(7)] + [
            x.new_empty((2 * (depth + 1) * rows * kp * 2 + rows * H * 4 + n * H * 4 + 4096 * (2 * depth + 6),),
                        dtype=torch.uint8)]
    need_z = act != 0
    return [mk(b), mk(depth + 1, e, H), mk(depth, e, H), mk(depth + 1, e, H) if need_z else mk(0), mk(n, H),
            mk(n, H), mk(n, H) if need_z else mk(0), mk(b, H), blob]


@torch.library.custom_op("cgr_b200::gnn_backward", mutates_args=())
def gnn_backward(grad_out: Tensor, x: Tensor, edge_attr: Tensor, src: Tensor, dst: Tensor, in_ptr: Tensor,
                 in_idx: Tensor, atom_ptr: Tensor, params: Sequence[Tensor], saved: Sequence[Tensor], depth: int,
                 act: int, use_skip: bool, dropout_ps: Sequence[float], seed: int, engine: int, tc_weights: Tensor,
                 x_hi: Tensor, x_lo: Tensor, tile_info: Tensor, n_tiles: int, tc_status: Tensor) -> List[Tensor]:
    """Dispatcher-registered form of :func:`gnn_backward_impl`."""
    pg = gnn_backward_impl(grad_out, x, edge_attr, src, dst, in_ptr, in_idx, atom_ptr, params, saved, depth, act,
                           use_skip, dropout_ps, seed, engine, tc_weights, x_hi, x_lo, tile_info, n_tiles, tc_status)
    return [g.clone() for g in pg]       # registered ops may not return views of one buffer


def gnn_backward_impl(grad_out: Tensor, x: Tensor, edge_attr: Tensor, src: Tensor, dst: Tensor, in_ptr: Tensor,
                      in_idx: Tensor, atom_ptr: Tensor, params: Sequence[Tensor], saved: Sequence[Tensor], depth: int,
                      act: int, use_skip: bool, dropout_ps: Sequence[float], seed: int, engine: int,
                      tc_weights: Tensor, x_hi: Tensor, x_lo: Tensor, tile_info: Tensor, n_tiles: int,
                      tc_status: Tensor) -> List[Tensor]:
    """Explicit backward: gradients of every parameter, in parameter-list order."""
    _require_cuda(grad_out, x, *params)
    lib = _lib.load()
    x, edge_attr = _f32c(x), _f32c(edge_attr)
    params = [_f32c(p) for p in params]
    grad_out = _f32c(grad_out)
    fa, fb = int(x.shape[1]), int(edge_attr.shape[1])
    ctx = _ctx_cached("bwd", params, depth, act, use_skip, fa, fb, dropout_ps)      # resets tc_weights / tc_throughput / tc_fast
    if tc_weights.numel() > 0:
        ctx.params.tc_weights = tc_weights.data_ptr()
    h_all, m_all, z_all, s, hv, zv, pooled, tc_blob = saved
    need_z = act != 0
    if tc_blob.numel() > 0:          # fused tile-local backward of the tcgen05 engine
        g = _graph_struct(x, edge_attr, src, dst, in_ptr, in_idx, atom_ptr, tile_info, n_tiles, tc_status, x_hi, x_lo)
        sv = _lib.CgrSaved(tc_blob=tc_blob.data_ptr(), tc_blob_bytes=tc_blob.numel())
    else:
        g = _graph_struct(x, edge_attr, src, dst, in_ptr, in_idx, atom_ptr, None, 0, None, x_hi, x_lo)
        sv = _lib.CgrSaved(h_all=h_all.data_ptr(), m_all=m_all.data_ptr(),
                           z_all=z_all.data_ptr() if need_z else None, s=s.data_ptr(), hv=hv.data_ptr(),
                           zv=zv.data_ptr() if need_z else None, pooled=pooled.data_ptr())
    # all gradients are views of ONE flat fp32 buffer (16-byte aligned pieces, parameter order), so data-parallel
    # training all-reduces them with a single collective and no packing copies (parallel.allreduce_gradients_)
    sizes = [(p.numel() + 3) // 4 * 4 for p in params]
    flat = None
    if _GRAD_ARENA is not None:              # a data-parallel optimizer wants the gradients in memory its peers can map
        flat = _GRAD_ARENA(sum(sizes), x.device, params)
    if flat is None:
        flat = torch.empty(sum(sizes), dtype=torch.float32, device=x.device)
    grads = [(c if c.numel() == p.numel() else c[:p.numel()]).view(p.shape) for c, p in zip(flat.split(sizes), params)]
    gw_init, gb_init, gw_conv, gb_conv, gw_e2n, gb_e2n, gw_ffn, gb_ffn, gskip = _unpack(grads, depth, use_skip)
    wc, bc = _lib.ptr_array(gw_conv), _lib.ptr_array(gb_conv)
    sk = _lib.ptr_array(gskip) if use_skip else None
    gs = _lib.CgrGrads(w_init=gw_init.data_ptr(), b_init=gb_init.data_ptr(), w_conv=C.cast(wc, _lib.c_void_pp),
                       b_conv=C.cast(bc, _lib.c_void_pp), skip=C.cast(sk, _lib.c_void_pp) if use_skip else None,
                       w_e2n=gw_e2n.data_ptr(), b_e2n=gb_e2n.data_ptr(), w_ffn=gw_ffn.data_ptr(),
                       b_ffn=gb_ffn.data_ptr())
    with torch.cuda.device(x.device):
        ws_bytes = lib.cgr_backward_workspace(C.byref(ctx.params), C.byref(g), engine)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        _lib.check(lib.cgr_gnn_backward(C.byref(ctx.params), C.byref(g), C.byref(sv), grad_out.data_ptr(),
                                        C.byref(gs), seed & 0xFFFFFFFFFFFFFFFF, engine, ws.data_ptr(), ws_bytes,
                                        _stream()), "cgr_gnn_backward")
    return grads


@gnn_backward.register_fake
def _(grad_out, x, edge_attr, src, dst, in_ptr, in_idx, atom_ptr, params, saved, depth, act, use_skip, dropout_ps,
      seed, engine, tc_weights, x_hi, x_lo, tile_info, n_tiles, tc_status):
    return [torch.empty_like(p) for p in params]


def _setup_context(ctx, inputs, output):
    (x, edge_attr, src, dst, in_ptr, in_idx, atom_ptr, params, depth, act, use_skip, dropout_ps, training, seed,
     engine, _tile_info, _n_tiles, _tc_status, _tc_weights, _x_hi, _x_lo, _tp, _fused) = inputs
    ctx.cfg = (depth, act, use_skip, list(dropout_ps), training, seed, engine, int(_n_tiles))
    ctx.n_params = len(params)
    ctx.set_materialize_grads(False)      # no zero-filled gradients for the saved-activation outputs
    ctx.save_for_backward(x, edge_attr, src, dst, in_ptr, in_idx, atom_ptr, *params, *output[1:], _tc_weights, _x_hi,
                          _x_lo, _tile_info, _tc_status)


def _backward(ctx, grads):
    depth, act, use_skip, dropout_ps, training, seed, engine, n_tiles = ctx.cfg
    if not training:
        raise RuntimeError("cgr_b200::gnn_forward was run with training=False; no activations were saved "
                           "(call model.train() before a forward that needs gradients)")
    t = ctx.saved_tensors
    x, edge_attr, src, dst, in_ptr, in_idx, atom_ptr = t[:7]
    params = list(t[7:7 + ctx.n_params])
    saved = list(t[7 + ctx.n_params:-5])
    tc_weights, x_hi, x_lo, tile_info, tc_status = t[-5:]
    g_out = grads[0]
    if g_out is None:
        g_out = torch.zeros(atom_ptr.shape[0] - 1, dtype=torch.float32, device=x.device)
    pg = gnn_backward(g_out, x, edge_attr, src, dst, in_ptr, in_idx, atom_ptr, params, saved, depth, act, use_skip,
                      dropout_ps, seed, engine, tc_weights, x_hi, x_lo, tile_info, n_tiles, tc_status)
    return (None, None, None, None, None, None, None, pg, None, None, None, None, None, None, None, None, None,
            None, None, None, None, None, None)


gnn_forward.register_autograd(_backward, setup_context=_setup_context)


class GnnFunction(torch.autograd.Function):
    """Lean autograd node around the same two C calls (what ``GNN.forward`` uses when gradients are needed): the
    dispatcher-registered ops above cost several hundred microseconds of host time per call, which is more than
    the whole training step takes on the GPU.  ``call`` carries every non-differentiable argument."""

    @staticmethod
    def forward(ctx, call, x, edge_attr, *params):
        res = gnn_forward_impl(x, edge_attr, call["src"], call["dst"], call["in_ptr"], call["in_idx"], call["atom_ptr"],
                               list(params), call["depth"], call["act"], call["use_skip"], call["dropout_ps"], True,
                               call["seed"], call["engine"], call["tile_info"], call["n_tiles"], call["tc_status"],
                               call["tc_weights"], call["x_hi"], call["x_lo"], call["tc_throughput"],
                               call["fused_train"])
        ctx.call = call
        ctx.n_params = len(params)
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(x, edge_attr, *params, *res[1:])
        return res[0]

    @staticmethod
    def backward(ctx, g_out):
        call = ctx.call
        t = ctx.saved_tensors
        x, edge_attr = t[0], t[1]
        params = list(t[2:2 + ctx.n_params])
        saved = list(t[2 + ctx.n_params:])
        if g_out is None:
            g_out = torch.zeros(call["atom_ptr"].shape[0] - 1, dtype=torch.float32, device=x.device)
        pg = gnn_backward_impl(g_out, x, edge_attr, call["src"], call["dst"], call["in_ptr"], call["in_idx"],
                               call["atom_ptr"], params, saved, call["depth"], call["act"], call["use_skip"],
                               call["dropout_ps"], call["seed"], call["engine"], call["tc_weights"], call["x_hi"],
                               call["x_lo"], call["tile_info"], call["n_tiles"], call["tc_status"])
        return (None, None, None, *pg)


def prepare_tc_weights(params: Sequence[Tensor], depth: int, act: int, use_skip: bool, fa: int, fb: int) -> Tensor:
    """FP16 (hi, lo) operand form of the parameter tree for the tcgen05 engine (``cgr_tc_prepare_weights``)."""
    _require_cuda(*params)
    lib = _lib.load()
    params = [_f32c(p.detach()) for p in params]
    ctx = _Ctx(params, depth, act, use_skip, fa, fb, [0.0] * depth)
    dev = params[0].device
    with torch.cuda.device(dev):
        nbytes = lib.cgr_tc_weights_bytes(C.byref(ctx.params))
        buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _lib.check(lib.cgr_tc_prepare_weights(C.byref(ctx.params), buf.data_ptr(), nbytes, _stream()),
                   "cgr_tc_prepare_weights")
    return buf


def tc_linear(x: Tensor, w: Tensor, bias=None) -> Tensor:
    """Test entry: ``x @ w.T + bias`` on the TMA + tcgen05 FP16x3 pipeline (``cgr_tc_linear``)."""
    _require_cuda(x, w)
    lib = _lib.load()
    x, w = _f32c(x), _f32c(w)
    m, k = x.shape
    n = w.shape[0]
    out = torch.empty((m, n), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        nbytes = lib.cgr_tc_linear_workspace(m, n, k)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        b = None if bias is None else _f32c(bias)
        _lib.check(lib.cgr_tc_linear(x.data_ptr(), m, k, w.data_ptr(), n, _lib.ptr(b), out.data_ptr(), ws.data_ptr(),
                                     nbytes, _stream()), "cgr_tc_linear")
    return out


def tc_gemm_test(a: Tensor, b: Tensor, a_mn: bool, b_mn: bool) -> Tensor:
    """Test entry of the training GEMM: ``a`` is [m,k] (or [k,m] when ``a_mn``), ``b`` is [n,k] (or [k,n])."""
    _require_cuda(a, b)
    lib = _lib.load()
    a, b = _f32c(a), _f32c(b)
    k, m = (a.shape if a_mn else a.shape[::-1])
    n = b.shape[1] if b_mn else b.shape[0]
    out = torch.empty((m, n), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        nbytes = lib.cgr_tc_gemm_test_workspace(m, n, k)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=a.device)
        _lib.check(lib.cgr_tc_gemm_test(a.data_ptr(), b.data_ptr(), m, n, k, int(a_mn), int(b_mn), out.data_ptr(),
                                        ws.data_ptr(), nbytes, _stream()), "cgr_tc_gemm_test")
    return out
