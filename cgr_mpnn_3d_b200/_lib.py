"""ctypes binding of ``libcgr_b200.so`` (the C ABI declared in ``include/cgr_b200.h``).

The library is loaded lazily at module scope so that ``torch.save(model)`` never has to pickle a
handle (reference ``cgr_mpnn_3D/training/trainer.py:208`` pickles the whole module).  There is no
fallback: if the shared library is missing or a call fails, a ``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcgr_b200.so")

ENGINE_SIMT = 0
ENGINE_TC = 1
ENGINE_TC_FAST = 2      # Python-level id: tcgen05 engine with cgr_params_t.tc_fast = 1 (single-pass fp16, inference only)
ACT_IDS = {"relu": 0, "silu": 1, "gelu": 2}

c_float_p = C.POINTER(C.c_float)
c_void_pp = C.POINTER(C.c_void_p)


class CgrParams(C.Structure):
    _fields_ = [
        ("fa", C.c_int32), ("fb", C.c_int32), ("hidden", C.c_int32), ("depth", C.c_int32),
        ("act", C.c_int32), ("use_skip", C.c_int32),
        ("w_init", C.c_void_p), ("b_init", C.c_void_p),
        ("w_conv", c_void_pp), ("b_conv", c_void_pp), ("skip", c_void_pp),
        ("w_e2n", C.c_void_p), ("b_e2n", C.c_void_p), ("w_ffn", C.c_void_p), ("b_ffn", C.c_void_p),
        ("host_dropout_p", c_float_p),
        ("tc_weights", C.c_void_p),
        ("tc_throughput", C.c_int32),
        ("tc_fast", C.c_int32),
    ]


class CgrGrads(C.Structure):
    _fields_ = [
        ("w_init", C.c_void_p), ("b_init", C.c_void_p),
        ("w_conv", c_void_pp), ("b_conv", c_void_pp), ("skip", c_void_pp),
        ("w_e2n", C.c_void_p), ("b_e2n", C.c_void_p), ("w_ffn", C.c_void_p), ("b_ffn", C.c_void_p),
    ]


class CgrGraph(C.Structure):
    _fields_ = [
        ("n_atoms", C.c_int64), ("n_bonds", C.c_int64), ("n_rxn", C.c_int64),
        ("x", C.c_void_p), ("edge_attr", C.c_void_p), ("src", C.c_void_p), ("dst", C.c_void_p),
        ("in_ptr", C.c_void_p), ("in_idx", C.c_void_p), ("atom_ptr", C.c_void_p),
        ("tile_info", C.c_void_p), ("n_tiles", C.c_int64), ("tc_status", C.c_void_p),
        ("x_hi", C.c_void_p), ("x_lo", C.c_void_p),
    ]


class CgrHostBatch(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("edge_attr", C.c_void_p), ("edge_index", C.c_void_p), ("ptr", C.c_void_p),
        ("batch", C.c_void_p), ("n_atoms", C.c_int64), ("n_bonds", C.c_int64), ("n_rxn", C.c_int64),
    ]


class CgrStore(C.Structure):
    _fields_ = [
        ("x_all", C.c_void_p), ("ea_all", C.c_void_p), ("ei_all", C.c_void_p), ("node_ptr", C.c_void_p),
        ("edge_ptr", C.c_void_p), ("node_ptr_host", C.c_void_p), ("edge_ptr_host", C.c_void_p),
        ("n_rxn", C.c_int64), ("e_all", C.c_int64), ("fa", C.c_int32), ("fb", C.c_int32),
    ]


class CgrAdamTensor(C.Structure):
    _fields_ = [
        ("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
        ("max_exp_avg_sq", C.c_void_p), ("numel", C.c_int64),
    ]


class CgrFeatureTables(C.Structure):
    _fields_ = [("symbol_z", C.c_int16 * 11), ("degrees", C.c_int16 * 6), ("charges", C.c_int16 * 5),
                ("num_hs", C.c_int16 * 5), ("hybridizations", C.c_int16 * 5)]


class CgrSaved(C.Structure):
    _fields_ = [
        ("h_all", C.c_void_p), ("m_all", C.c_void_p), ("z_all", C.c_void_p), ("s", C.c_void_p),
        ("hv", C.c_void_p), ("zv", C.c_void_p), ("pooled", C.c_void_p), ("tc_blob", C.c_void_p),
        ("tc_blob_bytes", C.c_size_t),
    ]


# name -> (restype, argtypes); must list every symbol include/cgr_b200.h declares
_V = C.c_void_p
_I64 = C.c_int64
_I32 = C.c_int32
_SZ = C.c_size_t
PROTOTYPES = {
    "cgr_version": (C.c_int, []),
    "cgr_last_error_string": (C.c_char_p, []),
    "cgr_collate_workspace": (_SZ, [_I64]),
    "cgr_collate_indices": (C.c_int, [_V, _V, _V, _I64, _I64, _I64, _V, _V, _V, _V, _V, _SZ, _V]),
    "cgr_csr_workspace": (_SZ, [_I64, _I64]),
    "cgr_csr_build": (C.c_int, [_V, _I64, _I64, _V, _V, _V, _V, _V, _V, _SZ, _V]),
    "cgr_csr_build_by_reaction": (C.c_int, [_V, _V, _V, _I64, _I64, _I64, _V, _V, _V, _V, _V, _V]),
    "cgr_infer_host_workspace": (C.c_int, [C.POINTER(CgrParams), _I64, _I64, _I64, C.POINTER(_SZ), C.POINTER(_SZ)]),
    "cgr_gnn_infer_host": (C.c_int, [C.POINTER(CgrParams), _V, _V, _V, _V, _V, _I64, _I64, _I64, _V, _V, _SZ, _V, _SZ, _V]),
    "cgr_gnn_infer_host_async": (C.c_int, [C.POINTER(CgrParams), _V, _V, _V, _V, _V, _I64, _I64, _I64, _V, _V, _SZ, _V, _SZ, _V]),
    "cgr_gnn_infer_host_multi_async": (C.c_int, [C.POINTER(CgrParams), _V, C.c_int32, _V, _V, _SZ, _V, _SZ, _V]),
    "cgr_tc_saved_bytes": (_SZ, [C.POINTER(CgrParams), C.POINTER(CgrGraph)]),
    "cgr_forward_group_workspace": (_SZ, [C.POINTER(CgrParams), C.POINTER(CgrGraph), C.c_int32]),
    "cgr_gnn_forward_group": (C.c_int, [C.POINTER(CgrParams), C.POINTER(CgrGraph), C.c_int32, _V, _V, _SZ, _V]),
    "cgr_tc_plan_host": (C.c_int, [_V, _V, _I64, _V, C.POINTER(C.c_int64)]),
    "cgr_store_pack_order": (C.c_int, [_V, _V, _I64, _V, _I64, _V]),
    "cgr_store_infer_workspace": (C.c_int, [C.POINTER(CgrParams), C.POINTER(CgrStore), _V, _I64, _I64, C.POINTER(_SZ),
                                            C.POINTER(_SZ)]),
    "cgr_store_infer": (C.c_int, [C.POINTER(CgrParams), C.POINTER(CgrStore), _V, _I64, _I64, _V, _V, _SZ, _V, _SZ,
                                  C.c_int32, _V]),
    "cgr_store_gather": (C.c_int, [_V, _V, _V, _V, _V, _V, _I64, _V, _V, _V, _I64, C.c_int32, C.c_int32, _I64, _V, _V,
                                   _V, _V, _V, _V]),
    "cgr_ipc_export": (C.c_int, [_V, _V, C.POINTER(C.c_int64)]),
    "cgr_ipc_open": (C.c_int, [_V, _I64, C.POINTER(C.c_void_p)]),
    "cgr_enable_peer_access": (C.c_int, [C.c_int32]),
    "cgr_peer_allreduce_adam": (C.c_int, [_V, C.c_int32, _V, _V, _V, _I64, C.c_int32, C.c_int32, C.c_int32, C.c_double,
                                          C.c_double, C.c_double, C.c_double, C.c_double, _I64, C.c_int32, C.c_float, _V]),
    "cgr_adam_step": (C.c_int, [_V, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, _I64,
                                C.c_int32, C.c_float, _V]),
    "cgr_infer_host_check": (C.c_int, [C.POINTER(CgrParams), _I64, _I64, _I64, _V]),
    "cgr_atom_ptr_from_batch": (C.c_int, [_V, _I64, _I64, _V, _V]),
    "cgr_edge_init_fwd": (C.c_int, [_V, _V, _V, _V, _V, _I64, _I64, _I32, _I32, _I32, _I32, _V, _V, _V, _SZ, _V]),
    "cgr_bond_update_fwd": (C.c_int, [_V, _V, _V, _V, _V, _V, _V, _V, _I32, C.c_float, C.c_uint64, C.c_uint32,
                                      _I32, _V, _V, _V, _I64, _I64, _I32, _V]),
    "cgr_conv_fwd": (C.c_int, [_V, _V, _V, _V, _V, _V, _V, _V, _V, _I64, _I64, _I32, _V]),
    "cgr_readout_fwd": (C.c_int, [_V, _V, _V, _V, _V, _V, _V, _V, _V, _I32, _V, _V, _V, _V, _V, _I64, _I64, _I64,
                                  _I32, _I32, _V]),
    "cgr_stage_bwd_workspace": (_SZ, [_I64, _I64, _I32, _I32, _I32]),
    "cgr_readout_bwd": (C.c_int, [_V, _V, _V, _V, _V, _V, _V, _V, _I32, _V, _V, _V, _V, _V, _V, _V, _V, _V, _I64, _I64,
                                  _I64, _I32, _I32, _V, _SZ, _V]),
    "cgr_bond_update_bwd": (C.c_int, [_V, _V, _V, _V, _V, _V, _V, _V, _V, _V, _I32, C.c_float, C.c_uint64, C.c_uint32,
                                      _I32, _V, _V, _V, _V, _V, _I32, _I64, _I64, _I32, _V, _SZ, _V]),
    "cgr_edge_init_bwd": (C.c_int, [_V, _V, _V, _V, _V, _V, _V, _I32, _V, _V, _I64, _I64, _I32, _I32, _I32, _V, _SZ, _V]),
    "cgr_forward_workspace": (_SZ, [C.POINTER(CgrParams), C.POINTER(CgrGraph), _I32, _I32]),
    "cgr_gnn_forward": (C.c_int, [C.POINTER(CgrParams), C.POINTER(CgrGraph), _V, C.POINTER(CgrSaved), _I32,
                                  C.c_uint64, _I32, _V, _SZ, _V]),
    "cgr_backward_workspace": (_SZ, [C.POINTER(CgrParams), C.POINTER(CgrGraph), _I32]),
    "cgr_gnn_backward": (C.c_int, [C.POINTER(CgrParams), C.POINTER(CgrGraph), C.POINTER(CgrSaved), _V,
                                   C.POINTER(CgrGrads), C.c_uint64, _I32, _V, _SZ, _V]),
    "cgr_tc_plan_build": (C.c_int, [_V, _V, _I64, _V, _V, _V]),
    "cgr_tc_plan_check": (C.c_int, [_V, _I64, _V, _V, _V, _V]),
    "cgr_tc_gemm_test_workspace": (_SZ, [_I64, _I64, _I64]),
    "cgr_tc_gemm_test": (C.c_int, [_V, _V, _I64, _I64, _I64, _I32, _I32, _V, _V, _SZ, _V]),
    "cgr_tc_features_ld": (_I64, [_I32]),
    "cgr_tc_split_features": (C.c_int, [_V, _I64, _I32, _V, _V, _V, _V]),
    "cgr_tc_debug_buffer": (C.c_int, [_V]),
    "cgr_tc_weights_bytes": (_SZ, [C.POINTER(CgrParams)]),
    "cgr_tc_prepare_weights": (C.c_int, [C.POINTER(CgrParams), _V, _SZ, _V]),
    "cgr_tc_linear_workspace": (_SZ, [_I64, _I64, _I64]),
    "cgr_tc_linear": (C.c_int, [_V, _I64, _I64, _V, _I64, _V, _V, _V, _SZ, _V]),
    "cgr_mse_sum_fwd_bwd": (C.c_int, [_V, _V, _I64, _V, _V, _V]),
    "cgr_featurize_cgr": (C.c_int, [C.POINTER(CgrFeatureTables), _V, _V, _V, _V, _I64, _V, _V, _I64, _V, _I64, _V, _V]),
    "cgr_launch_count": (C.c_longlong, []),
    "cgr_profile_enable": (C.c_int, [C.c_int]),
    "cgr_profile_count": (C.c_int, []),
    "cgr_profile_get": (C.c_int, [C.c_int, C.c_char_p, C.c_int, c_float_p]),
    "cgr_dropout_mask": (C.c_int, [C.c_uint64, C.c_uint32, C.c_float, _I64, _I32, _V, _V]),
}

_lib = None
_lock = threading.Lock()


def load() -> C.CDLL:
    """Load the shared library (once) and attach prototypes.  Raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m cgr_mpnn_3d_b200.build` "
                "(nvcc, sm_100a).  There is no CPU or PyTorch fallback for the CGR hot path.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)      # AttributeError if a declared symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().cgr_last_error_string()
        raise RuntimeError(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")


def ptr(t) -> int:
    """Device (or host) address of a torch tensor, or 0 for None."""
    return 0 if t is None else t.data_ptr()


def ptr_array(tensors):
    arr = (C.c_void_p * max(1, len(tensors)))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


_raw_stream = None


def current_stream_handle() -> int:
    """Raw ``cudaStream_t`` of the current device's current stream.  ``torch.cuda.current_stream()`` builds a Stream
    object and resolves the device index through several Python layers (~20 us per call, several calls per step)."""
    global _raw_stream
    import torch
    if _raw_stream is None:
        _raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", False)
    if _raw_stream:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream
