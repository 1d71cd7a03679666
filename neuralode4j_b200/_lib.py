"""ctypes binding of libnode4j_b200.so (include/node4j.h).

This is the same C ABI a Java shim binds over Panama/JNI.  There is no CPU fallback: if the
library is missing it is built with nvcc, and if that fails or no CUDA device is present the compute entry
points raise.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

N4J_OK, ERR_INVALID_ARGUMENT, ERR_ILLEGAL_STATE, ERR_UNSUPPORTED, ERR_CUDA, ERR_MAX_STEPS, ERR_COMM = range(7)

LAYER_DENSE, LAYER_CONV3X3, LAYER_BATCHNORM, LAYER_ACTIVATION, LAYER_TIME_CONCAT = 1, 2, 3, 4, 5
ACT = {"identity": 0, "relu": 1, "elu": 2, "tanh": 3}
FLAG_EXACT_STAGE_TIMES, FLAG_NO_GRAPH, FLAG_FAST_TF32, FLAG_EXACT_CONTRACTIONS, FLAG_NO_TENSOR, FLAG_NO_PDL, FLAG_FORCE_TENSOR, FLAG_NO_PROLOGUE_FUSION, FLAG_NO_BWD_STATS_FUSION, FLAG_NO_DEFERRED_STATS, FLAG_NO_MLP_FUSION, FLAG_SIMT_WGRAD = 1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048
TIMEGRAD_NONE, TIMEGRAD_ZERO, TIMEGRAD_CALC = 0, 1, 2

EXPORTS = [
    "node4j_abi_version", "node4j_last_error", "node4j_device_count", "node4j_create", "node4j_destroy",
    "node4j_num_params", "node4j_num_real_params", "node4j_state_len", "node4j_bind_params", "node4j_forward",
    "node4j_backward", "node4j_get_step_log", "node4j_rhs", "node4j_rhs_vjp", "node4j_rhs_vjp_time", "node4j_error_norm",
    "node4j_comm_init", "node4j_comm_export", "node4j_bench_solver_pass", "node4j_bench_piece", "node4j_get_stream",
    "node4j_net_create", "node4j_net_destroy", "node4j_net_num_params", "node4j_net_ode_slice", "node4j_net_train_step", "node4j_net_output",
]
BENCH_RHS_FWD, BENCH_RHS_AUG, BENCH_CONV_FWD, BENCH_CONV_DGRAD, BENCH_CONV_WGRAD, BENCH_CONV_FWD_FUSED, BENCH_CONV_DGRAD_FUSED = 1, 2, 3, 4, 5, 6, 7


class LayerDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("activation", C.c_int32), ("n_in", C.c_int64), ("n_out", C.c_int64),
                ("has_bias", C.c_int32), ("eps", C.c_float)]


class GraphDesc(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("rank", C.c_int32), ("shape", C.c_int64 * 4), ("layers", C.POINTER(LayerDesc))]


class SolverCfg(C.Structure):
    _fields_ = [("abs_tol", C.c_double), ("rel_tol", C.c_double), ("min_step", C.c_double), ("max_step", C.c_double),
                ("safety", C.c_double), ("max_growth", C.c_double), ("min_reduction", C.c_double),
                ("order", C.c_int32), ("flags", C.c_uint32), ("max_trials", C.c_int64)]


class NetDesc(C.Structure):
    _fields_ = [("batch", C.c_int32), ("channels", C.c_int32), ("solver", SolverCfg), ("t0", C.c_float), ("t1", C.c_float)]


class Stats(C.Structure):
    _fields_ = [("nfe", C.c_int64), ("accepted", C.c_int64), ("rejected", C.c_int64), ("solves", C.c_int64),
                ("last_h", C.c_float), ("last_err", C.c_float), ("gpu_ms", C.c_float), ("gpu_launches", C.c_int32)]


class StepRec(C.Structure):
    _fields_ = [("t", C.c_float), ("h", C.c_float), ("err", C.c_float), ("accepted", C.c_int32), ("solve", C.c_int32)]


class Node4jError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


_EXC = {ERR_INVALID_ARGUMENT: ValueError,          # IllegalArgumentException
        ERR_ILLEGAL_STATE: RuntimeError,           # IllegalStateException
        ERR_UNSUPPORTED: NotImplementedError,      # UnsupportedOperationException
        }

_lib = None


def lib_path() -> str:
    return _build.LIB


def load():
    """Loads (building first if necessary) the native library.  Raises if it cannot be produced."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.build()
    L = C.CDLL(path)
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    L.node4j_abi_version.restype = i32
    L.node4j_last_error.restype = C.c_char_p
    L.node4j_device_count.restype = i32
    L.node4j_create.restype = i32
    L.node4j_create.argtypes = [C.POINTER(GraphDesc), C.POINTER(SolverCfg), i32, C.POINTER(vp)]
    L.node4j_destroy.restype = i32
    L.node4j_destroy.argtypes = [vp]
    for f in (L.node4j_num_params, L.node4j_num_real_params, L.node4j_state_len):
        f.restype = i64
        f.argtypes = [vp]
    L.node4j_bind_params.restype = i32
    L.node4j_bind_params.argtypes = [vp, vp, i64]
    L.node4j_forward.restype = i32
    L.node4j_forward.argtypes = [vp, vp, vp, i32, i32, vp, C.POINTER(Stats)]
    L.node4j_backward.restype = i32
    L.node4j_backward.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp, vp, C.POINTER(Stats)]
    L.node4j_get_step_log.restype = i64
    L.node4j_get_step_log.argtypes = [vp, C.POINTER(StepRec), i64]
    L.node4j_rhs.restype = i32
    L.node4j_rhs.argtypes = [vp, vp, f32, vp]
    L.node4j_rhs_vjp.restype = i32
    L.node4j_rhs_vjp.argtypes = [vp, vp, f32, vp, vp, vp, vp]
    L.node4j_rhs_vjp_time.restype = i32
    L.node4j_rhs_vjp_time.argtypes = [vp, vp, f32, vp, vp, vp, vp, C.POINTER(C.c_float)]
    L.node4j_error_norm.restype = i32
    L.node4j_error_norm.argtypes = [vp, vp, vp, vp, f32, i64, C.POINTER(f32)]
    L.node4j_comm_init.restype = i32
    L.node4j_comm_init.argtypes = [vp, i32, i32, vp, i64]
    L.node4j_comm_export.restype = i32
    L.node4j_comm_export.argtypes = [vp, vp]
    L.node4j_bench_solver_pass.restype = i32
    L.node4j_bench_solver_pass.argtypes = [vp, i64, i32, C.POINTER(f32)]
    L.node4j_bench_piece.restype = i32
    L.node4j_bench_piece.argtypes = [vp, i32, i32, C.POINTER(f32)]
    L.node4j_get_stream.restype = i32
    L.node4j_get_stream.argtypes = [vp, C.POINTER(vp)]
    L.node4j_net_create.restype = i32
    L.node4j_net_create.argtypes = [C.POINTER(NetDesc), i32, C.POINTER(vp)]
    L.node4j_net_destroy.restype = i32
    L.node4j_net_destroy.argtypes = [vp]
    L.node4j_net_num_params.restype = i64
    L.node4j_net_num_params.argtypes = [vp]
    L.node4j_net_ode_slice.restype = i32
    L.node4j_net_ode_slice.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
    L.node4j_net_train_step.restype = i32
    L.node4j_net_train_step.argtypes = [vp, vp, vp, vp, vp, f32, f32, C.POINTER(f32), vp, C.POINTER(Stats), C.POINTER(Stats)]
    L.node4j_net_output.restype = i32
    L.node4j_net_output.argtypes = [vp, vp, vp, vp, i32]
    if L.node4j_abi_version() != 2:
        raise RuntimeError("libnode4j_b200.so ABI version mismatch")
    _lib = L
    return L


def check(rc: int):
    if rc == N4J_OK:
        return
    msg = load().node4j_last_error().decode(errors="replace")
    exc = _EXC.get(rc)
    if exc is not None:
        raise exc(msg)
    raise Node4jError(rc, msg)
