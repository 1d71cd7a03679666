"""ctypes wrapper around oracle/libnode_oracle.so — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs import this
module.  It is the checker for the CUDA path, never the thing shipped or measured as the product.

The wrapper speaks the *reference's* tensor conventions so that tests read like the reference's tests:
  * single time pair: state in -> state out (SingleStep.java:32-46)
  * multi time: output ``[B,H,T]`` (rank-2 state) / ``[B,T,C,H,W]`` (rank-4 state), z(t0) included
    (MultiStep.alignOutShape, MultiStep.java:49-60); backward takes the same layouts
    (MultiStepAdjoint.alignInShapeToTimeFirst, MultiStepAdjoint.java:112-123).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libnode_oracle.so")

KIND = {"dense": 1, "conv3x3": 2, "batchnorm": 3, "activation": 4, "time_concat": 5}
ACT = {"identity": 0, "relu": 1, "elu": 2, "tanh": 3}
TG_NONE, TG_ZERO, TG_CALC = 0, 1, 2


class LayerDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("activation", C.c_int32), ("n_in", C.c_int64), ("n_out", C.c_int64),
                ("has_bias", C.c_int32), ("eps", C.c_float)]


class GraphDesc(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("rank", C.c_int32), ("shape", C.c_int64 * 4), ("layers", C.POINTER(LayerDesc))]


class SolverCfg(C.Structure):
    _fields_ = [("abs_tol", C.c_double), ("rel_tol", C.c_double), ("min_step", C.c_double), ("max_step", C.c_double),
                ("safety", C.c_double), ("max_growth", C.c_double), ("min_reduction", C.c_double),
                ("order", C.c_int32), ("exact_stage_times", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("nfe", C.c_int64), ("accepted", C.c_int64), ("rejected", C.c_int64), ("solves", C.c_int64)]


class StepRec(C.Structure):
    _fields_ = [("t", C.c_double), ("h", C.c_double), ("err", C.c_double), ("accepted", C.c_int32), ("solve", C.c_int32)]


def build(force: bool = False) -> str:
    """Compile the oracle with oracle/Makefile (g++).  Building the checker is not using it."""
    src = os.path.join(_HERE, "node_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.orc_last_error.restype = C.c_char_p
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(GraphDesc), C.POINTER(SolverCfg), C.c_int]
        L.orc_destroy.argtypes = [C.c_void_p]
        for f in (L.orc_num_params, L.orc_num_real_params, L.orc_state_len):
            f.restype = C.c_int64
            f.argtypes = [C.c_void_p]
        L.orc_last_dt.restype = C.c_double
        L.orc_last_dt.argtypes = [C.c_void_p]
        L.orc_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(Stats)]
        L.orc_backward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Stats)]
        L.orc_get_step_log.restype = C.c_int64
        L.orc_get_step_log.argtypes = [C.c_void_p, C.POINTER(StepRec), C.c_int64]
        L.orc_rhs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
        L.orc_rhs_vjp.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_error_norm.restype = C.c_double
        L.orc_error_norm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int64, C.c_double, C.c_double, C.c_int]
        L.orc_next_step.restype = C.c_double
        L.orc_next_step.argtypes = [C.c_double, C.c_double, C.POINTER(SolverCfg), C.c_int]
        L.orc_init_step.restype = C.c_double
        L.orc_init_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
        L.orc_interpolate.argtypes = [C.c_void_p] * 5 + [C.c_double] * 4 + [C.c_void_p, C.c_int64, C.c_int]
        L.orc_tableau.argtypes = [C.c_void_p] * 5
        L.orc_num_threads.restype = C.c_int
        _lib = L
    return _lib


def solver_cfg(abs_tol=1e-3, rel_tol=1e-3, min_step=1e-10, max_step=100.0, exact_stage_times=False) -> SolverCfg:
    # controller constants: AdaptiveRungeKuttaStepPolicy.java:43-45, order 5: DormandPrince54Solver.java:100
    return SolverCfg(abs_tol, rel_tol, min_step, max_step, 0.9, 10.0, 0.2, 5, int(exact_stage_times))


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


@dataclass
class StepLog:
    t: np.ndarray
    h: np.ndarray
    err: np.ndarray
    accepted: np.ndarray
    solve: np.ndarray


class Oracle:
    """One OdeVertex worth of oracle: inner graph (chain of layers) + DP54 config.

    ``layers``: list of dicts ``{"kind": "dense"|"conv3x3"|"batchnorm"|"activation", "n_in":, "n_out":,
    "activation": "identity"|"relu"|"elu"|"tanh", "has_bias": bool, "eps": float}``.
    ``shape``: state shape ``[B,F]`` or ``[B,C,H,W]``.
    """

    def __init__(self, layers, shape, cfg: SolverCfg | None = None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        assert self.dtype in (np.dtype(np.float32), np.dtype(np.float64))
        self.shape = tuple(int(s) for s in shape)
        self.cfg = cfg or solver_cfg()
        self._layers = (LayerDesc * len(layers))()
        for i, l in enumerate(layers):
            self._layers[i] = LayerDesc(KIND[l["kind"]], ACT[l.get("activation", "identity")], int(l.get("n_in", 0)),
                                        int(l.get("n_out", 0)), int(bool(l.get("has_bias", l["kind"] == "dense"))),
                                        float(l.get("eps", 1e-5)))
        g = GraphDesc(len(layers), len(self.shape), (C.c_int64 * 4)(*(list(self.shape) + [0] * (4 - len(self.shape)))),
                      self._layers)
        self._h = lib().orc_create(C.byref(g), C.byref(self.cfg), int(self.dtype == np.float64))
        if not self._h:
            raise ValueError(lib().orc_last_error().decode())
        self.n_params = lib().orc_num_params(self._h)
        self.n_real = lib().orc_num_real_params(self._h)
        self.n_state = lib().orc_state_len(self._h)
        self.stats = Stats()

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_destroy(self._h)
            self._h = None

    def _a(self, x, shape=None):
        a = np.ascontiguousarray(np.asarray(x, dtype=self.dtype))
        if shape is not None:
            a = a.reshape(shape)
        return a

    # -- layout helpers (MultiStep.java:49-60 / MultiStepAdjoint.java:112-123) ------------------------
    def _to_ref_layout(self, tf: np.ndarray) -> np.ndarray:
        if len(self.shape) == 2:
            return np.ascontiguousarray(tf.transpose(1, 2, 0))          # [T,B,H] -> [B,H,T]
        return np.ascontiguousarray(tf.transpose(1, 0, 2, 3, 4))        # [T,B,C,H,W] -> [B,T,C,H,W]

    def _to_time_first(self, a: np.ndarray) -> np.ndarray:
        if len(self.shape) == 2:
            return np.ascontiguousarray(a.transpose(2, 0, 1))
        return np.ascontiguousarray(a.transpose(1, 0, 2, 3, 4))

    # -- API ---------------------------------------------------------------------------------------
    def forward(self, params, y0, t, interpolate=True):
        params = self._a(params, (self.n_params,))
        y0 = self._a(y0, self.shape)
        t = self._a(t).ravel()
        nt = t.size
        out = np.empty(((nt,) if nt > 2 else ()) + self.shape, dtype=self.dtype)
        rc = lib().orc_forward(self._h, _p(params), _p(y0), _p(t), nt, int(interpolate), _p(out), C.byref(self.stats))
        if rc:
            raise ValueError(lib().orc_last_error().decode())
        return self._to_ref_layout(out) if nt > 2 else out

    def backward(self, params, z_out, dL_dz, t, grad_inout=None, time_grad=TG_NONE):
        """Returns (dL/dy0, grad_flat[n_params], dL/dt or None)."""
        params = self._a(params, (self.n_params,))
        t = self._a(t).ravel()
        nt = t.size
        if nt > 2:
            zt = self._to_time_first(self._a(z_out))
            g = self._to_time_first(self._a(dL_dz))
        else:
            zt = self._a(z_out, self.shape)
            g = self._a(dL_dz, self.shape)
        grad = np.zeros(self.n_params, dtype=self.dtype) if grad_inout is None else self._a(grad_inout, (self.n_params,)).copy()
        dy0 = np.empty(self.shape, dtype=self.dtype)
        dt = np.zeros(nt, dtype=self.dtype) if time_grad != TG_NONE else None
        rc = lib().orc_backward(self._h, _p(params), _p(zt), _p(g), _p(t), nt, int(time_grad), _p(grad), _p(dy0),
                                _p(dt) if dt is not None else None, C.byref(self.stats))
        if rc:
            raise ValueError(lib().orc_last_error().decode())
        return dy0, grad, dt

    def step_log(self) -> StepLog:
        n = lib().orc_get_step_log(self._h, None, 0)
        buf = (StepRec * max(n, 1))()
        lib().orc_get_step_log(self._h, buf, n)
        return StepLog(np.array([buf[i].t for i in range(n)]), np.array([buf[i].h for i in range(n)]),
                       np.array([buf[i].err for i in range(n)]), np.array([buf[i].accepted for i in range(n)]),
                       np.array([buf[i].solve for i in range(n)]))

    def rhs(self, params, z, t=0.0):
        params = self._a(params, (self.n_params,))
        z = self._a(z, self.shape)
        out = np.empty_like(z)
        lib().orc_rhs(self._h, _p(params), _p(z), float(t), _p(out))
        return out

    def last_dt(self):
        """g^T df/dt of the last :meth:`rhs_vjp` (the epsilon of the time input, TimeInput.java:46-50); 0 for autonomous graphs."""
        return float(lib().orc_last_dt(self._h))

    def rhs_vjp(self, params, z, g, t=0.0):
        """Returns (f(z), g^T df/dz, g^T df/dtheta_real[n_real])."""
        params = self._a(params, (self.n_params,))
        z = self._a(z, self.shape)
        g = self._a(g, self.shape)
        fz = np.empty_like(z)
        dz = np.empty_like(z)
        dreal = np.zeros(max(self.n_real, 1), dtype=self.dtype)
        lib().orc_rhs_vjp(self._h, _p(params), _p(z), float(t), _p(g), _p(fz), _p(dz), _p(dreal))
        return fz, dz, dreal[: self.n_real]

    def init_step(self, params, y, t0, t1):
        params = self._a(params, (self.n_params,))
        y = self._a(y, self.shape)
        k0 = np.empty_like(y)
        k1 = np.empty_like(y)
        h = lib().orc_init_step(self._h, _p(params), _p(y), float(t0), float(t1), _p(k0), _p(k1))
        return h, k0, k1


def error_norm(K, y0, y1, h, abs_tol, rel_tol, dtype=np.float32):
    K = np.ascontiguousarray(np.asarray(K, dtype=dtype))
    y0 = np.ascontiguousarray(np.asarray(y0, dtype=dtype)).ravel()
    y1 = np.ascontiguousarray(np.asarray(y1, dtype=dtype)).ravel()
    assert K.shape == (7, y0.size)
    return lib().orc_error_norm(_p(K), _p(y0), _p(y1), float(h), y0.size, float(abs_tol), float(rel_tol),
                                int(np.dtype(dtype) == np.float64))


def next_step(h, err, cfg: SolverCfg, dtype=np.float32):
    return lib().orc_next_step(float(h), float(err), C.byref(cfg), int(np.dtype(dtype) == np.float64))


def interpolate(y0, y1, ymid, f0, f1, dt, t0, t1, t, dtype=np.float32):
    arrs = [np.ascontiguousarray(np.asarray(a, dtype=dtype)) for a in (y0, y1, ymid, f0, f1)]
    out = np.empty_like(arrs[0])
    lib().orc_interpolate(*[_p(a) for a in arrs], float(dt), float(t0), float(t1), float(t), _p(out), out.size,
                          int(np.dtype(dtype) == np.float64))
    return out


def tableau():
    a = np.zeros((6, 6)); b = np.zeros(7); e = np.zeros(7); c = np.zeros(6); cm = np.zeros(7)
    lib().orc_tableau(_p(a), _p(b), _p(e), _p(c), _p(cm))
    return a, b, e, c, cm


def num_threads() -> int:
    return lib().orc_num_threads()


def set_num_threads(n: int) -> None:
    """Overrides OMP_NUM_THREADS (torchrun exports 1 for its children) for the calls that follow."""
    lib().orc_set_num_threads(int(n))
