"""The reference's MNIST ODE-net as a whole: host-side mirror of examples/mnist/OdeNetModel.java:60-88 —
ResBlockStem, OdeVertex, Output — whose `fit` / `output` run on the device through `node4j_net_*` (include/node4j.h).

    model = OdeNetModel(batch=128, nrofKernels=64, solver=DormandPrince54Solver(SolverConfig(1e-3, 1e-3, 1e-20, 100)))
    score = model.fit(images, labels_onehot, lr=0.1)       # ComputationGraph.fit(DataSet): forward, backward, Nesterovs update
    probs = model.output(images)                           # ComputationGraph.output

`params` / `updater_state` are the flat vectors DL4J keeps (and ModelSerializer stores: neuralode4j_b200/dl4j_zip.py).  There is no CPU
fallback: without a CUDA device `OdeNetModel(...)` raises."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .conf import DormandPrince54Solver, SolverConfig


class OdeNetModel:
    def __init__(self, batch: int = 128, nrofKernels: int = 64, solver: DormandPrince54Solver | None = None, time=(0.0, 1.0), device: int = 0,
                 flags: int = 0, momentum: float = 0.9):
        self._L = _lib.load()
        self._h = C.c_void_p()
        solver = solver or DormandPrince54Solver(SolverConfig(1e-3, 1e-3, 1e-20, 100))      # OdeNetModel.java:47
        c = solver.config
        cfg = _lib.SolverCfg(c.absTol, c.relTol, c.minStep, c.maxStep, 0.9, 10.0, 0.2, 5, int(flags), 0)
        desc = _lib.NetDesc(int(batch), int(nrofKernels), cfg, float(time[0]), float(time[1]))
        _lib.check(self._L.node4j_net_create(C.byref(desc), int(device), C.byref(self._h)))
        self.batch, self.channels, self.momentum = int(batch), int(nrofKernels), float(momentum)
        n = int(self._L.node4j_net_num_params(self._h))
        off, ln = C.c_int64(), C.c_int64()
        _lib.check(self._L.node4j_net_ode_slice(self._h, C.byref(off), C.byref(ln)))
        self.ode_slice = (int(off.value), int(ln.value))
        self.params = np.zeros(n, dtype=np.float32)
        self.updater_state = np.zeros(n, dtype=np.float32)
        self.ode_fwd_stats, self.ode_bwd_stats = _lib.Stats(), _lib.Stats()
        self.last_gradient = None

    def numParams(self) -> int:
        return self.params.size

    def setParams(self, params) -> None:
        p = np.ascontiguousarray(np.asarray(params, dtype=np.float32).ravel())
        if p.size != self.params.size:
            raise ValueError(f"expected {self.params.size} parameters, got {p.size}")
        self.params = p.copy()

    def close(self) -> None:
        if self._h:
            self._L.node4j_net_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _p(self, a):
        return a.ctypes.data_as(C.c_void_p)

    def fit(self, images, labels, lr: float = 0.1, keep_gradient: bool = False) -> float:
        """One minibatch: ComputationGraph.fit.  Returns the score (LossMCXENT / B) of the minibatch before the update."""
        x = np.ascontiguousarray(np.asarray(images, dtype=np.float32).reshape(self.batch, 1, 28, 28))
        y = np.ascontiguousarray(np.asarray(labels, dtype=np.float32).reshape(self.batch, 10))
        score = C.c_float()
        g = np.zeros_like(self.params) if keep_gradient else None
        _lib.check(self._L.node4j_net_train_step(self._h, self._p(self.params), self._p(self.updater_state), self._p(x), self._p(y), float(lr),
                                                 self.momentum, C.byref(score), self._p(g) if g is not None else None,
                                                 C.byref(self.ode_fwd_stats), C.byref(self.ode_bwd_stats)))
        self.last_gradient = g
        return float(score.value)

    def output(self, images, train: bool = False):
        x = np.ascontiguousarray(np.asarray(images, dtype=np.float32).reshape(self.batch, 1, 28, 28))
        probs = np.empty((self.batch, 10), dtype=np.float32)
        _lib.check(self._L.node4j_net_output(self._h, self._p(self.params), self._p(x), self._p(probs), int(train)))
        return probs
