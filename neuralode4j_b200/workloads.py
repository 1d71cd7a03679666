"""Synthetic workloads for the configurations named in BASELINE.json (shapes, seeds and distributions from
SURVEY.md section 8d).  Used by bench.py and the parity tests; no dataset files are needed."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

from .conf import (ActivationLayer, BatchNormalization, Convolution2D, DenseLayer, DormandPrince54Solver, FixedStep, InputStep,
                   MergeVertex, ShapeMatchVertex, SolverConfig)
from .vertex import OdeVertex


@dataclass
class Workload:
    name: str
    conf: OdeVertex
    shape: tuple
    params: np.ndarray
    y0: np.ndarray
    time: np.ndarray            # FixedStep: conf time; InputStep: the time input
    eps_seed: int
    eps_scale: float
    time_is_input: bool = False

    def epsilon(self, out_shape) -> np.ndarray:
        rng = np.random.default_rng(self.eps_seed)
        return (rng.standard_normal(out_shape) * self.eps_scale).astype(np.float32)


def _uniform(rng, n, fan_in):
    b = 1.0 / np.sqrt(fan_in)
    return rng.uniform(-b, b, size=n).astype(np.float32)


def mnist_block(batch: int = 128, channels: int = 64, hw: int = 7, abs_tol=1e-3, rel_tol=1e-3, seed: int = 666) -> Workload:
    """Config #2: OdeVertex(BN-ReLU, conv3x3, BN-ReLU, conv3x3, BN-ReLU), t=[0,1]
    (examples/mnist/OdeNetModel.java:72-88, solver cfg :47, batch MnistTrainDataProvider.java:26)."""
    solver = DormandPrince54Solver(SolverConfig(abs_tol, rel_tol, 1e-20, 100))
    conf = (OdeVertex.Builder(None, "normFirst", BatchNormalization(channels, "relu"))
            .addLayer("convFirst", Convolution2D(channels), "normFirst")
            .addLayer("normSecond", BatchNormalization(channels, "relu"), "convFirst")
            .addLayer("convSecond", Convolution2D(channels), "normSecond")
            .addLayer("normThird", BatchNormalization(channels, "relu"), "convSecond")
            .odeConf(FixedStep(solver, np.arange(2)))
            .build())
    rng = np.random.default_rng(seed)
    C = channels

    def bn():
        return np.concatenate([np.ones(C), np.zeros(C), np.zeros(C), np.ones(C)]).astype(np.float32)   # gamma, beta, mean, var
    params = np.concatenate([bn(), _uniform(rng, C * C * 9, C * 9), bn(), _uniform(rng, C * C * 9, C * 9), bn()])
    y0 = np.random.default_rng(1234).standard_normal((batch, C, hw, hw)).astype(np.float32)
    return Workload("mnist_odeblock", conf, (batch, C, hw, hw), params, y0, np.array([0, 1], dtype=np.float32), 4321, 1e-2)


def spiral_block(batch: int = 1000, latent: int = 4, hidden: int = 20, nt: int = 100, seed: int = 666) -> Workload:
    """Config #1: latent ODE block fc 4-20-20-4 (ELU, ELU, identity), first `nt` of linspace(0, 6*pi, 500), interpolated
    forward, no time gradient (examples/spiral/LatentOdeBlock.java:31-45, OdeNetModel.java:44-47, SpiralFactory.java:71)."""
    solver = DormandPrince54Solver(SolverConfig(1e-12, 1e-6, 1e-20, 1e2))
    conf = (OdeVertex.Builder(None, "fc1", DenseLayer(hidden, latent, "elu"))
            .addLayer("fc2", DenseLayer(hidden, hidden, "elu"), "fc1")
            .addLayer("fc3", DenseLayer(latent, hidden, "identity"), "fc2")
            .odeConf(InputStep(solver, 1, True, False))
            .build())
    rng = np.random.default_rng(seed)
    parts = []
    for nin, nout in ((latent, hidden), (hidden, hidden), (hidden, latent)):
        parts += [_uniform(rng, nin * nout, nin), _uniform(rng, nout, nin)]
    params = np.concatenate(parts)
    y0 = np.random.default_rng(1).standard_normal((batch, latent)).astype(np.float32)
    t = np.linspace(0, 6 * np.pi, 500)[:nt].astype(np.float32)
    return Workload("spiral_latent_ode", conf, (batch, latent), params, y0, t, 3, 1e-2, time_is_input=True)


def mlp_block(batch: int = 1024, dim: int = 4096, act: str = "tanh", abs_tol=1e-3, rel_tol=1e-3, seed: int = 2) -> Workload:
    """Config #3: synthetic dense-MLP OdeVertex, Dense(dim->dim)+act -> Dense(dim->dim), t=[0,1]."""
    solver = DormandPrince54Solver(SolverConfig(abs_tol, rel_tol, 1e-20, 100))
    conf = (OdeVertex.Builder(None, "fc1", DenseLayer(dim, dim, act))
            .addLayer("fc2", DenseLayer(dim, dim, "identity"), "fc1")
            .odeConf(FixedStep(solver, np.arange(2)))
            .build())
    rng = np.random.default_rng(seed)
    params = np.concatenate([_uniform(rng, dim * dim, dim), np.zeros(dim, np.float32), _uniform(rng, dim * dim, dim), np.zeros(dim, np.float32)])
    y0 = np.random.default_rng(1).standard_normal((batch, dim)).astype(np.float32)
    return Workload(f"mlp{dim}_odeblock", conf, (batch, dim), params, y0, np.array([0, 1], dtype=np.float32), 3, 1.0 / np.sqrt(dim))


def timeseries_block(batch: int = 1024, dim: int = 256, nt: int = 32, seed: int = 5) -> Workload:
    """Config #5 (timing variant): InputStep time series with needTimeGradient=true, single Dense(dim->dim) identity."""
    solver = DormandPrince54Solver(SolverConfig(1e-12, 1e-6, 1e-20, 1e2))
    conf = (OdeVertex.Builder(None, "0", DenseLayer(dim, dim, "identity"))
            .odeConf(InputStep(solver, 1, True, True)).build())
    rng = np.random.default_rng(seed)
    params = np.concatenate([_uniform(rng, dim * dim, dim), _uniform(rng, dim, dim)])
    y0 = np.random.default_rng(6).standard_normal((batch, dim)).astype(np.float32)
    t = np.linspace(0, 4, nt).astype(np.float32)
    return Workload("timeseries_inputstep", conf, (batch, dim), params, y0, t, 7, 1.0, time_is_input=True)


def anode_time_block(batch: int = 256, dim: int = 3, hidden: int = 32, time_input: bool = False, nt: int = 2, seed: int = 11,
                     abs_tol=1e-3, rel_tol=1e-3) -> Workload:
    """Time-dependent f, dense: the reference's ANODE block (examples/anode/OdeBlockWithTime.java:26-40 around MlpBlock.java:42-60):
    "timeConc" = ShapeMatchVertex(MergeVertex) with the time as last input, then Dense(hidden, ReLU) x2 and Dense(dim).
    `time_input=True` drives it through InputStep with needTimeGradient (the adjoint's time slot then integrates -a^T df/dt)."""
    solver = DormandPrince54Solver(SolverConfig(abs_tol, rel_tol, 1e-20, 100))
    b = (OdeVertex.Builder(None, "timeConc", ShapeMatchVertex(MergeVertex()), True)
         .addLayer("h1", DenseLayer(hidden, dim + 1, "relu"), "timeConc")
         .addLayer("h2", DenseLayer(hidden, hidden, "relu"), "h1")
         .addLayer("out", DenseLayer(dim, hidden, "identity"), "h2"))
    t = np.linspace(0, 1, nt).astype(np.float32)
    conf = b.odeConf(InputStep(solver, 1, True, True) if time_input else FixedStep(solver, t, True)).build()
    rng = np.random.default_rng(seed)
    params = np.concatenate([_uniform(rng, (dim + 1) * hidden, dim + 1), _uniform(rng, hidden, dim + 1),
                             _uniform(rng, hidden * hidden, hidden), _uniform(rng, hidden, hidden),
                             _uniform(rng, hidden * dim, hidden), _uniform(rng, dim, hidden)]) * np.float32(2.0)
    y0 = np.random.default_rng(seed + 1).standard_normal((batch, dim)).astype(np.float32)
    return Workload("anode_time_block", conf, (batch, dim), params, y0, t, seed + 2, 1.0, time_is_input=time_input)


def conv_time_block(batch: int = 8, channels: int = 16, hw: int = 6, time_channels: int = 10, seed: int = 21,
                    abs_tol=1e-3, rel_tol=1e-3) -> Workload:
    """Time-dependent f, convolutional: the vertex of examples/cifar10/OdeBlockTime.java:35-45 — ShapeMatchVertex(MergeVertex,
    {1: 10}) appends ten time channels, then a chain conv3x3 - BN/ReLU - conv3x3 and the trailing ReLU ActivationLayer (:45)."""
    solver = DormandPrince54Solver(SolverConfig(abs_tol, rel_tol, 1e-20, 100))
    C, k = channels, time_channels
    conf = (OdeVertex.Builder(None, "timeConc", ShapeMatchVertex(MergeVertex(), {1: k}), True)
            .addLayer("conv1", Convolution2D(C, C + k), "timeConc")
            .addLayer("bn1", BatchNormalization(C, "relu"), "conv1")
            .addLayer("conv2", Convolution2D(C, C), "bn1")
            .addLayer("relu", ActivationLayer("relu"), "conv2")
            .odeConf(FixedStep(solver, np.arange(2))).build())
    rng = np.random.default_rng(seed)
    bn = np.concatenate([np.ones(C), np.zeros(C), np.zeros(C), np.ones(C)]).astype(np.float32)
    params = np.concatenate([_uniform(rng, C * (C + k) * 9, (C + k) * 9), bn, _uniform(rng, C * C * 9, C * 9)])
    y0 = np.random.default_rng(seed + 1).standard_normal((batch, C, hw, hw)).astype(np.float32)
    return Workload("conv_time_block", conf, (batch, C, hw, hw), params, y0, np.array([0, 1], dtype=np.float32), seed + 2, 1e-1)


def golden_linear(sign: int = 1) -> Workload:
    """Config #5 (parity variant): the exact problem of MultiStepAdjointTest.java:174-327 — f(y) = y*A + b with
    A = 0.2*ones(2,2), b = 3, t = linspace(+-1, +-8, 10), y0 = linspace(-1.23, 2.34, 4).reshape(2,2),
    InputStep(DP54(1e-12, 1e-6, 1e-20, 1e2), 1, interpolate=true, needTimeGradient=true)."""
    solver = DormandPrince54Solver(SolverConfig(1e-12, 1e-6, 1e-20, 1e2))
    conf = (OdeVertex.Builder(None, "0", DenseLayer(2, 2, "identity"))
            .odeConf(InputStep(solver, 1, True, True)).build())
    params = np.array([0.2, 0.2, 0.2, 0.2, 3, 3], dtype=np.float32)
    y0 = np.linspace(-1.23, 2.34, 4).reshape(2, 2).astype(np.float32)
    t = np.linspace(sign * 1, sign * 8, 10).astype(np.float32)
    return Workload("golden_linear", conf, (2, 2), params, y0, t, 0, 1.0, time_is_input=True)


def golden_lossgrad(shape=(2, 2, 10)) -> np.ndarray:
    """lossgrad of MultiStepAdjointTest.java:205-208."""
    g = np.linspace(3, 13, int(np.prod(shape))).reshape(shape)
    for i in range(shape[0] // 2):
        g[i * 2, :, i * 2 + 1] *= -1
    return g.astype(np.float32)
