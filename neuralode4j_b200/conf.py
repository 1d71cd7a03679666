"""Configuration classes with the reference's names, constructor arguments and error behaviour.

Mirrors (all under /root/reference/src/main/java):
  ode/solve/conf/SolverConfig.java:17-30            -> SolverConfig
  ode/solve/conf/DormandPrince54Solver.java:24-57   -> DormandPrince54Solver
  ode/vertex/conf/helper/FixedStep.java:21-39       -> FixedStep
  ode/vertex/conf/helper/InputStep.java:29-52       -> InputStep
  util/listen/step/StepCounter.java:46-63, Mask.java:24-47 -> StepCounter, Mask
DL4J layer confs used inside an OdeVertex (third party in the reference):
  DenseLayer, Convolution2D (3x3, ConvolutionMode.Same, hasBias(false)), BatchNormalization, ActivationLayer

These objects hold no arithmetic; `OdeVertex.instantiate` hands them to the native engine through the C ABI.
"""
from __future__ import annotations

import copy
import json
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _lib


# ----------------------------------------------------------------------------------------------------------
# activation types (DL4J's InputType, the part OdeVertex.getOutputType needs)
# ----------------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class InputType:
    """org.deeplearning4j.nn.conf.inputs.InputType restricted to what the ODE block's output-type inference touches
    (ode/vertex/conf/helper/forward/OutputTypeFromConfig.java:27-35, OutputTypeAddTimeAsDimension.java:47-53).
    ``shape`` is per example: FF (size,), RNN (size, timeSeriesLength), CNN (channels, height, width),
    CNN3D in NDHWC (depth, height, width, channels)."""
    kind: str
    shape: tuple

    @staticmethod
    def feedForward(size): return InputType("FF", (int(size),))
    @staticmethod
    def recurrent(size, timeSeriesLength): return InputType("RNN", (int(size), int(timeSeriesLength)))
    @staticmethod
    def convolutional(height, width, channels): return InputType("CNN", (int(channels), int(height), int(width)))
    @staticmethod
    def convolutional3D(depth, height, width, channels): return InputType("CNN3D", (int(depth), int(height), int(width), int(channels)))
    @staticmethod
    def inferInputType(array): return InputType.feedForward(np.asarray(array).size)     # a time vector [1,T] is FF(T)

    def arrayElementsPerExample(self) -> int:
        return int(np.prod(self.shape))


def _output_type_from_config(vertex, inputs):
    """OutputTypeFromConfig.java:27-35: the inner graph's single output must have the input's shape."""
    if len(inputs) != 1:
        raise ValueError("Can only support one single output!")
    out = vertex.activationType(inputs[0])
    if out != inputs[0]:
        raise ValueError(f"input shape must match output shape! Got input {inputs[0]} and output: {out}!")
    return out


def _add_time_dim(out: InputType, time: InputType) -> InputType:
    """OutputTypeAddTimeAsDimension.java:25-53: FF -> recurrent(size, T); CNN -> convolutional3D(NDHWC, T, h, w, c)."""
    if time.kind != "FF":
        raise ValueError("Time must be 1D!")
    t = time.arrayElementsPerExample()
    if out.kind == "FF":
        return InputType.recurrent(out.arrayElementsPerExample(), t)
    if out.kind == "CNN":
        c, h, w = out.shape
        return InputType.convolutional3D(t, h, w, c)
    raise ValueError("Input type not supported with time as input!")


# ----------------------------------------------------------------------------------------------------------
# solver configuration
# ----------------------------------------------------------------------------------------------------------
class SolverConfig:
    """ode.solve.conf.SolverConfig(absoluteTolerance, relativeTolerance, minStep, maxStep)."""

    def __init__(self, absoluteTolerance: float, relativeTolerance: float, minStep: float, maxStep: float):
        if minStep >= maxStep:   # SolverConfig.java:22-24
            raise ValueError(f"Max step smaller than min step! Swapped arguments? max: {maxStep} min {minStep}")
        self.absTol = float(absoluteTolerance)
        self.relTol = float(relativeTolerance)
        self.minStep = float(minStep)
        self.maxStep = float(maxStep)

    def getAbsTol(self): return self.absTol
    def getRelTol(self): return self.relTol
    def getMinStep(self): return self.minStep
    def getMaxStep(self): return self.maxStep

    def __eq__(self, o):
        return isinstance(o, SolverConfig) and (self.absTol, self.relTol, self.minStep, self.maxStep) == (o.absTol, o.relTol, o.minStep, o.maxStep)

    def __hash__(self):
        return hash((self.absTol, self.relTol, self.minStep, self.maxStep))

    def to_json(self):
        return {"absoluteTolerance": self.absTol, "relativeTolerance": self.relTol, "minStep": self.minStep, "maxStep": self.maxStep}

    @staticmethod
    def from_json(d):
        return SolverConfig(d["absoluteTolerance"], d["relativeTolerance"], d["minStep"], d["maxStep"])


class StepListener:
    """ode.solve.api.StepListener: begin(t, y0) / step(solverState, step, error) / done().

    The native solve runs on the device without host callbacks; listeners are replayed from the device step log
    after each integrate call, once per ACCEPTED step (rejected steps are invisible, as in the reference).  The
    replayed ``solverState`` exposes ``time()`` and ``getInterpolationMidpoints()``; the state accessors raise.
    """

    def begin(self, t, y0): ...
    def step(self, solverState, step, error): ...
    def done(self): ...


class _ReplayState:
    """ode.solve.impl.util.SolverState as the replayed listeners see it (the Java shim's ReplayedSolverState): time and the solver's
    dense-output midpoint coefficients are available; the state vectors stay on the device, so their accessors raise (the reference's
    listeners that only look at time / step / error — StepCounter, Mask, the plotting listeners — work unchanged)."""

    # DormandPrince54Solver.java:44-47 (cMid); the device-side dense output uses the same coefficients
    _C_MID = (6025192743.0 / 30085553152.0 / 2.0, 0.0, 51252292925.0 / 65400821598.0 / 2.0, -2691868925.0 / 45128329728.0 / 2.0,
              187940372067.0 / 1594534317056.0 / 2.0, -1776094331.0 / 19743644256.0 / 2.0, 11237099.0 / 235043384.0 / 2.0)

    def __init__(self, t):
        self._t = t

    def time(self):
        return self._t

    def getInterpolationMidpoints(self):
        return list(self._C_MID)

    def getCurrentState(self):
        raise NotImplementedError("native ODE helper: the solver state stays on the device (UnsupportedOperationException in the Java shim)")

    def getStateDot(self, stage):
        raise NotImplementedError("native ODE helper: stage derivatives stay on the device (UnsupportedOperationException in the Java shim)")


class DormandPrince54Solver:
    """ode.solve.conf.DormandPrince54Solver: serialisable configuration of the DP5(4) solver."""

    JCLASS = "ode.solve.conf.DormandPrince54Solver"

    def __init__(self, config: Optional[SolverConfig] = None):
        self.config = config if config is not None else SolverConfig(1e-3, 1e-3, 1e-10, 100)   # :31-33
        self.listeners: List[StepListener] = []

    def getConfig(self):
        return self.config

    def addListeners(self, *listeners):
        for l in listeners:
            if l not in self.listeners:
                self.listeners.append(l)

    def clearListeners(self, *listeners):
        if not listeners:
            self.listeners.clear()
        else:
            self.listeners = [l for l in self.listeners if l not in listeners]

    def clone(self):   # :46-52 (listeners are not cloned)
        c = self.config
        return DormandPrince54Solver(SolverConfig(c.absTol, c.relTol, c.minStep, c.maxStep))

    def __eq__(self, o):
        return isinstance(o, DormandPrince54Solver) and self.config == o.config

    def __hash__(self):
        return hash(self.config)

    def to_json(self):
        return {"@class": self.JCLASS, "config": self.config.to_json()}

    @staticmethod
    def from_json(d):
        if d.get("@class") != DormandPrince54Solver.JCLASS:
            raise NotImplementedError(f"solver conf {d.get('@class')} has no native implementation")
        return DormandPrince54Solver(SolverConfig.from_json(d["config"]))

    def native_cfg(self, flags: int = 0, max_trials: int = 0) -> _lib.SolverCfg:
        c = self.config
        # controller constants hard-coded in the reference: AdaptiveRungeKuttaStepPolicy.java:43-45, order 5
        return _lib.SolverCfg(c.absTol, c.relTol, c.minStep, c.maxStep, 0.9, 10.0, 0.2, 5, flags, max_trials)


# ----------------------------------------------------------------------------------------------------------
# step listeners shipped with the reference
# ----------------------------------------------------------------------------------------------------------
class StepCounter(StepListener):
    """util.listen.step.StepCounter(nrToAccum, msg): average number of accepted steps per solve, reported every
    nrToAccum solves (StepCounter.java:46-63)."""

    def __init__(self, nrToAccum: int = 20, msg: str = "average number of solver steps: ", consumer=None):
        self.nrToAccum = nrToAccum
        self.msg = msg
        self.nrofSteps = 0
        self.nrofSolves = 0
        self.consumer = consumer if consumer is not None else self._default
        self.last_average = None

    def _default(self, nrofSteps, nrofSolves):
        self.last_average = nrofSteps / nrofSolves

    def begin(self, t, y0):
        self.nrofSolves += 1

    def step(self, solverState, step, error):
        self.nrofSteps += 1

    def done(self):
        if self.nrofSolves == self.nrToAccum:
            self.consumer(self.nrofSteps, self.nrofSolves)
            self.nrofSolves = 0
            self.nrofSteps = 0


class Mask(StepListener):
    """util.listen.step.Mask: forwards callbacks only for solves in one time direction."""

    def __init__(self, listener: StepListener, want_forward: bool):
        self.listener = listener
        self.want_forward = want_forward
        self.active = False

    @staticmethod
    def forward(listener):    # masks FORWARD solves, i.e. lets backward-in-time solves through (Mask.java:24-30)
        return Mask(listener, want_forward=False)

    @staticmethod
    def backward(listener):   # masks BACKWARD solves
        return Mask(listener, want_forward=True)

    def begin(self, t, y0):
        t = np.asarray(t).ravel()
        is_forward = t[-1] > t[0]
        self.active = (is_forward == self.want_forward)
        if self.active:
            self.listener.begin(t, y0)

    def step(self, solverState, step, error):
        if self.active:
            self.listener.step(solverState, step, error)

    def done(self):
        if self.active:
            self.listener.done()


# ----------------------------------------------------------------------------------------------------------
# ODE helper confs
# ----------------------------------------------------------------------------------------------------------
def _time_array(time) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(time, dtype=np.float32).ravel())


class FixedStep:
    """ode.vertex.conf.helper.FixedStep(solver, time[, interpolateForwardIfMultiStep])."""

    JCLASS = "ode.vertex.conf.helper.FixedStep"

    def __init__(self, solver: DormandPrince54Solver, time, interpolateForwardIfMultiStep: bool = False):
        self.solver = solver
        self.time = _time_array(time).copy()      # time.dup() :27
        self.interpolateForwardIfMultiStep = bool(interpolateForwardIfMultiStep)
        if self.time.size < 2:
            raise ValueError(f"time must be a vector with two elements! Was of shape: {list(self.time.shape)}!")

    # OdeHelper.forward()/backward(): the two halves are the same native handle here
    def forward(self): return self
    def backward(self): return self

    time_input_index = None
    needTimeGradient = False

    def nrofTimeInputs(self) -> int:          # conf/helper/forward/FixedStep.java:45-47
        return 0

    def getOutputType(self, vertex, *vertexInputs):
        """conf/helper/forward/FixedStep.java:55-67: two time points keep the input's type, more add the time dimension."""
        if self.time.size == 2:
            return _output_type_from_config(vertex, vertexInputs)
        return _add_time_dim(_output_type_from_config(vertex, vertexInputs), InputType.inferInputType(self.time))

    def clone(self):
        return FixedStep(self.solver.clone(), self.time, self.interpolateForwardIfMultiStep)

    def __eq__(self, o):
        return (isinstance(o, FixedStep) and self.solver == o.solver and np.array_equal(self.time, o.time)
                and self.interpolateForwardIfMultiStep == o.interpolateForwardIfMultiStep)

    def to_json(self):
        return {"@class": self.JCLASS, "solver": self.solver.to_json(), "time": self.time.tolist(),
                "interpolateForwardIfMultiStep": self.interpolateForwardIfMultiStep}

    @staticmethod
    def from_json(d):
        return FixedStep(DormandPrince54Solver.from_json(d["solver"]), d["time"], d.get("interpolateForwardIfMultiStep", False))


class InputStep:
    """ode.vertex.conf.helper.InputStep(solverConf, timeInputIndex[, interpolateForwardIfMultiStep[, needTimeGradient]])."""

    JCLASS = "ode.vertex.conf.helper.InputStep"

    def __init__(self, solverConf: DormandPrince54Solver, timeInputIndex: int, interpolateForwardIfMultiStep: bool = False,
                 needTimeGradient: bool = False):
        self.solver = solverConf
        self.time_input_index = int(timeInputIndex)
        self.interpolateForwardIfMultiStep = bool(interpolateForwardIfMultiStep)
        self.needTimeGradient = bool(needTimeGradient)
        self.time = None

    def forward(self): return self
    def backward(self): return self

    def nrofTimeInputs(self) -> int:          # conf/helper/forward/InputStep.java: one of the vertex inputs is the time
        return 1

    def getOutputType(self, vertex, *vertexInputs):
        """conf/helper/forward/InputStep.java:50-66."""
        if len(vertexInputs) <= self.time_input_index:
            raise ValueError("Time input index was not part of input types!!")
        time = vertexInputs[self.time_input_index]
        rest = [v for i, v in enumerate(vertexInputs) if i != self.time_input_index]
        out = _output_type_from_config(vertex, rest)
        return _add_time_dim(out, time) if time.arrayElementsPerExample() > 2 else out

    def clone(self):
        return InputStep(self.solver.clone(), self.time_input_index, self.interpolateForwardIfMultiStep, self.needTimeGradient)

    def __eq__(self, o):
        return (isinstance(o, InputStep) and self.solver == o.solver and self.time_input_index == o.time_input_index
                and self.interpolateForwardIfMultiStep == o.interpolateForwardIfMultiStep
                and self.needTimeGradient == o.needTimeGradient)

    def to_json(self):
        return {"@class": self.JCLASS, "solverConf": self.solver.to_json(), "timeInputIndex": self.time_input_index,
                "interpolateForwardIfMultiStep": self.interpolateForwardIfMultiStep, "needTimeGradient": self.needTimeGradient}

    @staticmethod
    def from_json(d):
        return InputStep(DormandPrince54Solver.from_json(d["solverConf"]), d["timeInputIndex"],
                         d.get("interpolateForwardIfMultiStep", False), d.get("needTimeGradient", False))


def ode_helper_from_json(d):
    cls = {FixedStep.JCLASS: FixedStep, InputStep.JCLASS: InputStep}.get(d.get("@class"))
    if cls is None:
        raise NotImplementedError(f"ODE helper {d.get('@class')} has no native implementation")
    return cls.from_json(d)


# ----------------------------------------------------------------------------------------------------------
# layers allowed inside a native OdeVertex
# ----------------------------------------------------------------------------------------------------------
@dataclass
class Layer:
    kind: int = 0
    nIn: int = 0
    nOut: int = 0
    activation: str = "identity"
    hasBias: bool = False
    eps: float = 1e-5

    def desc(self) -> _lib.LayerDesc:
        if self.activation not in _lib.ACT:
            raise NotImplementedError(f"activation {self.activation} is outside the native set {sorted(_lib.ACT)}")
        return _lib.LayerDesc(self.kind, _lib.ACT[self.activation], int(self.nIn), int(self.nOut), int(self.hasBias), float(self.eps))

    def spec(self) -> dict:
        """Plain description (what tests hand to the oracle)."""
        kind = {1: "dense", 2: "conv3x3", 3: "batchnorm", 4: "activation", 5: "time_concat"}[self.kind]
        return {"kind": kind, "n_in": self.nIn, "n_out": self.nOut, "activation": self.activation,
                "has_bias": self.hasBias, "eps": self.eps}


def DenseLayer(nOut: int, nIn: int = 0, activation: str = "identity", hasBias: bool = True) -> Layer:
    return Layer(_lib.LAYER_DENSE, nIn, nOut, activation, hasBias)


def Convolution2D(nOut: int, nIn: int = 0, kernelSize=(3, 3), activation: str = "identity", hasBias: bool = False,
                  convolutionMode: str = "Same", stride=(1, 1)) -> Layer:
    """examples/mnist/LayerUtil.conv3x3Same (LayerUtil.java:65-72) is the only convolution inside the reference's ODE blocks."""
    if tuple(kernelSize) != (3, 3) or convolutionMode != "Same" or tuple(stride) != (1, 1) or hasBias:
        raise NotImplementedError("only Convolution2D 3x3, ConvolutionMode.Same, stride 1, hasBias(false) is in the native set")
    return Layer(_lib.LAYER_CONV3X3, nIn, nOut, activation, False)


def BatchNormalization(nOut: int = 0, activation: str = "identity", eps: float = 1e-5) -> Layer:
    return Layer(_lib.LAYER_BATCHNORM, nOut, nOut, activation, False, eps)


def ActivationLayer(activation: str) -> Layer:
    return Layer(_lib.LAYER_ACTIVATION, 0, 0, activation, False)


# ----------------------------------------------------------------------------------------------------------
# time as input to the inner graph: the reference's "timeConc" first vertex
# ----------------------------------------------------------------------------------------------------------
class MergeVertex:
    """org.deeplearning4j.nn.conf.graph.MergeVertex: concatenation along dimension 1 (the only vertex ShapeMatchVertex wraps natively)."""

    def __eq__(self, o):
        return isinstance(o, MergeVertex)


class ShapeMatchVertex:
    """ode.vertex.conf.ShapeMatchVertex (ShapeMatchVertex.java:35-50): duplicates the LAST input (the solver's current time) to the
    shape of the first one, with the sizes in ``overrideSizeDims`` replaced, and hands all inputs to the wrapped vertex.
    ``ShapeMatchVertex(MergeVertex())`` defaults to ``{1: 1}`` (one time channel / feature), the cifar10 block uses ``{1: 10}``
    (examples/cifar10/OdeBlockTime.java:38)."""

    def __init__(self, graphVertex, overrideSizeDims=None):
        if not isinstance(graphVertex, MergeVertex):
            # ShapeMatchVertex(ElementWiseVertex ...) etc. exist in the reference; only the merge is in the native set
            raise NotImplementedError("only ShapeMatchVertex(MergeVertex) is in the native set")
        self.graphVertex = graphVertex
        self.overrideSizeDims = {1: 1} if overrideSizeDims is None else {int(k): int(v) for k, v in overrideSizeDims.items()}
        if set(self.overrideSizeDims) != {1} or self.overrideSizeDims[1] < 1:
            raise NotImplementedError("the native time merge overrides the size of dimension 1 only")

    def __eq__(self, o):
        return isinstance(o, ShapeMatchVertex) and self.overrideSizeDims == o.overrideSizeDims

    def layer(self) -> Layer:
        return Layer(_lib.LAYER_TIME_CONCAT, 0, self.overrideSizeDims[1], "identity", False)
