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


def _act(name):
    return {"@class": "org.nd4j.linalg.activations.impl." + name}


def _layer_vertex(layer):
    return {"@class": "org.deeplearning4j.nn.conf.graph.LayerVertex", "layerConf": {"layer": layer}}


def _conv(nin, nout, k, stride, mode, bias=False, act="ActivationIdentity"):
    return _layer_vertex({"@class": "org.deeplearning4j.nn.conf.layers.ConvolutionLayer", "nin": nin, "nout": nout, "kernelSize": [k, k],
                          "stride": [stride, stride], "hasBias": bias, "convolutionMode": mode, "activationFn": _act(act)})


def _bn(n):
    return _layer_vertex({"@class": "org.deeplearning4j.nn.conf.layers.BatchNormalization", "nin": n, "nout": n, "eps": 1e-5, "decay": 0.9,
                          "lockGammaBeta": False, "activationFn": _act("ActivationReLU")})


def odenet_configuration(nrofKernels: int = 64, solver: DormandPrince54Solver | None = None, time=(0.0, 1.0)) -> dict:
    """The reference's MNIST ODE-net as a ComputationGraphConfiguration in DL4J's JSON shape — what `configuration.json` of a
    ModelSerializer checkpoint of that model holds, as far as this package reads it (neuralode4j_b200/dl4j_zip.py): vertex names, inputs
    and layer fields of examples/mnist/ResBlockStem.java:25-48, OdeNetModel.java:72-88, Output.java:27-35, LayerUtil.java:54-154, in
    the builders' insertion order (DL4J's topological ordering breaks ties by it, and hands out the flat parameter views in that order —
    the order `node4j_net_*` uses).  Unpinned against a real DL4J file like the rest of row f4 (no JVM here)."""
    C = int(nrofKernels)
    solver = solver or DormandPrince54Solver(SolverConfig(1e-3, 1e-3, 1e-20, 100))
    sc = {"@class": "ode.solve.conf.DormandPrince54Solver", "config": solver.config.to_json()}
    t = [float(time[0]), float(time[1])]
    vertices, inputs = {}, {}

    def add(name, vertex, *src):
        vertices[name] = vertex
        inputs[name] = list(src)

    add("inputConv", _conv(1, C, 3, 1, "Truncate", bias=True), "input")
    prev = "inputConv"
    for r in (0, 1):
        add(f"firstNormStem_{r}", _bn(C), prev)
        add(f"downSample_{r}", _conv(C, C, 1, 2, "Truncate"), f"firstNormStem_{r}")
        add(f"firstConvStem_{r}", _conv(C, C, 3, 2, "Same", act="ActivationReLU"), f"firstNormStem_{r}")
        add(f"secondNormStem_{r}", _bn(C), f"firstConvStem_{r}")
        add(f"secondConvStem_{r}", _conv(C, C, 3, 1, "Same"), f"secondNormStem_{r}")
        add(f"stemAdd_{r}", {"@class": "org.deeplearning4j.nn.conf.graph.ElementWiseVertex", "op": "Add"}, f"downSample_{r}", f"secondConvStem_{r}")
        prev = f"stemAdd_{r}"
    inner = {"networkInputs": ["input"], "networkOutputs": ["normThird"],
             "vertexInputs": {"normFirst": ["input"], "convFirst": ["normFirst"], "normSecond": ["convFirst"], "convSecond": ["normSecond"],
                              "normThird": ["convSecond"]},
             "vertices": {"normFirst": _bn(C), "convFirst": _conv(C, C, 3, 1, "Same"), "normSecond": _bn(C), "convSecond": _conv(C, C, 3, 1, "Same"),
                          "normThird": _bn(C)}}
    add("odeBlock", {"@class": "ode.vertex.conf.OdeVertex", "conf": inner, "firstVertex": "normFirst",
                     "odeForwardConf": {"@class": "ode.vertex.conf.helper.forward.FixedStep", "solverConf": sc, "time": t, "interpolateIfMultiStep": False},
                     "odeBackwardConf": {"@class": "ode.vertex.conf.helper.backward.FixedStepAdjoint", "solverConf": sc, "time": t}}, prev)
    add("normOutput", _bn(C), "odeBlock")
    add("globPool", _layer_vertex({"@class": "org.deeplearning4j.nn.conf.layers.GlobalPoolingLayer", "poolingType": "AVG"}), "normOutput")
    add("output", _layer_vertex({"@class": "org.deeplearning4j.nn.conf.layers.OutputLayer", "nin": C, "nout": 10, "hasBias": True,
                                 "activationFn": _act("ActivationSoftmax"), "lossFn": {"@class": "org.nd4j.linalg.lossfunctions.impl.LossMCXENT"}}), "globPool")
    return {"networkInputs": ["input"], "networkOutputs": ["output"], "vertexInputs": inputs, "vertices": vertices}


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
        self._solver, self._time = solver, (float(time[0]), float(time[1]))
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

    # ---- checkpoints: model.save(file) / ModelSerializer.restoreComputationGraph(file, true) (examples/mnist/Main.java:140-142) ----
    def configuration(self) -> dict:
        return odenet_configuration(self.channels, self._solver, self._time)

    def save(self, path, saveUpdater: bool = True) -> None:
        """ModelSerializer zip: configuration.json, coefficients.bin (the flat parameter vector), updaterState.bin (Nesterovs momenta)."""
        from . import dl4j_zip
        dl4j_zip.write_model_zip(path, self.configuration(), self.params, self.updater_state if saveUpdater else None)

    def load(self, path, loadUpdater: bool = True) -> None:
        """Takes parameters (and updater state) from a checkpoint of the same model; the configuration must place the ODE block where this
        model has it and hold as many parameters."""
        from . import dl4j_zip
        conf, params, upd = dl4j_zip.read_model_zip(path)
        slices = dl4j_zip.ode_block_slices(conf)
        if len(slices) != 1 or next(iter(slices.values())) != self.ode_slice or params.size != self.params.size:
            raise ValueError(f"{path}: not a checkpoint of this model (ODE block at {slices}, {params.size} parameters; "
                             f"expected {self.ode_slice}, {self.params.size})")
        self.setParams(params)
        if loadUpdater and upd is not None and upd.size == self.params.size:
            self.updater_state = np.ascontiguousarray(upd, dtype=np.float32).copy()

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
