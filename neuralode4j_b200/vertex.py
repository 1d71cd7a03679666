"""OdeVertex: configuration (+ Builder) and runtime vertex backed by the native engine.

Mirrors
  ode/vertex/conf/OdeVertex.java:57-148,164-336   (conf, Builder: addLayer / odeConf / odeForward / odeBackward / build)
  ode/vertex/impl/OdeVertex.java:67-156           (runtime: params, numParams, setBackpropGradientsViewArray,
                                                   doForward, doBackward)
  ode/vertex/impl/helper/OdeGraphHelper.java:121-156  (lastOutput bookkeeping, nfe counter)
under /root/reference/src/main/java.  The inner graph must be a chain of layers from the native set (conf.py);
anything else raises NotImplementedError at build time — there is no CPU fallback.

Arrays may be numpy float32 arrays (host) or torch CUDA tensors (device pointers go straight through the C ABI).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from .conf import (DormandPrince54Solver, FixedStep, InputStep, Layer, ShapeMatchVertex, _ReplayState, ode_helper_from_json)


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _ptr(x) -> C.c_void_p:
    if x is None:
        return C.c_void_p(0)
    if _is_torch(x):
        assert x.is_contiguous() and str(x.dtype) == "torch.float32", "need a contiguous float32 tensor"
        return C.c_void_p(x.data_ptr())
    assert isinstance(x, np.ndarray) and x.dtype == np.float32 and x.flags["C_CONTIGUOUS"], "need a contiguous float32 array"
    return C.c_void_p(x.ctypes.data)


def _as_f32(x):
    if _is_torch(x):
        return x.contiguous().float() if (not x.is_contiguous() or str(x.dtype) != "torch.float32") else x
    return np.ascontiguousarray(np.asarray(x, dtype=np.float32))


def _empty_like(x, shape):
    if _is_torch(x):
        import torch
        return torch.empty(tuple(shape), dtype=torch.float32, device=x.device)
    return np.empty(tuple(shape), dtype=np.float32)


class OdeVertex:
    """ode.vertex.conf.OdeVertex — configuration of one ODE block."""

    JCLASS = "ode.vertex.conf.OdeVertex"

    def __init__(self, layer_names: List[str], layers: List[Layer], first: str, odeForwardConf, odeBackwardConf):
        self.layer_names = layer_names
        self.layers = layers
        self.first = first
        self.odeForwardConf = odeForwardConf
        self.odeBackwardConf = odeBackwardConf

    class Builder:
        """OdeVertex.Builder(globalConf, name, layer) (OdeVertex.java:164-336)."""

        def __init__(self, globalConf, name: str, layer, timeAsInput: bool = False, *otherInputs):
            self.globalConf = globalConf
            self.first = name
            self._names = []
            self._layers = []
            if isinstance(layer, Layer):          # Builder(globalConf, name, layer) :178-185
                self._names.append(name)
                self._layers.append(layer)
            elif timeAsInput:                      # Builder(globalConf, name, vertex, timeAsInput, ...) :194-208
                self.addTimeVertex(name, layer, "input")
            else:
                self.addVertex(name, layer, "input")
            default = FixedStep(DormandPrince54Solver(), np.arange(2), True)   # :168-169
            self.odeForwardConf = default
            self.odeBackwardConf = default

        def addLayer(self, name: str, layer: Layer, *inputs: str):
            if len(inputs) != 1 or inputs[0] != self._names[-1]:
                raise NotImplementedError("native ODE blocks are chains: each layer must take the previous layer as its only input")
            self._names.append(name)
            self._layers.append(layer)
            return self

        def addVertex(self, name, vertex, *inputs):
            raise NotImplementedError("GraphVertex inside an ODE block is outside the native set (except the time merge, addTimeVertex)")

        def addTimeLayer(self, name, layer):
            raise NotImplementedError("a layer fed by the time alone (addTimeLayer) is outside the native set; use the time merge vertex")

        def addTimeVertex(self, name, vertex, *inputs):
            """addTimeVertex (OdeVertex.java:243-246): the vertex also receives the solver's current time as its last input.  Native
            for ShapeMatchVertex(MergeVertex) as the FIRST vertex of the block (the reference's examples/anode and examples/cifar10
            blocks, OdeBlockWithTime.java:30-36, OdeBlockTime.java:35-41)."""
            if not isinstance(vertex, ShapeMatchVertex):
                raise NotImplementedError("only ShapeMatchVertex(MergeVertex) can take the time input natively")
            if self._names:
                raise NotImplementedError("the time merge is only supported as the first vertex of the inner graph")
            self._names.append(name)
            self._layers.append(vertex.layer())
            return self

        def odeConf(self, odeConf):            # :271-274
            self.odeForward(odeConf.forward())
            return self.odeBackward(odeConf.backward())

        def odeForward(self, odeForwardConf):  # :282-285
            self.odeForwardConf = odeForwardConf
            return self

        def odeBackward(self, odeBackwardConf):  # :293-296
            self.odeBackwardConf = odeBackwardConf
            return self

        def build(self) -> "OdeVertex":
            return OdeVertex(list(self._names), list(self._layers), self.first, self.odeForwardConf, self.odeBackwardConf)

    # ---- conf API ----------------------------------------------------------------------------------------
    def clone(self):
        return OdeVertex(list(self.layer_names), [Layer(**vars(l)) for l in self.layers], self.first,
                         self.odeForwardConf.clone(), self.odeBackwardConf.clone())

    def __eq__(self, o):
        return (isinstance(o, OdeVertex) and self.layer_names == o.layer_names and self.layers == o.layers
                and self.odeForwardConf == o.odeForwardConf and self.odeBackwardConf == o.odeBackwardConf)

    def numParams(self, input_shape: Sequence[int]) -> int:
        c = input_shape[1]
        n = 0
        for l in self.layers:
            if l.kind == _lib.LAYER_DENSE:
                n += c * l.nOut + (l.nOut if l.hasBias else 0)
                c = l.nOut
            elif l.kind == _lib.LAYER_CONV3X3:
                n += l.nOut * c * 9
                c = l.nOut
            elif l.kind == _lib.LAYER_BATCHNORM:
                n += 4 * c
            elif l.kind == _lib.LAYER_TIME_CONCAT:
                c += l.nOut
        return n

    def activationType(self, inp):
        """Type of the inner graph's output for one input (ComputationGraphConfiguration.getLayerActivationTypes restricted to the
        chain of layers the native set allows): 3x3 Same stride-1 convolutions keep height and width, the time merge adds features."""
        kind, shape = inp.kind, list(inp.shape)
        if kind not in ("FF", "CNN"):
            raise ValueError(f"input type {inp} is not supported inside a native OdeVertex")
        for l in self.layers:
            if l.kind == _lib.LAYER_DENSE:
                if kind != "FF":
                    raise ValueError("DenseLayer inside an OdeVertex needs feed-forward activations")
                shape[0] = l.nOut
            elif l.kind == _lib.LAYER_CONV3X3:
                if kind != "CNN":
                    raise ValueError("Convolution2D inside an OdeVertex needs convolutional activations")
                shape[0] = l.nOut
            elif l.kind == _lib.LAYER_TIME_CONCAT:
                shape[0] += l.nOut
        return type(inp)(kind, tuple(shape))

    def getOutputType(self, layerIndex, *vertexInputs):
        """OdeVertex.getOutputType (conf/OdeVertex.java:151-154): delegated to the forward helper conf."""
        return self.odeForwardConf.getOutputType(self, *vertexInputs)

    def graph_spec(self) -> List[dict]:
        return [l.spec() for l in self.layers]

    def to_json(self):
        return {"@class": self.JCLASS, "first": self.first,
                "layers": [dict(name=n, **l.spec()) for n, l in zip(self.layer_names, self.layers)],
                "odeForwardConf": self.odeForwardConf.to_json(), "odeBackwardConf": self.odeBackwardConf.to_json()}

    @staticmethod
    def from_json(d):
        kinds = {"dense": _lib.LAYER_DENSE, "conv3x3": _lib.LAYER_CONV3X3, "batchnorm": _lib.LAYER_BATCHNORM,
                 "activation": _lib.LAYER_ACTIVATION, "time_concat": _lib.LAYER_TIME_CONCAT}
        layers = [Layer(kinds[l["kind"]], l["n_in"], l["n_out"], l["activation"], l["has_bias"], l["eps"]) for l in d["layers"]]
        return OdeVertex([l["name"] for l in d["layers"]], layers, d["first"], ode_helper_from_json(d["odeForwardConf"]),
                         ode_helper_from_json(d["odeBackwardConf"]))

    def instantiate(self, input_shape: Sequence[int], device: int = 0, flags: int = 0, max_trials: int = 0) -> "OdeVertexImpl":
        """OdeVertex.instantiate (OdeVertex.java:112-141): creates the runtime vertex for a given (local) input shape."""
        return OdeVertexImpl(self, input_shape, device, flags, max_trials)


class OdeVertexImpl:
    """ode.vertex.impl.OdeVertex backed by node4j_* calls."""

    def __init__(self, conf: OdeVertex, input_shape: Sequence[int], device: int = 0, flags: int = 0, max_trials: int = 0):
        self.conf = conf
        self.shape = tuple(int(s) for s in input_shape)
        if len(self.shape) not in (2, 4):
            raise NotImplementedError(f"Rank not supported: {len(self.shape)}")   # MultiStep.java:57-59
        if conf.odeForwardConf.solver != conf.odeBackwardConf.solver:
            raise NotImplementedError("forward and backward helpers must share one solver configuration in the native path")
        self._L = _lib.load()
        self._layers = (_lib.LayerDesc * len(conf.layers))(*[l.desc() for l in conf.layers])
        g = _lib.GraphDesc(len(conf.layers), len(self.shape), (C.c_int64 * 4)(*(list(self.shape) + [0] * (4 - len(self.shape)))), self._layers)
        cfg = conf.odeForwardConf.solver.native_cfg(flags, max_trials)
        h = C.c_void_p()
        _lib.check(self._L.node4j_create(C.byref(g), C.byref(cfg), device, C.byref(h)))
        self._h = h
        self._n_params = self._L.node4j_num_params(h)
        self._n_real = self._L.node4j_num_real_params(h)
        self._params = None
        self._grad = None
        self._inputs = None
        self._epsilon = None
        self._last_time = None
        self.fwd_stats = _lib.Stats()
        self.bwd_stats = _lib.Stats()

    def close(self):
        if getattr(self, "_h", None):
            self._L.node4j_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- DL4J GraphVertex surface -------------------------------------------------------------------------
    def numParams(self) -> int:
        return int(self._n_params)

    def numRealParams(self) -> int:
        return int(self._n_real)

    def setParams(self, params):
        """Binds the flat parameter view shared with the outer network (OdeVertex.java:124-129)."""
        p = _as_f32(params)
        n = p.numel() if _is_torch(p) else p.size
        if n != self._n_params:
            raise ValueError(f"expected {self._n_params} parameters")
        self._params = p
        _lib.check(self._L.node4j_bind_params(self._h, _ptr(p), self._n_params))

    def params(self):
        return self._params

    def setBackpropGradientsViewArray(self, grad):
        g = grad
        n = g.numel() if _is_torch(g) else g.size
        if n != self._n_params:
            raise ValueError(f"gradient view must have {self._n_params} entries")
        _ptr(g)   # validates dtype/contiguity: the view is written in place
        self._grad = g

    def getGradientsViewArray(self):
        return self._grad

    def setInputs(self, *inputs):
        self._inputs = [_as_f32(i) for i in inputs]

    def setEpsilon(self, epsilon):
        self._epsilon = _as_f32(epsilon)

    def clear(self):   # OdeGraphHelper.clear :93-96
        self._inputs = None
        self._epsilon = None

    # ---- helpers -----------------------------------------------------------------------------------------
    def _time_and_y0(self):
        helper = self.conf.odeForwardConf
        if self._inputs is None:
            raise RuntimeError("Cannot do forward pass: inputs not set")   # OdeVertex.validateForward
        if helper.time_input_index is None:
            return self._inputs[0], np.ascontiguousarray(helper.time, dtype=np.float32)
        idx = helper.time_input_index
        others = [x for i, x in enumerate(self._inputs) if i != idx]   # NoTimeInput.removeInput :52-57
        t = self._inputs[idx]
        t = t.detach().cpu().numpy() if _is_torch(t) else t
        return others[0], np.ascontiguousarray(np.asarray(t, dtype=np.float32).ravel())

    def _replay_listeners(self, solver, t_pairs):
        if not solver.listeners:
            return
        n = self._L.node4j_get_step_log(self._h, None, 0)
        buf = (_lib.StepRec * max(int(n), 1))()
        self._L.node4j_get_step_log(self._h, buf, n)
        # one record per trial, tagged with the solve (segment) it belongs to: begin/done boundaries come from that tag, not from
        # comparing times (t + (t_end - t) need not equal t_end in fp32).  Records beyond the engine's log capacity are not replayed.
        held = min(int(n), len(buf))
        i = 0
        for k, (t0, t1) in enumerate(t_pairs):
            for l in solver.listeners:
                l.begin(np.array([t0, t1], dtype=np.float32), None)
            while i < held and buf[i].solve == k and buf[i].h != 0.0:     # h == 0: past the engine's log capacity
                r = buf[i]
                i += 1
                if r.accepted:
                    for l in solver.listeners:
                        l.step(_ReplayState(r.t), r.h, r.err)
            for l in solver.listeners:
                l.done()

    # ---- forward / backward ----------------------------------------------------------------------------------
    def doForward(self, training: bool = True):
        """OdeVertex.doForward :111-121 -> OdeGraphHelper.doForward :121-132.  Always training mode (SingleStep.java:37)."""
        helper = self.conf.odeForwardConf
        y0, t = self._time_and_y0()
        if tuple(y0.shape) != self.shape:
            raise ValueError(f"input shape {tuple(y0.shape)} does not match the instantiated shape {self.shape}")
        if self._params is None and self._n_params > 0:
            raise RuntimeError("parameters not set")
        nt = int(t.size)
        if nt == 2:
            out_shape = self.shape
        elif len(self.shape) == 2:
            out_shape = (self.shape[0], self.shape[1], nt)             # [B,H,T]
        else:
            out_shape = (self.shape[0], nt) + self.shape[1:]           # [B,T,C,H,W]
        out = _empty_like(y0, out_shape)
        _lib.check(self._L.node4j_forward(self._h, _ptr(y0), _ptr(t), nt, int(helper.interpolateForwardIfMultiStep), _ptr(out),
                                          C.byref(self.fwd_stats)))
        self._last_time = t
        if nt == 2 or helper.interpolateForwardIfMultiStep:
            pairs = [(t[0], t[-1])]
        else:
            pairs = [(t[k - 1], t[k]) for k in range(1, nt)]
        self._replay_listeners(helper.solver, pairs)
        return out

    def doBackward(self, tbptt: bool = False, last_output=None):
        """OdeVertex.doBackward :124-140 -> OdeGraphHelper.doBackward :134-156.

        Returns ``(gradient_view, epsilons)``; epsilons has one entry (FixedStep) or two (InputStep: state epsilon and
        time gradient placed at ``timeInputIndex``)."""
        helper = self.conf.odeBackwardConf
        if self._inputs is None:
            raise RuntimeError("Cannot do backward pass: inputs not set.")
        if self._epsilon is None:
            raise RuntimeError("Cannot do backward pass: all epsilons not set.")
        y0, t = self._time_and_y0()
        nt = int(t.size)
        if self._grad is None:
            self._grad = np.zeros(self._n_params, dtype=np.float32) if not _is_torch(y0) else __import__("torch").zeros(
                self._n_params, dtype=__import__("torch").float32, device=y0.device)
        eps = self._epsilon
        dy0 = _empty_like(y0, self.shape)
        if helper.time_input_index is None:
            mode, dt = _lib.TIMEGRAD_NONE, None
        else:
            mode = _lib.TIMEGRAD_CALC if helper.needTimeGradient else _lib.TIMEGRAD_ZERO
            dt = np.zeros(nt, dtype=np.float32)
        lo = _as_f32(last_output) if last_output is not None else None
        _lib.check(self._L.node4j_backward(self._h, _ptr(lo), _ptr(eps), _ptr(t), nt, mode, _ptr(self._grad), _ptr(dy0),
                                           _ptr(dt), C.byref(self.bwd_stats)))
        pairs = [(t[k], t[k - 1]) for k in range(nt - 1, 0, -1)]
        self._replay_listeners(helper.solver, pairs)
        if helper.time_input_index is None:
            return self._grad, [dy0]
        epsilons = [None, None]
        epsilons[helper.time_input_index] = dt.reshape(1, nt)
        epsilons[(helper.time_input_index + 1) % 2] = dy0
        return self._grad, epsilons

    # ---- multi-GPU ---------------------------------------------------------------------------------------------
    def comm_export(self) -> bytes:
        """node4j_comm_export: the 64-byte IPC handle of this rank's exchange region.  The host framework all-gathers the handles
        of all ranks (any transport: MPI, a file, torch.distributed, the JVM's own RPC) and passes them to `init_comm_handles`."""
        mine = C.create_string_buffer(64)
        _lib.check(self._L.node4j_comm_export(self._h, mine))
        return mine.raw

    def init_comm_handles(self, rank: int, world: int, handles) -> None:
        """node4j_comm_init with pre-gathered handles (rank order): no torch on this path.  `handles` is a sequence of `world`
        64-byte strings or one bytes object of 64 * world bytes.  Every rank must have called `comm_export` first and must not
        launch work on the handle until all ranks have returned from this call (the caller's barrier)."""
        raw = bytes(handles) if isinstance(handles, (bytes, bytearray)) else b"".join(bytes(h) for h in handles)
        if len(raw) != 64 * world:
            raise ValueError(f"expected {world} handles of 64 bytes, got {len(raw)} bytes")
        if not 0 <= rank < world:
            raise ValueError("rank out of range")
        allh = C.create_string_buffer(raw, 64 * world)
        _lib.check(self._L.node4j_comm_init(self._h, rank, world, allh, self.shape[0] * world))

    def init_comm(self, group=None):
        """Convenience for hosts that already run torch.distributed (bench.py, the tests): row-sharded minibatch over the ranks of a
        process group (one process per GPU, one box).  torch.distributed only carries the 64-byte IPC handles and the barrier; the data
        path exchanges run inside the engine's kernels over NVLink peer memory.  Hosts without torch use comm_export / init_comm_handles."""
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        gathered = [None] * world
        dist.all_gather_object(gathered, self.comm_export(), group=group)
        self.init_comm_handles(rank, world, gathered)
        dist.barrier(group)
        return rank, world

    # ---- extras ------------------------------------------------------------------------------------------------
    def step_log(self):
        n = int(self._L.node4j_get_step_log(self._h, None, 0))
        buf = (_lib.StepRec * max(n, 1))()
        self._L.node4j_get_step_log(self._h, buf, n)
        m = min(n, len(buf))
        return {"t": np.array([buf[i].t for i in range(m)], dtype=np.float32), "h": np.array([buf[i].h for i in range(m)], dtype=np.float32),
                "err": np.array([buf[i].err for i in range(m)], dtype=np.float32), "accepted": np.array([buf[i].accepted for i in range(m)]),
                "solve": np.array([buf[i].solve for i in range(m)])}

    def rhs(self, z, t: float = 0.0):
        z = _as_f32(z)
        out = _empty_like(z, self.shape)
        _lib.check(self._L.node4j_rhs(self._h, _ptr(z), float(t), _ptr(out)))
        return out

    def rhs_vjp(self, z, g, t: float = 0.0):
        z, g = _as_f32(z), _as_f32(g)
        fz, dz = _empty_like(z, self.shape), _empty_like(z, self.shape)
        dreal = np.zeros(max(self._n_real, 1), dtype=np.float32)
        _lib.check(self._L.node4j_rhs_vjp(self._h, _ptr(z), float(t), _ptr(g), _ptr(fz), _ptr(dz), _ptr(dreal)))
        return fz, dz, dreal[: self._n_real]

    def rhs_vjp_time(self, z, g, t: float = 0.0):
        """(f(z,t), g^T df/dz, g^T df/dtheta_real, g^T df/dt): the last entry is the epsilon of the time input (TimeInput.java:46-50)."""
        z, g = _as_f32(z), _as_f32(g)
        fz, dz = _empty_like(z, self.shape), _empty_like(z, self.shape)
        dreal = np.zeros(max(self._n_real, 1), dtype=np.float32)
        dt = C.c_float()
        _lib.check(self._L.node4j_rhs_vjp_time(self._h, _ptr(z), float(t), _ptr(g), _ptr(fz), _ptr(dz), _ptr(dreal), C.byref(dt)))
        return fz, dz, dreal[: self._n_real], dt.value

    def error_norm(self, K, y0, y1, h):
        K, y0, y1 = _as_f32(K), _as_f32(y0), _as_f32(y1)
        out = C.c_float()
        n = y0.numel() if _is_torch(y0) else y0.size
        _lib.check(self._L.node4j_error_norm(self._h, _ptr(K), _ptr(y0), _ptr(y1), float(h), n, C.byref(out)))
        return out.value

    def bench_piece(self, what: int, iters: int = 20) -> float:
        us = C.c_float()
        _lib.check(self._L.node4j_bench_piece(self._h, what, iters, C.byref(us)))
        return us.value

    def cuda_stream(self) -> int:
        s = C.c_void_p()
        _lib.check(self._L.node4j_get_stream(self._h, C.byref(s)))
        return int(s.value or 0)

    def bench_solver_pass(self, n: int, iters: int = 20) -> float:
        us = C.c_float()
        _lib.check(self._L.node4j_bench_solver_pass(self._h, n, iters, C.byref(us)))
        return us.value
