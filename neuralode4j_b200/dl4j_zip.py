"""On-disk format of the reference's checkpoints: DL4J `ModelSerializer` zip files.

The reference saves and restores whole models with `model.save(file)` / `ModelSerializer.restoreComputationGraph(file, true)`
(examples/mnist/Main.java:140-142, examples/mnist/DeserializingModelFactory.java:34, and the cifar10 / spiral twins).  This module reads
and writes that container far enough to move the ODE block's parameter slice (and the matching updater-state slice) between a DL4J
checkpoint and the native engine:

    zip
      configuration.json   ComputationGraphConfiguration.toJson()  — the OdeVertex sits in "vertices" with "@class": "ode.vertex.conf.OdeVertex"
      coefficients.bin     Nd4j.write(model.params(), DataOutputStream)     one [1, numParams] row vector, all vertices in topological order
      updaterState.bin     Nd4j.write(updater state view)                   optional, same length conventions (Nesterovs: one momentum per param)

`Nd4j.write` = shape-info buffer then data buffer, each as `BaseDataBuffer.write`:  writeUTF(allocation mode), length (int32 for the
legacy modes DIRECT/HEAP/JAVACPP, int64 for LONG_SHAPE/MIXED_DATA_TYPES), writeUTF(data type), then the elements big-endian.
The shape-info buffer is [rank, shape..., stride..., offset/extras, elementWiseStride, order ('c' = 99, 'f' = 102)].

PARITY UNPINNED: ND4J / DL4J 1.0.0-beta3 are third-party dependencies of the reference that are not vendored in /root/reference and
cannot run here (no JVM), and the reference's tests create and delete their zip files (OdeVertexTest.java:96-118) — no fixture exists
to pin this reader against.  The layout above restates the published ND4J serialisation (both length conventions are accepted on
read; beta3's LONG_SHAPE form is written).  Tests check self-consistency (write -> read round trip, slice arithmetic against the
engine's own parameter counts), not compatibility with a real DL4J file.
"""
from __future__ import annotations

import io
import json
import struct
import zipfile
from typing import Dict, List, Optional, Tuple

import numpy as np

CONFIG_JSON = "configuration.json"
COEFFICIENTS_BIN = "coefficients.bin"
UPDATER_BIN = "updaterState.bin"
ODE_VERTEX_CLASS = "ode.vertex.conf.OdeVertex"

_LEGACY_MODES = ("DIRECT", "HEAP", "JAVACPP")                 # AllocationMode.ordinal() < 3: int32 length
_DTYPES = {"FLOAT": (">f4", 4), "DOUBLE": (">f8", 8), "INT": (">i4", 4), "LONG": (">i8", 8), "HALF": (">f2", 2)}


# ----------------------------------------------------------------------------------------------------------------------
# Nd4j.write / Nd4j.read
# ----------------------------------------------------------------------------------------------------------------------
def _write_utf(out: io.BufferedIOBase, s: str) -> None:       # DataOutputStream.writeUTF (ASCII subset: identical to modified UTF-8)
    b = s.encode("ascii")
    out.write(struct.pack(">H", len(b)))
    out.write(b)


def _read_utf(inp: io.BufferedIOBase) -> str:
    (n,) = struct.unpack(">H", inp.read(2))
    return inp.read(n).decode("ascii")


def _write_buffer(out, values: np.ndarray, dtype_name: str) -> None:
    _write_utf(out, "LONG_SHAPE")
    out.write(struct.pack(">q", int(values.size)))
    _write_utf(out, dtype_name)
    out.write(np.ascontiguousarray(values).astype(_DTYPES[dtype_name][0]).tobytes())


def _read_buffer(inp) -> np.ndarray:
    mode = _read_utf(inp)
    if mode in _LEGACY_MODES:
        (n,) = struct.unpack(">i", inp.read(4))
    else:
        (n,) = struct.unpack(">q", inp.read(8))
    dtype_name = _read_utf(inp)
    if dtype_name not in _DTYPES:
        raise ValueError(f"unsupported ND4J data type {dtype_name!r}")
    fmt, width = _DTYPES[dtype_name]
    raw = inp.read(n * width)
    if len(raw) != n * width:
        raise ValueError("truncated ND4J buffer")
    return np.frombuffer(raw, dtype=fmt).astype(fmt[1:])


def nd4j_write(arr: np.ndarray) -> bytes:
    """`Nd4j.write(arr, dos)` for a float32 array: shape-info buffer (LONG), then the data buffer (FLOAT), c-order."""
    a = np.ascontiguousarray(arr, dtype=np.float32)
    if a.ndim == 1:
        a = a.reshape(1, -1)                                  # ND4J vectors are rank 2
    strides = [int(s // a.itemsize) for s in a.strides]
    info = [a.ndim] + list(a.shape) + strides + [0, 1, ord("c")]
    out = io.BytesIO()
    _write_buffer(out, np.asarray(info, dtype=np.int64), "LONG")
    _write_buffer(out, a.ravel(), "FLOAT")
    return out.getvalue()


def nd4j_read(data: bytes) -> np.ndarray:
    """`Nd4j.read(dis)`: returns the array in its stored shape as float32 (f-ordered arrays are transposed into c-order)."""
    inp = io.BytesIO(data)
    info = _read_buffer(inp).astype(np.int64)
    rank = int(info[0])
    if rank < 1 or info.size < 2 * rank + 4:
        raise ValueError("malformed ND4J shape information")
    shape = [int(x) for x in info[1:1 + rank]]
    order = chr(int(info[2 * rank + 3]))
    vals = _read_buffer(inp)
    if vals.size != int(np.prod(shape)):
        raise ValueError(f"ND4J data length {vals.size} does not match shape {shape}")
    return np.ascontiguousarray(vals.astype(np.float32).reshape(shape, order="F" if order == "f" else "C"))


# ----------------------------------------------------------------------------------------------------------------------
# ModelSerializer zip
# ----------------------------------------------------------------------------------------------------------------------
def read_model_zip(path) -> Tuple[dict, np.ndarray, Optional[np.ndarray]]:
    """ModelSerializer.restoreComputationGraph's inputs: (configuration dict, flat params, flat updater state or None)."""
    with zipfile.ZipFile(path) as z:
        names = set(z.namelist())
        if CONFIG_JSON not in names or COEFFICIENTS_BIN not in names:
            raise ValueError(f"{path}: not a ModelSerializer zip ({CONFIG_JSON} / {COEFFICIENTS_BIN} missing)")
        conf = json.loads(z.read(CONFIG_JSON).decode("utf-8"))
        params = nd4j_read(z.read(COEFFICIENTS_BIN)).ravel()
        upd = nd4j_read(z.read(UPDATER_BIN)).ravel() if UPDATER_BIN in names else None
    return conf, params, upd


def write_model_zip(path, conf: dict, params: np.ndarray, updater_state: Optional[np.ndarray] = None) -> None:
    """ModelSerializer.writeModel(model, file, saveUpdater)."""
    with zipfile.ZipFile(path, "w", zipfile.ZIP_DEFLATED) as z:
        z.writestr(CONFIG_JSON, json.dumps(conf, indent=2))
        z.writestr(COEFFICIENTS_BIN, nd4j_write(np.asarray(params, dtype=np.float32).ravel()))
        if updater_state is not None:
            z.writestr(UPDATER_BIN, nd4j_write(np.asarray(updater_state, dtype=np.float32).ravel()))


# ----------------------------------------------------------------------------------------------------------------------
# locating the ODE block inside the flat parameter vector
# ----------------------------------------------------------------------------------------------------------------------
def _layer_of(vertex: dict) -> Optional[dict]:
    lc = vertex.get("layerConf")
    return lc.get("layer") if isinstance(lc, dict) else None


def _get(d: dict, *names, default=None):
    for n in names:
        if n in d:
            return d[n]
    return default


def layer_num_params(layer: dict) -> int:
    """numParams of one DL4J layer configuration (the layer kinds the reference's MNIST / spiral models use)."""
    cls = layer.get("@class", "").rsplit(".", 1)[-1]
    nin, nout = int(_get(layer, "nin", "nIn", default=0)), int(_get(layer, "nout", "nOut", default=0))
    bias = bool(_get(layer, "hasBias", default=True))
    if cls in ("ConvolutionLayer", "Convolution2D"):
        k = _get(layer, "kernelSize", default=[1, 1])
        return nout * nin * int(k[0]) * int(k[1]) + (nout if bias else 0)
    if cls == "BatchNormalization":
        return (2 if _get(layer, "lockGammaBeta", default=False) else 4) * nout          # gamma, beta, mean, var
    if cls in ("DenseLayer", "OutputLayer"):
        return nin * nout + (nout if bias else 0)
    if cls in ("ActivationLayer", "GlobalPoolingLayer", "CnnLossLayer", "LossLayer", "SubsamplingLayer", "DropoutLayer"):
        return 0
    raise NotImplementedError(f"parameter count of layer class {layer.get('@class')!r}")


def vertex_num_params(vertex: dict) -> int:
    cls = vertex.get("@class", "")
    if cls == ODE_VERTEX_CLASS:                                # conf/OdeVertex.java:85-89: sum over the inner graph's vertices
        return sum(vertex_num_params(v) for v in vertex["conf"]["vertices"].values())
    layer = _layer_of(vertex)
    return layer_num_params(layer) if layer is not None else 0


def topological_order(conf: dict) -> List[str]:
    """ComputationGraphConfiguration.topologicalOrdering(): Kahn's algorithm over `vertexInputs`, network inputs first, ties in
    insertion order of the `vertices` map (DL4J keeps LinkedHashMaps, JSON keeps their order)."""
    inputs = list(conf.get("networkInputs", []))
    vin: Dict[str, List[str]] = {k: list(v) for k, v in conf.get("vertexInputs", {}).items()}
    names = inputs + [n for n in conf["vertices"] if n not in inputs]
    indeg = {n: len(vin.get(n, [])) for n in names}
    out_edges: Dict[str, List[str]] = {n: [] for n in names}
    for n, srcs in vin.items():
        for s in srcs:
            out_edges.setdefault(s, []).append(n)
    queue = [n for n in names if indeg[n] == 0]
    order = []
    while queue:
        n = queue.pop(0)
        order.append(n)
        for m in out_edges.get(n, []):
            indeg[m] -= 1
            if indeg[m] == 0:
                queue.append(m)
    if len(order) != len(names):
        raise ValueError("configuration.json: vertex graph has a cycle or dangling input")
    return order


def ode_block_slices(conf: dict) -> Dict[str, Tuple[int, int]]:
    """{OdeVertex name: (offset, length)} of every ODE block in the flat parameter vector (ComputationGraph.init hands out the
    parameter views vertex by vertex in topological order)."""
    off, out = 0, {}
    for name in topological_order(conf):
        v = conf["vertices"].get(name)
        if v is None:
            continue
        n = vertex_num_params(v)
        if v.get("@class") == ODE_VERTEX_CLASS:
            out[name] = (off, n)
        off += n
    return out


def extract_ode_block(path, vertex_name: Optional[str] = None):
    """(vertex name, its DL4J configuration dict, parameter slice, updater-state slice or None) of one ODE block of a checkpoint.
    The parameter slice is in DL4J's flat view order, i.e. exactly what node4j_bind_params / OdeVertexImpl.setParams take."""
    conf, params, upd = read_model_zip(path)
    slices = ode_block_slices(conf)
    if not slices:
        raise ValueError(f"{path}: no {ODE_VERTEX_CLASS} in the configuration")
    name = vertex_name or next(iter(slices))
    off, n = slices[name]
    if off + n > params.size:
        raise ValueError(f"{path}: parameter vector has {params.size} entries, the ODE block needs [{off}, {off + n})")
    u = None
    if upd is not None and upd.size == params.size:            # Nesterovs / Sgd-like: one state value per parameter, same layout
        u = upd[off:off + n].copy()
    return name, conf["vertices"][name], params[off:off + n].copy(), u


def replace_ode_block(src_path, dst_path, block_params: np.ndarray, vertex_name: Optional[str] = None, block_updater_state=None) -> None:
    """Writes a copy of checkpoint `src_path` whose ODE-block parameter slice (and optionally updater-state slice) is replaced."""
    conf, params, upd = read_model_zip(src_path)
    slices = ode_block_slices(conf)
    name = vertex_name or next(iter(slices))
    off, n = slices[name]
    block_params = np.asarray(block_params, dtype=np.float32).ravel()
    if block_params.size != n:
        raise ValueError(f"ODE block {name!r} has {n} parameters, got {block_params.size}")
    params = params.copy()
    params[off:off + n] = block_params
    if block_updater_state is not None:
        if upd is None or upd.size != params.size:
            raise ValueError("checkpoint has no per-parameter updater state to replace")
        upd = upd.copy()
        upd[off:off + n] = np.asarray(block_updater_state, dtype=np.float32).ravel()
    write_model_zip(dst_path, conf, params, upd)


# ----------------------------------------------------------------------------------------------------------------------
# DL4J-shaped OdeVertex JSON -> native configuration
# ----------------------------------------------------------------------------------------------------------------------
_ACT = {"ActivationIdentity": "identity", "ActivationReLU": "relu", "ActivationELU": "elu", "ActivationTanH": "tanh"}


def _activation(layer: dict) -> str:
    a = layer.get("activationFn") or {}
    cls = a.get("@class", "ActivationIdentity").rsplit(".", 1)[-1]
    if cls not in _ACT:
        raise NotImplementedError(f"activation {cls} is outside the native set")
    return _ACT[cls]


def ode_vertex_from_dl4j(vertex: dict):
    """`ode.vertex.conf.OdeVertex` as DL4J's Jackson mapper writes it (fields of conf/OdeVertex.java:50-72: conf, firstVertex,
    odeForwardConf, odeBackwardConf, ...) -> the native `neuralode4j_b200.OdeVertex` configuration.  Layers outside the native set
    raise NotImplementedError (no CPU fallback)."""
    from .conf import (ActivationLayer, BatchNormalization, Convolution2D, DenseLayer, DormandPrince54Solver, FixedStep, InputStep, SolverConfig)
    from .vertex import OdeVertex
    if vertex.get("@class") != ODE_VERTEX_CLASS:
        raise ValueError("not an ode.vertex.conf.OdeVertex")
    inner = vertex["conf"]
    order = [n for n in topological_order(inner) if n in inner["vertices"]]
    builder = None
    prev = None
    for name in order:
        layer = _layer_of(inner["vertices"][name])
        if layer is None:
            raise NotImplementedError(f"inner vertex {name!r} is not a layer vertex")
        cls = layer["@class"].rsplit(".", 1)[-1]
        nin, nout = int(_get(layer, "nin", "nIn", default=0)), int(_get(layer, "nout", "nOut", default=0))
        act = _activation(layer)
        if cls == "BatchNormalization":
            l = BatchNormalization(nout, act, eps=float(_get(layer, "eps", default=1e-5)))
        elif cls in ("ConvolutionLayer", "Convolution2D"):
            k, s = _get(layer, "kernelSize", default=[3, 3]), _get(layer, "stride", default=[1, 1])
            if list(k) != [3, 3] or list(s) != [1, 1] or _get(layer, "convolutionMode") != "Same" or _get(layer, "hasBias", default=True):
                raise NotImplementedError(f"convolution {name!r} must be 3x3, stride 1, Same, no bias for the native path")
            l = Convolution2D(nout, nin, activation=act)
        elif cls == "DenseLayer":
            l = DenseLayer(nout, nin, act, hasBias=bool(_get(layer, "hasBias", default=True)))
        elif cls == "ActivationLayer":
            l = ActivationLayer(act)
        else:
            raise NotImplementedError(f"layer class {layer['@class']} is outside the native set")
        if builder is None:
            builder = OdeVertex.Builder(None, name, l)
        else:
            builder.addLayer(name, l, prev)
        prev = name

    def solver_of(h):
        sc = _get(h, "solverConf", "solver")
        if not sc or not sc.get("@class", "").endswith("DormandPrince54Solver"):
            raise NotImplementedError("only DormandPrince54Solver is native")
        c = sc["config"]
        return DormandPrince54Solver(SolverConfig(float(_get(c, "absoluteTolerance", "absTol")), float(_get(c, "relativeTolerance", "relTol")),
                                                  float(c["minStep"]), float(c["maxStep"])))

    fwd, bwd = vertex["odeForwardConf"], vertex["odeBackwardConf"]
    fcls = fwd.get("@class", "").rsplit(".", 1)[-1]
    if fcls == "FixedStep":
        t = fwd["time"]
        t = t.get("data", t) if isinstance(t, dict) else t
        helper = FixedStep(solver_of(fwd), np.asarray(t, dtype=np.float32).ravel(), bool(_get(fwd, "interpolateIfMultiStep", default=False)))
    elif fcls == "InputStep":
        helper = InputStep(solver_of(fwd), int(fwd["timeInputIndex"]), bool(_get(fwd, "interpolateIfMultiStep", default=False)),
                           bool(_get(bwd, "needTimeGradient", default=False)))
    else:
        raise NotImplementedError(f"forward helper {fwd.get('@class')!r}")
    return builder.odeConf(helper).build()
