"""CPU tests of the host-side mirror of the reference interface and of the C-ABI library surface (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest

import neuralode4j_b200 as N
from neuralode4j_b200 import _lib, build as nbuild
from neuralode4j_b200 import workloads as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "node4j.h")).read()
    declared = sorted(set(re.findall(r"\b(node4j_[a-z_0-9]+)\s*\(", header)))
    assert declared, "no entry points found in include/node4j.h"
    L = _lib.load()
    for name in declared:
        assert hasattr(L, name), f"libnode4j_b200.so does not export {name}"
    assert sorted(_lib.EXPORTS) == declared
    assert L.node4j_abi_version() == 2


def test_library_is_built_for_sm_100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", nbuild.LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out)


def test_no_cpu_fallback_without_device():
    L = _lib.load()
    if L.node4j_device_count() > 0:
        pytest.skip("a CUDA device is present")
    w = W.golden_linear()
    with pytest.raises(_lib.Node4jError) as e:
        w.conf.instantiate(w.shape)
    assert e.value.code == _lib.ERR_CUDA


def test_product_never_imports_the_oracle():
    """Nothing under neuralode4j_b200/ (or the public header) may import, include, dlopen or spawn anything under oracle/: python files
    are checked through their syntax tree (imports, and string constants that name the oracle's files), native sources through their
    #include lines and any dlopen / system / popen call.  Comments may talk about the oracle; code may not touch it."""
    import ast
    pkg = os.path.join(ROOT, "neuralode4j_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            path = os.path.join(dirpath, f)
            if f.endswith(".py"):
                tree = ast.parse(open(path).read())
                for node in ast.walk(tree):
                    if isinstance(node, ast.Import):
                        assert not any(a.name.split(".")[0] == "oracle" for a in node.names), path
                    elif isinstance(node, ast.ImportFrom):
                        assert (node.module or "").split(".")[0] != "oracle", path
                    elif isinstance(node, ast.Constant) and isinstance(node.value, str) and not isinstance(getattr(node, "parent", None), ast.Expr):
                        if node.value.lstrip().startswith(('"""', "On-disk", "The reference")) or "\n" in node.value:
                            continue                                   # docstrings / prose
                        assert "libnode_oracle" not in node.value and "node_oracle" not in node.value and "outer_oracle" not in node.value, path
            elif f.endswith((".cu", ".cuh", ".h", ".cpp", ".c")):
                for line in open(path):
                    code = line.split("//")[0]
                    if code.lstrip().startswith("#include"):
                        assert "oracle" not in code, (path, line)
                    assert not re.search(r"\b(dlopen|popen|system|execv\w*|fork)\s*\(", code), (path, line)
    for line in open(os.path.join(ROOT, "include", "node4j.h")):
        assert not (line.lstrip().startswith("#include") and "oracle" in line)


def test_solver_config_validation_and_equality():
    with pytest.raises(ValueError):                       # SolverConfig.java:22-24
        N.SolverConfig(1e-3, 1e-3, 1.0, 1.0)
    a = N.DormandPrince54Solver(N.SolverConfig(1e-3, 1e-4, 1e-10, 10))
    b = a.clone()
    assert a == b and a is not b and a.config is not b.config
    assert N.DormandPrince54Solver().config == N.SolverConfig(1e-3, 1e-3, 1e-10, 100)   # DormandPrince54Solver.java:31-33
    a.addListeners(N.StepCounter())
    assert a.clone().listeners == []                      # listeners are not cloned (:46-52)


def test_conf_json_round_trip():
    # AbstractConfTest.java:21-43 / OdeVertexTest.java:56-120: serialise, deserialise, equal
    for w in (W.mnist_block(batch=2), W.spiral_block(batch=2), W.golden_linear()):
        d = w.conf.to_json()
        back = N.OdeVertex.from_json(d)
        assert back == w.conf
        assert back.to_json() == d
        assert back.clone() == w.conf
    d = W.golden_linear().conf.to_json()
    assert d["odeForwardConf"]["@class"] == "ode.vertex.conf.helper.InputStep"
    assert d["odeForwardConf"]["solverConf"]["@class"] == "ode.solve.conf.DormandPrince54Solver"
    assert set(d["odeForwardConf"]["solverConf"]["config"]) == {"absoluteTolerance", "relativeTolerance", "minStep", "maxStep"}


def test_builder_rejects_what_is_outside_the_native_set():
    b = N.OdeVertex.Builder(None, "a", N.DenseLayer(4, 4))
    with pytest.raises(NotImplementedError):
        b.addLayer("b", N.DenseLayer(4, 4), "nonexistent")
    with pytest.raises(NotImplementedError):
        b.addVertex("v", object(), "a")
    with pytest.raises(NotImplementedError):
        b.addTimeLayer("t", N.DenseLayer(4, 4))
    with pytest.raises(NotImplementedError):
        b.addTimeVertex("tv", N.ShapeMatchVertex(N.MergeVertex()), "a")       # the time merge must be the first vertex
    with pytest.raises(NotImplementedError):
        N.ShapeMatchVertex(object())                                           # only the MergeVertex is wrapped natively
    with pytest.raises(NotImplementedError):
        N.ShapeMatchVertex(N.MergeVertex(), {2: 3})
    with pytest.raises(NotImplementedError):
        N.OdeVertex.Builder(None, "v", N.ShapeMatchVertex(N.MergeVertex()), False)   # vertex without the time input
    with pytest.raises(NotImplementedError):
        N.Convolution2D(8, kernelSize=(5, 5))
    with pytest.raises(NotImplementedError):
        N.Convolution2D(8, hasBias=True)
    with pytest.raises(NotImplementedError):
        N.DenseLayer(3, 3, activation="softmax").desc()


def test_time_vertex_builder_mirrors_reference():
    # OdeVertexTest.fitWithFirstTimeVertex (OdeVertexTest.java:226-254) / examples OdeBlockWithTime, OdeBlockTime
    v = (N.OdeVertex.Builder(None, "ode0", N.ShapeMatchVertex(N.MergeVertex()), True)
         .addLayer("ode1", N.DenseLayer(8), "ode0").build())
    assert [l.kind for l in v.layers] == [_lib.LAYER_TIME_CONCAT, _lib.LAYER_DENSE]
    assert v.layers[0].nOut == 1                                  # ShapeMatchVertex(MergeVertex): overrideSizeDims {1: 1}
    assert v.numParams((5, 8)) == 9 * 8 + 8                       # the dense layer sees nIn + 1 inputs
    assert N.ShapeMatchVertex(N.MergeVertex(), {1: 10}).layer().nOut == 10    # cifar10/OdeBlockTime.java:38
    for w in (W.anode_time_block(batch=2), W.conv_time_block(batch=2)):
        back = N.OdeVertex.from_json(w.conf.to_json())
        assert back == w.conf and w.conf.numParams(w.shape) == w.params.size


def test_default_ode_conf_matches_reference_default():
    v = N.OdeVertex.Builder(None, "a", N.DenseLayer(4, 4)).build()   # OdeVertex.java:168-169
    assert isinstance(v.odeForwardConf, N.FixedStep)
    np.testing.assert_array_equal(v.odeForwardConf.time, [0, 1])
    assert v.odeForwardConf.interpolateForwardIfMultiStep is True


def test_num_params_matches_layout():
    w = W.mnist_block(batch=2)
    assert w.conf.numParams(w.shape) == 2 * 36864 + 3 * 256 == w.params.size          # 74 496
    s = W.spiral_block(batch=2)
    assert s.conf.numParams(s.shape) == 4 * 20 + 20 + 20 * 20 + 20 + 20 * 4 + 4 == s.params.size   # 604


def test_step_counter_and_mask_semantics():
    # StepCounter.java:46-63, Mask.java:24-47
    got = []
    c = N.StepCounter(2, "avg: ", consumer=lambda steps, solves: got.append((steps, solves)))
    fwd_only = N.Mask.backward(c)            # "masks backward": passes forward-in-time solves
    for t, nsteps in (([0, 1], 3), ([1, 0], 5), ([0, 2], 4)):
        fwd_only.begin(np.array(t, dtype=np.float32), None)
        for _ in range(nsteps):
            fwd_only.step(None, 0.1, 0.5)
        fwd_only.done()
    assert got == [(7, 2)]


def test_replayed_solver_state_surface():
    """The SolverState the replayed listeners get (ode/solve/impl/util/SolverState.java): time, the DP54 dense-output midpoint coefficients
    (DormandPrince54Solver.java:44-47 — the same numbers the Java shim's ReplayedSolverState returns), and state accessors that raise."""
    import re
    from neuralode4j_b200.conf import _ReplayState
    st = _ReplayState(0.25)
    assert st.time() == 0.25
    mid = st.getInterpolationMidpoints()
    assert len(mid) == 7 and mid[1] == 0.0 and abs(sum(mid) - 0.5) < 1e-12        # the midpoint weights sum to c = 1/2
    java = open(os.path.join(ROOT, "java", "src", "main", "java", "ode", "vertex", "impl", "helper", "nativeb200", "ReplayedSolverState.java")).read()
    nums = re.search(r"new double\[\]\{(.*?)\};", java, flags=re.S).group(1)
    jvals = [eval(x.replace("d", "")) for x in nums.replace("\n", " ").split(",")]
    np.testing.assert_allclose(jvals, mid, rtol=1e-15)
    for call in (st.getCurrentState, lambda: st.getStateDot(0)):
        with pytest.raises(NotImplementedError):
            call()


def test_workload_shapes():
    w = W.mnist_block()
    assert w.shape == (128, 64, 7, 7) and int(np.prod(w.shape)) == 401408
    assert 2 * 401408 + w.params.size == 877312                                          # N_aug of SURVEY.md section 8a
    s = W.spiral_block()
    assert s.time.size == 100 and abs(s.time[1] - 6 * np.pi / 499) < 1e-6


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU leg of the driver's contract) prints ONE JSON line with the agreed keys; it must not
    need a GPU.  One step of the oracle port on the FULL 128-row minibatch (a few seconds on the host cores)."""
    import json
    import subprocess
    import sys
    env = dict(os.environ, OMP_NUM_THREADS="1")        # what torchrun exports: the arm must still use the host's cores
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "samples/s" and line["higher_is_better"] is True
    for k in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["vs_baseline"] is None and line["cpu_baseline"]["kind"] == "port"
    assert line["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert "workload" in line["config"]


def test_bench_reference_arm_under_torchrun_prints_once():
    """Launched the way the driver launches N > 1 (torchrun, one process per rank): rank 0 alone runs the CPU arm and prints the
    line, the other ranks exit 0 without work."""
    import json
    import subprocess
    import sys
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29541",
           os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["config"]["global_batch"] == 256


def test_output_type_inference_mirrors_reference():
    """OdeVertex.getOutputType (conf/OdeVertex.java:151-154) -> FixedStep / InputStep.getOutputType
    (conf/helper/forward/FixedStep.java:55-67, InputStep.java:50-66, OutputTypeAddTimeAsDimension.java:25-53, OutputTypeFromConfig.java:27-35)."""
    from neuralode4j_b200 import (BatchNormalization, Convolution2D, DenseLayer, DormandPrince54Solver, FixedStep, InputStep, InputType,
                                  OdeVertex, SolverConfig)
    solver = DormandPrince54Solver(SolverConfig(1e-3, 1e-3, 1e-10, 10))
    conv = lambda helper: (OdeVertex.Builder(None, "bn", BatchNormalization(6, "relu")).addLayer("c", Convolution2D(6), "bn").odeConf(helper).build())
    dense = lambda helper: (OdeVertex.Builder(None, "d", DenseLayer(5, 5, "tanh")).odeConf(helper).build())
    cnn, ff = InputType.convolutional(9, 9, 6), InputType.feedForward(5)
    # two time points: the input's own type
    assert conv(FixedStep(solver, [0, 1])).getOutputType(0, cnn) == cnn
    assert dense(FixedStep(solver, [0, 1])).getOutputType(0, ff) == ff
    # more: time becomes a dimension — recurrent(size, T) for dense, NDHWC convolutional3D(T, h, w, c) for conv states
    assert dense(FixedStep(solver, np.linspace(0, 1, 7))).getOutputType(0, ff) == InputType.recurrent(5, 7)
    assert conv(FixedStep(solver, np.linspace(0, 1, 4))).getOutputType(0, cnn) == InputType.convolutional3D(4, 9, 9, 6)
    # InputStep: the time is one of the vertex inputs
    assert dense(InputStep(solver, 1)).getOutputType(0, ff, InputType.feedForward(2)) == ff
    assert dense(InputStep(solver, 1)).getOutputType(0, ff, InputType.feedForward(10)) == InputType.recurrent(5, 10)
    assert conv(InputStep(solver, 0)).getOutputType(0, InputType.feedForward(3), cnn) == InputType.convolutional3D(3, 9, 9, 6)
    with pytest.raises(ValueError, match="Time input index"):
        dense(InputStep(solver, 1)).getOutputType(0, ff)
    with pytest.raises(ValueError, match="Time must be 1D"):
        dense(InputStep(solver, 1)).getOutputType(0, ff, InputType.recurrent(3, 4))
    # the block must map its input type onto itself (OutputTypeFromConfig.java:27-35)
    widening = OdeVertex.Builder(None, "d", DenseLayer(7, 5, "tanh")).odeConf(FixedStep(solver, [0, 1])).build()
    with pytest.raises(ValueError, match="must match output shape"):
        widening.getOutputType(0, ff)
    assert FixedStep(solver, [0, 1]).nrofTimeInputs() == 0 and InputStep(solver, 1).nrofTimeInputs() == 1


def test_java_shim_matches_header():
    """java/src/main/java/.../Node4j.java (the Panama binding a neuralODE4j maintainer adds; not compilable here: no JDK) must agree
    with include/node4j.h: every struct layout field by field (name, order, width) and every downcall handle's symbol and arity."""
    import re
    hdr = open(os.path.join(ROOT, "include", "node4j.h")).read()
    hdr_nc = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    java = open(os.path.join(ROOT, "java", "src", "main", "java", "ode", "vertex", "impl", "helper", "nativeb200", "Node4j.java")).read()
    jtype = {"int32_t": "JAVA_INT", "uint32_t": "JAVA_INT", "int64_t": "JAVA_LONG", "float": "JAVA_FLOAT", "double": "JAVA_DOUBLE"}
    pairs = {"n4j_layer_desc": "LAYER", "n4j_graph_desc": "GRAPH", "n4j_solver_cfg": "CFG", "n4j_stats": "STATS", "n4j_step_rec": "STEP_REC"}
    for cname, jname in pairs.items():
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), hdr_nc, flags=re.S).group(1)
        want = []
        for decl in [d.strip() for d in body.split(";") if d.strip()]:
            m = re.match(r"(const\s+)?(\w+)\s*(\*?)\s*(.+)$", decl)
            ctype, ptr, names = m.group(2), m.group(3), m.group(4)
            for nm in [x.strip() for x in names.split(",")]:
                arr = re.match(r"(\w+)\[(\d+)\]", nm)
                if ptr or nm.startswith("*"):
                    want.append((nm.lstrip("* "), "ADDRESS"))
                elif arr:
                    want.append((arr.group(1), "SEQ%s:%s" % (arr.group(2), jtype[ctype])))
                else:
                    want.append((nm, jtype[ctype]))
        jbody = re.search(r"StructLayout %s = MemoryLayout\.structLayout\((.*?)\);" % jname, java, flags=re.S).group(1)
        got = []
        for m in re.finditer(r"(MemoryLayout\.sequenceLayout\((\d+), (\w+)\)|\b(JAVA_\w+|ADDRESS))\.withName\(\"(\w+)\"\)", jbody):
            got.append((m.group(5), "SEQ%s:%s" % (m.group(2), m.group(3)) if m.group(2) else m.group(4)))
        assert got == want, (cname, got, want)
    # downcalls: symbol exists in the header with the same number of parameters
    for m in re.finditer(r"h\(\"(node4j_\w+)\",\s*FunctionDescriptor\.of\(([^;]*?)\)\);", java, flags=re.S):
        sym, desc = m.group(1), [x.strip() for x in m.group(2).split(",") if x.strip()]
        proto = re.search(r"\b%s\s*\(([^)]*)\)\s*;" % sym, hdr_nc, flags=re.S)
        assert proto, f"{sym} is not declared in node4j.h"
        params = [p.strip() for p in proto.group(1).split(",") if p.strip() and p.strip() != "void"]
        assert len(desc) - 1 == len(params), (sym, desc, params)
        for d, p in zip(desc[1:], params):
            if "*" in p:
                assert d == "ADDRESS", (sym, p, d)
            else:
                assert d == jtype[p.split()[0] if p.split()[0] != "const" else p.split()[1]], (sym, p, d)
    assert "ABI_VERSION = %d" % int(re.search(r"#define N4J_ABI_VERSION (\d+)", hdr).group(1)) in java
    # status codes and enums used by the shim carry the header's values
    for name in ("LAYER_DENSE", "LAYER_CONV3X3", "LAYER_BATCHNORM", "LAYER_ACTIVATION", "LAYER_TIME_CONCAT", "ACT_RELU", "ACT_ELU", "ACT_TANH",
                 "ERR_INVALID_ARGUMENT", "ERR_ILLEGAL_STATE", "ERR_UNSUPPORTED", "ERR_CUDA", "ERR_MAX_STEPS", "ERR_COMM",
                 "TIMEGRAD_ZERO", "TIMEGRAD_CALC"):
        hv = int(re.search(r"N4J_%s = (\d+)" % name, hdr_nc).group(1))
        jv = int(re.search(r"\b%s = (\d+)" % name, java).group(1))
        assert hv == jv, name
    # byte offsets the helper writes the structs at (NativeOdeHelper.create) equal the ctypes layout of the same structs
    from neuralode4j_b200 import _lib
    helper = open(os.path.join(ROOT, "java", "src", "main", "java", "ode", "vertex", "impl", "helper", "nativeb200", "NativeOdeHelper.java")).read()
    assert _lib.LayerDesc.has_bias.offset == 24 and _lib.LayerDesc.eps.offset == 28 and "s.set(JAVA_FLOAT, 28, l.eps)" in helper
    assert _lib.GraphDesc.layers.offset == 40 and "g.set(ADDRESS, 40, layers)" in helper
    assert _lib.SolverCfg.order.offset == 56 and _lib.SolverCfg.flags.offset == 60 and _lib.SolverCfg.max_trials.offset == 64
    assert ctypes.sizeof(_lib.StepRec) == 20 and "i * 20 + 16" in helper
    # the network entry points: NET_DESC nests the solver configuration; NativeOdeNet writes it at the ctypes offsets
    net = open(os.path.join(ROOT, "java", "src", "main", "java", "ode", "vertex", "impl", "helper", "nativeb200", "NativeOdeNet.java")).read()
    assert 'CFG.withName("solver")' in java
    assert _lib.NetDesc.solver.offset == 8 and _lib.NetDesc.t0.offset == 80 and _lib.NetDesc.t1.offset == 84 and ctypes.sizeof(_lib.NetDesc) == 88
    assert "d.set(JAVA_DOUBLE, 8, absTol)" in net and "d.set(JAVA_INT, 68, flags)" in net and "d.set(JAVA_LONG, 72, 0L)" in net
    assert "d.set(JAVA_FLOAT, 80, 0f)" in net and "d.set(JAVA_FLOAT, 84, 1f)" in net
    assert _lib.Stats.accepted.offset == 8 and "fwdStats.get(JAVA_LONG, 8)" in net
