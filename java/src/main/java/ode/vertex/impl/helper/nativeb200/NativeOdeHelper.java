package ode.vertex.impl.helper.nativeb200;

import ode.solve.api.FirstOrderSolverConf;
import ode.solve.api.StepListener;
import ode.solve.conf.DormandPrince54Solver;
import ode.solve.conf.SolverConfig;
import ode.vertex.impl.helper.GraphInputOutput;
import ode.vertex.impl.helper.backward.OdeHelperBackward;
import ode.vertex.impl.helper.forward.GraphInput;
import ode.vertex.impl.helper.forward.OdeHelperForward;
import org.deeplearning4j.nn.graph.ComputationGraph;
import org.deeplearning4j.nn.workspace.ArrayType;
import org.deeplearning4j.nn.workspace.LayerWorkspaceMgr;
import org.nd4j.jita.allocator.impl.AtomicAllocator;
import org.nd4j.linalg.api.ndarray.INDArray;
import org.nd4j.linalg.factory.Nd4j;
import org.nd4j.linalg.primitives.Pair;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;
import java.util.Map;
import java.util.WeakHashMap;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_FLOAT;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

/**
 * B200-native replacement for the helper pair of one OdeVertex: implements BOTH
 * {@link OdeHelperForward#solve} (ode/vertex/impl/helper/forward/OdeHelperForward.java:22) and
 * {@link OdeHelperBackward#solve} (ode/vertex/impl/helper/backward/OdeHelperBackward.java:55) on top of
 * {@code node4j_forward} / {@code node4j_backward}: the whole DormandPrince54 solve INCLUDING f runs on the device.
 * <p>
 * Created by the four conf classes' {@code instantiate()} (see java/patches/conf-helper-instantiate.patch):
 * FixedStep / InputStep / FixedStepAdjoint / InputStepAdjoint keep their constructors and JSON shape.
 * The forward and the backward helper of one inner graph share one native handle (registry keyed by the graph), so
 * {@code lastOutput} of the forward pass is already resident on the device when the backward pass starts.
 * <p>
 * NOT COMPILED IN THIS REPOSITORY (no JDK / DL4J in the build image); what it does call for call is what
 * tests/abi_driver.c does in plain C against the same library.
 */
public final class NativeOdeHelper implements OdeHelperForward, OdeHelperBackward {

    /** one native engine per inner graph and input shape */
    private static final class Handle {
        MemorySegment ptr;
        long[] shape;
        long numParams;
    }

    private static final Map<ComputationGraph, Handle> HANDLES = new WeakHashMap<>();

    private final FirstOrderSolverConf solverConf;
    private final INDArray fixedTime;        // FixedStep: the configured time; InputStep: null
    private final int timeInputIndex;        // InputStep: index of the time input; FixedStep: -1
    private final boolean interpolateIfMultiStep;
    private final boolean needTimeGradient;
    private final StepListener[] listeners;

    /**
     * True when the conf classes should hand out this helper: the library loads, a CUDA device is visible, the solver is
     * a DormandPrince54Solver and the property {@code node4j.disable} is not set.  Whether the INNER GRAPH is inside the
     * native set is only known at the first solve; outside it the helper throws (no silent CPU fallback once selected).
     */
    public static boolean supports(FirstOrderSolverConf solverConf) {
        return !Boolean.getBoolean("node4j.disable") && solverConf instanceof DormandPrince54Solver && Node4j.deviceAvailable();
    }

    public NativeOdeHelper(FirstOrderSolverConf solverConf, INDArray fixedTime, int timeInputIndex, boolean interpolateIfMultiStep,
                           boolean needTimeGradient, StepListener... listeners) {
        this.solverConf = solverConf;
        this.fixedTime = fixedTime;
        this.timeInputIndex = timeInputIndex;
        this.interpolateIfMultiStep = interpolateIfMultiStep;
        this.needTimeGradient = needTimeGradient;
        this.listeners = listeners;
    }

    // ------------------------------------------------------------------------------------------------------------
    // forward: replaces forward/FixedStep.solve :20-29 and forward/InputStep.solve :29-36 (-> SingleStep / MultiStep)
    // ------------------------------------------------------------------------------------------------------------
    @Override
    public INDArray solve(ComputationGraph graph, LayerWorkspaceMgr wsMgr, GraphInput input) {
        GraphInput in = input;
        INDArray time = fixedTime;
        if (timeInputIndex >= 0) {
            final Pair<? extends GraphInput, INDArray> r = input.removeInput(timeInputIndex);     // InputStep.java:30-31
            in = r.getFirst();
            time = r.getSecond();
        }
        final INDArray y0 = in.y0();
        final int nt = (int) time.length();
        final Handle h = handleFor(graph, y0.shape());
        // SingleStep.java:42 / MultiStep.java:42,49-60: caller-allocated output, [B,F,T] or [B,T,C,H,W] for multi-step
        final INDArray out = wsMgr.createUninitialized(ArrayType.ACTIVATIONS, outputShape(y0.shape(), nt), 'c');
        try (Arena a = Arena.ofConfined()) {
            final MemorySegment stats = a.allocate(Node4j.STATS);
            Node4j.check((int) Node4j.BIND_PARAMS.invokeExact(h.ptr, address(graph.params()), h.numParams));
            Node4j.check((int) Node4j.FORWARD.invokeExact(h.ptr, address(y0), address(time.castTo(y0.dataType())), nt, interpolateIfMultiStep ? 1 : 0,
                    address(out), stats));
            // nfe counter: ForwardPass.java:42 increments the inner graph's iteration count once per RHS evaluation
            graph.getConfiguration().setIterationCount((int) stats.get(JAVA_LONG, 0));
            replayListeners(a, h, time, true, interpolateIfMultiStep);
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new IllegalStateException(t);
        }
        return out;
    }

    // ------------------------------------------------------------------------------------------------------------
    // backward: replaces FixedStepAdjoint.solve :19-23 and InputStepAdjoint.solve :29-53 (-> SingleStepAdjoint / MultiStepAdjoint)
    // ------------------------------------------------------------------------------------------------------------
    @Override
    public INDArray[] solve(ComputationGraph graph, InputArrays input, MiscPar miscPars) {
        GraphInputOutput gio = input.getGraphInputOutput();
        INDArray time = fixedTime;
        if (timeInputIndex >= 0) {
            final Pair<GraphInputOutput, INDArray> r = gio.removeInput(timeInputIndex);           // InputStepAdjoint.java:31
            gio = r.getFirst();
            time = r.getSecond();
        }
        final INDArray y0 = gio.y0();
        final int nt = (int) time.length();
        final Handle h = handleFor(graph, y0.shape());
        final int mode = timeInputIndex < 0 ? Node4j.TIMEGRAD_NONE : (needTimeGradient ? Node4j.TIMEGRAD_CALC : Node4j.TIMEGRAD_ZERO);
        final INDArray dLdy0 = miscPars.getWsMgr().createUninitialized(ArrayType.ACTIVATION_GRAD, y0.shape(), 'c');
        final INDArray dLdt = Nd4j.zeros(y0.dataType(), 1, nt);
        // The flat gradient view DL4J handed to the vertex (OdeVertex.java:149).  The theta-adjoint is SEEDED from its current
        // content (SingleStepAdjoint.java:59-60) and the result written back into the real-gradient slots (:85); BatchNorm
        // mean / var slots are left untouched (GradientViewSelectionFromBlacklisted.java:33-37) — all inside node4j_backward.
        final INDArray gradView = graph.getFlattenedGradients();
        try (Arena a = Arena.ofConfined()) {
            final MemorySegment stats = a.allocate(Node4j.STATS);
            Node4j.check((int) Node4j.BIND_PARAMS.invokeExact(h.ptr, address(graph.params()), h.numParams));
            Node4j.check((int) Node4j.BACKWARD.invokeExact(h.ptr, address(input.getLastOutput()), address(input.getLossGradient()),
                    address(time.castTo(y0.dataType())), nt, mode, address(gradView), address(dLdy0), address(dLdt), stats));
            graph.getConfiguration().setIterationCount((int) stats.get(JAVA_LONG, 0));
            replayListeners(a, h, time, false, false);
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new IllegalStateException(t);
        }
        if (timeInputIndex < 0) {
            return new INDArray[]{dLdy0};                                                          // NoTimeGrad.java:26-28
        }
        final INDArray[] epsilons = new INDArray[2];                                               // TimeGrad.createLossGradient: time at its input index
        epsilons[timeInputIndex] = dLdt;
        epsilons[(timeInputIndex + 1) % 2] = dLdy0;
        return epsilons;
    }

    // ------------------------------------------------------------------------------------------------------------
    private Handle handleFor(ComputationGraph graph, long[] shape) {
        synchronized (HANDLES) {
            Handle h = HANDLES.get(graph);
            if (h != null && java.util.Arrays.equals(h.shape, shape)) {
                return h;
            }
            if (h != null) {
                destroy(h);
            }
            h = create(graph, shape);
            HANDLES.put(graph, h);
            return h;
        }
    }

    private Handle create(ComputationGraph graph, long[] shape) {
        if (shape.length != 2 && shape.length != 4) {
            throw new UnsupportedOperationException("Rank not supported: " + shape.length);        // MultiStep.java:57-59
        }
        final NativeGraphDescription desc = new NativeGraphDescription(graph, shape);
        final SolverConfig sc = ((DormandPrince54Solver) solverConf).getConfig();
        final Handle h = new Handle();
        try (Arena a = Arena.ofConfined()) {
            final int n = desc.layers.size();
            final MemorySegment layers = a.allocate(Node4j.LAYER, n);
            for (int i = 0; i < n; i++) {
                final NativeGraphDescription.LayerDesc l = desc.layers.get(i);
                final MemorySegment s = layers.asSlice(i * Node4j.LAYER.byteSize(), Node4j.LAYER.byteSize());
                s.set(JAVA_INT, 0, l.kind);
                s.set(JAVA_INT, 4, l.activation);
                s.set(JAVA_LONG, 8, l.nIn);
                s.set(JAVA_LONG, 16, l.nOut);
                s.set(JAVA_INT, 24, l.hasBias ? 1 : 0);
                s.set(JAVA_FLOAT, 28, l.eps);
            }
            final MemorySegment g = a.allocate(Node4j.GRAPH);
            g.set(JAVA_INT, 0, n);
            g.set(JAVA_INT, 4, shape.length);
            for (int i = 0; i < shape.length; i++) {
                g.set(JAVA_LONG, 8 + 8L * i, shape[i]);
            }
            g.set(ADDRESS, 40, layers);
            final MemorySegment cfg = a.allocate(Node4j.CFG);
            cfg.set(java.lang.foreign.ValueLayout.JAVA_DOUBLE, 0, sc.getAbsTol());
            cfg.set(java.lang.foreign.ValueLayout.JAVA_DOUBLE, 8, sc.getRelTol());
            cfg.set(java.lang.foreign.ValueLayout.JAVA_DOUBLE, 16, sc.getMinStep());
            cfg.set(java.lang.foreign.ValueLayout.JAVA_DOUBLE, 24, sc.getMaxStep());
            cfg.set(java.lang.foreign.ValueLayout.JAVA_DOUBLE, 32, 0.9);      // AdaptiveRungeKuttaStepPolicy.java:43-45
            cfg.set(java.lang.foreign.ValueLayout.JAVA_DOUBLE, 40, 10.0);
            cfg.set(java.lang.foreign.ValueLayout.JAVA_DOUBLE, 48, 0.2);
            cfg.set(JAVA_INT, 56, 5);                                          // DormandPrince54Solver.java:100
            cfg.set(JAVA_INT, 60, Integer.getInteger("node4j.flags", 0));
            cfg.set(JAVA_LONG, 64, 0L);
            final MemorySegment out = a.allocate(ADDRESS);
            Node4j.check((int) Node4j.CREATE.invokeExact(g, cfg, Nd4j.getAffinityManager().getDeviceForCurrentThread(), out));
            h.ptr = out.get(ADDRESS, 0);
            h.shape = shape.clone();
            h.numParams = (long) Node4j.NUM_PARAMS.invokeExact(h.ptr);
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new IllegalStateException(t);
        }
        if (h.numParams != graph.numParams()) {
            destroy(h);
            throw new IllegalStateException("native ODE helper: parameter count " + h.numParams + " != " + graph.numParams() + " of the inner graph");
        }
        return h;
    }

    private static void destroy(Handle h) {
        try {
            Node4j.check((int) Node4j.DESTROY.invokeExact(h.ptr));
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new IllegalStateException(t);
        }
    }

    /** StepListener.begin / step / done (ode/solve/api/StepListener.java:17-32) replayed from the device-side trial log. */
    private void replayListeners(Arena a, Handle h, INDArray time, boolean forward, boolean oneSolve) throws Throwable {
        if (listeners == null || listeners.length == 0) {
            return;
        }
        final long n = (long) Node4j.GET_STEP_LOG.invokeExact(h.ptr, MemorySegment.NULL, 0L);
        final MemorySegment recs = a.allocate(Node4j.STEP_REC, Math.max(n, 1));
        Node4j.GET_STEP_LOG.invokeExact(h.ptr, recs, n);
        final int nt = (int) time.length();
        final int solves = (nt == 2 || oneSolve) ? 1 : nt - 1;
        long i = 0;
        for (int k = 0; k < solves; k++) {
            final INDArray pair = forward
                    ? (solves == 1 ? Nd4j.create(new float[]{time.getFloat(0), time.getFloat(nt - 1)}) : time.get(org.nd4j.linalg.indexing.NDArrayIndex.interval(k, k + 2)))
                    : Nd4j.create(new float[]{time.getFloat(nt - 1 - k), time.getFloat(nt - 2 - k)});
            for (StepListener l : listeners) {
                l.begin(pair, null);
            }
            while (i < n && recs.get(JAVA_INT, i * 20 + 16) == k && recs.get(JAVA_FLOAT, i * 20 + 4) != 0f) {
                if (recs.get(JAVA_INT, i * 20 + 12) != 0) {
                    final float t = recs.get(JAVA_FLOAT, i * 20), step = recs.get(JAVA_FLOAT, i * 20 + 4), err = recs.get(JAVA_FLOAT, i * 20 + 8);
                    for (StepListener l : listeners) {
                        l.step(new ReplayedSolverState(t), Nd4j.scalar(step), Nd4j.scalar(err));
                    }
                }
                i++;
            }
            for (StepListener l : listeners) {
                l.done();
            }
        }
    }

    /** Host pointer of a CPU-backend array or CUDA device pointer of a GPU-backend one; the engine detects which (node4j.h). */
    private static MemorySegment address(INDArray arr) {
        if (arr == null) {
            return MemorySegment.NULL;
        }
        if (!arr.isView() && arr.ordering() != 'c' && arr.rank() > 1) {
            arr = arr.dup('c');
        }
        final long base = "CUDA".equalsIgnoreCase(Nd4j.getExecutioner().getEnvironmentInformation().getProperty("backend"))
                ? AtomicAllocator.getInstance().getPointer(arr).address()
                : arr.data().addressPointer().address();
        return MemorySegment.ofAddress(base + arr.offset() * arr.data().getElementSize());
    }

    private static long[] outputShape(long[] in, int nt) {
        if (nt == 2) {
            return in;
        }
        if (in.length == 2) {
            return new long[]{in[0], in[1], nt};                      // [B,H,T]     MultiStep.java:49-60
        }
        return new long[]{in[0], nt, in[1], in[2], in[3]};            // [B,T,C,H,W]
    }
}
