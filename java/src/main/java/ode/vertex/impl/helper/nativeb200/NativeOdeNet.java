package ode.vertex.impl.helper.nativeb200;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_FLOAT;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

/**
 * The whole MNIST ODE-net of {@code examples.mnist.OdeNetModel} (ResBlockStem, OdeVertex, Output; OdeNetModel.java:60-88) on the device:
 * one {@link #fit} is what {@code ComputationGraph.fit(DataSet)} does for that model — forward, LossMCXENT, backward through the adjoint
 * solve, Nesterovs, BatchNorm running statistics — and {@link #output} what {@code ComputationGraph.output} does.
 * <p>
 * {@code params} and {@code updaterState} are the flat vectors DL4J itself keeps ({@code model.params()}, the updater's state view) and
 * {@code ModelSerializer} stores, so a model moves between DL4J and this class without conversion: pass off-heap segments that wrap the
 * INDArrays' host buffers (or device pointers), they are updated in place.
 * <p>
 * NOT COMPILED IN THIS REPOSITORY (no JDK in the build image, see java/README.md); the same entry points are exercised through ctypes by
 * neuralode4j_b200/net.py and tests/test_outer_net.py.
 */
public final class NativeOdeNet implements AutoCloseable {

    private final Arena arena = Arena.ofShared();
    private final MemorySegment handle;
    private final MemorySegment fwdStats, bwdStats, score;
    private final int batch;

    /**
     * @param batch       minibatch size (fixed per instance, like the engine's buffers)
     * @param nrofKernels OdeNetModel's nrofKernels (64 in the reference; only 64 takes the tensor-core path)
     * @param absTol      DormandPrince54Solver(new SolverConfig(absTol, relTol, minStep, maxStep)) of the block (OdeNetModel.java:47)
     * @param flags       N4J_FLAG_* bits (0: default mode)
     */
    public NativeOdeNet(int batch, int nrofKernels, double absTol, double relTol, double minStep, double maxStep, int flags, int device) {
        this.batch = batch;
        final MemorySegment d = arena.allocate(Node4j.NET_DESC);
        d.set(JAVA_INT, 0, batch);
        d.set(JAVA_INT, 4, nrofKernels);
        // n4j_solver_cfg at offset 8
        d.set(JAVA_DOUBLE, 8, absTol); d.set(JAVA_DOUBLE, 16, relTol); d.set(JAVA_DOUBLE, 24, minStep); d.set(JAVA_DOUBLE, 32, maxStep);
        d.set(JAVA_DOUBLE, 40, 0.9); d.set(JAVA_DOUBLE, 48, 10.0); d.set(JAVA_DOUBLE, 56, 0.2);     // safety, max growth, min reduction
        d.set(JAVA_INT, 64, 5); d.set(JAVA_INT, 68, flags); d.set(JAVA_LONG, 72, 0L);
        d.set(JAVA_FLOAT, 80, 0f); d.set(JAVA_FLOAT, 84, 1f);                                        // FixedStep time of the block (OdeNetModel.java:86)
        final MemorySegment out = arena.allocate(ADDRESS);
        try {
            Node4j.check((int) Node4j.NET_CREATE.invokeExact(d, device, out));
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new IllegalStateException(t);
        }
        handle = out.get(ADDRESS, 0);
        fwdStats = arena.allocate(Node4j.STATS);
        bwdStats = arena.allocate(Node4j.STATS);
        score = arena.allocate(JAVA_FLOAT);
    }

    public long numParams() {
        try {
            return (long) Node4j.NET_NUM_PARAMS.invokeExact(handle);
        } catch (Throwable t) {
            throw new IllegalStateException(t);
        }
    }

    /** {offset, length} of the OdeVertex's parameters inside the flat vectors */
    public long[] odeSlice() {
        final MemorySegment o = arena.allocate(JAVA_LONG), n = arena.allocate(JAVA_LONG);
        try {
            Node4j.check((int) Node4j.NET_ODE_SLICE.invokeExact(handle, o, n));
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new IllegalStateException(t);
        }
        return new long[]{o.get(JAVA_LONG, 0), n.get(JAVA_LONG, 0)};
    }

    /**
     * One minibatch of ComputationGraph.fit: images [batch,1,28,28], one-hot labels [batch,10]; params / updaterState updated in place.
     * gradOut (may be MemorySegment.NULL) receives the gradient view as the updater saw it.
     *
     * @return the score (LossMCXENT / batch) before the update
     */
    public float fit(MemorySegment params, MemorySegment updaterState, MemorySegment images, MemorySegment labels, float learningRate,
                     float momentum, MemorySegment gradOut) {
        try {
            Node4j.check((int) Node4j.NET_TRAIN_STEP.invokeExact(handle, params, updaterState, images, labels, learningRate, momentum, score, gradOut,
                    fwdStats, bwdStats));
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new IllegalStateException(t);
        }
        return score.get(JAVA_FLOAT, 0);
    }

    /** ComputationGraph.output: probabilities [batch,10]; train: batch statistics in the outer BatchNorm layers (what fit sees) */
    public void output(MemorySegment params, MemorySegment images, MemorySegment probs, boolean train) {
        try {
            Node4j.check((int) Node4j.NET_OUTPUT.invokeExact(handle, params, images, probs, train ? 1 : 0));
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new IllegalStateException(t);
        }
    }

    /** accepted steps of the block's last forward / adjoint solve (StepListener.step callbacks) */
    public long acceptedForward() {
        return fwdStats.get(JAVA_LONG, 8);
    }

    public long acceptedBackward() {
        return bwdStats.get(JAVA_LONG, 8);
    }

    public int batch() {
        return batch;
    }

    @Override
    public void close() {
        try {
            Node4j.check((int) Node4j.NET_DESTROY.invokeExact(handle));
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new IllegalStateException(t);
        } finally {
            arena.close();
        }
    }
}
