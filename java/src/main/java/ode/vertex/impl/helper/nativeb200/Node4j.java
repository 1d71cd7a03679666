package ode.vertex.impl.helper.nativeb200;

import java.lang.foreign.Arena;
import java.lang.foreign.FunctionDescriptor;
import java.lang.foreign.Linker;
import java.lang.foreign.MemoryLayout;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.StructLayout;
import java.lang.foreign.SymbolLookup;
import java.lang.invoke.MethodHandle;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_FLOAT;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

/**
 * Panama (java.lang.foreign, JDK 22+) binding of {@code include/node4j.h} — the C ABI of libnode4j_b200.so.
 * One downcall handle per entry point; struct layouts mirror the POD structs of the header field by field.
 * <p>
 * NOT COMPILED IN THIS REPOSITORY: the build image has no JDK / Maven / DL4J (see java/README.md). The same ABI is
 * exercised by neuralode4j_b200/_lib.py (ctypes) and tests/abi_driver.c (plain C) in the test-suite.
 */
final class Node4j {

    static final int ABI_VERSION = 2;

    // status codes (node4j.h)
    static final int OK = 0, ERR_INVALID_ARGUMENT = 1, ERR_ILLEGAL_STATE = 2, ERR_UNSUPPORTED = 3, ERR_CUDA = 4, ERR_MAX_STEPS = 5, ERR_COMM = 6;
    // layer kinds / activations
    static final int LAYER_DENSE = 1, LAYER_CONV3X3 = 2, LAYER_BATCHNORM = 3, LAYER_ACTIVATION = 4, LAYER_TIME_CONCAT = 5;
    static final int ACT_IDENTITY = 0, ACT_RELU = 1, ACT_ELU = 2, ACT_TANH = 3;
    // time gradient modes
    static final int TIMEGRAD_NONE = 0, TIMEGRAD_ZERO = 1, TIMEGRAD_CALC = 2;

    private static final Linker LINKER = Linker.nativeLinker();
    private static final SymbolLookup LIB = SymbolLookup.libraryLookup(System.getProperty("node4j.library", "libnode4j_b200.so"), Arena.global());

    /** n4j_layer_desc */
    static final StructLayout LAYER = MemoryLayout.structLayout(
            JAVA_INT.withName("kind"), JAVA_INT.withName("activation"),
            JAVA_LONG.withName("n_in"), JAVA_LONG.withName("n_out"),
            JAVA_INT.withName("has_bias"), JAVA_FLOAT.withName("eps"));
    /** n4j_graph_desc */
    static final StructLayout GRAPH = MemoryLayout.structLayout(
            JAVA_INT.withName("n_layers"), JAVA_INT.withName("rank"),
            MemoryLayout.sequenceLayout(4, JAVA_LONG).withName("shape"), ADDRESS.withName("layers"));
    /** n4j_solver_cfg */
    static final StructLayout CFG = MemoryLayout.structLayout(
            JAVA_DOUBLE.withName("abs_tol"), JAVA_DOUBLE.withName("rel_tol"), JAVA_DOUBLE.withName("min_step"),
            JAVA_DOUBLE.withName("max_step"), JAVA_DOUBLE.withName("safety"), JAVA_DOUBLE.withName("max_growth"),
            JAVA_DOUBLE.withName("min_reduction"), JAVA_INT.withName("order"), JAVA_INT.withName("flags"),
            JAVA_LONG.withName("max_trials"));
    /** n4j_stats */
    static final StructLayout STATS = MemoryLayout.structLayout(
            JAVA_LONG.withName("nfe"), JAVA_LONG.withName("accepted"), JAVA_LONG.withName("rejected"), JAVA_LONG.withName("solves"),
            JAVA_FLOAT.withName("last_h"), JAVA_FLOAT.withName("last_err"), JAVA_FLOAT.withName("gpu_ms"), JAVA_INT.withName("gpu_launches"));
    /** n4j_step_rec (ABI version 2: carries the index of the solve the trial belongs to) */
    static final StructLayout STEP_REC = MemoryLayout.structLayout(
            JAVA_FLOAT.withName("t"), JAVA_FLOAT.withName("h"), JAVA_FLOAT.withName("err"), JAVA_INT.withName("accepted"), JAVA_INT.withName("solve"));

    /** n4j_net_desc: the whole MNIST ODE-net behind node4j_net_* (batch, nrofKernels, the block's solver, its FixedStep time) */
    static final StructLayout NET_DESC = MemoryLayout.structLayout(
            JAVA_INT.withName("batch"), JAVA_INT.withName("channels"), CFG.withName("solver"),
            JAVA_FLOAT.withName("t0"), JAVA_FLOAT.withName("t1"));

    private static MethodHandle h(String name, FunctionDescriptor d) {
        return LINKER.downcallHandle(LIB.find(name).orElseThrow(() -> new UnsatisfiedLinkError(name)), d);
    }

    static final MethodHandle ABI = h("node4j_abi_version", FunctionDescriptor.of(JAVA_INT));
    static final MethodHandle LAST_ERROR = h("node4j_last_error", FunctionDescriptor.of(ADDRESS));
    static final MethodHandle DEVICE_COUNT = h("node4j_device_count", FunctionDescriptor.of(JAVA_INT));
    static final MethodHandle CREATE = h("node4j_create", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, ADDRESS));
    static final MethodHandle DESTROY = h("node4j_destroy", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    static final MethodHandle NUM_PARAMS = h("node4j_num_params", FunctionDescriptor.of(JAVA_LONG, ADDRESS));
    static final MethodHandle NUM_REAL_PARAMS = h("node4j_num_real_params", FunctionDescriptor.of(JAVA_LONG, ADDRESS));
    static final MethodHandle BIND_PARAMS = h("node4j_bind_params", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG));
    static final MethodHandle FORWARD = h("node4j_forward",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS, ADDRESS));
    static final MethodHandle BACKWARD = h("node4j_backward",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    static final MethodHandle GET_STEP_LOG = h("node4j_get_step_log", FunctionDescriptor.of(JAVA_LONG, ADDRESS, ADDRESS, JAVA_LONG));
    static final MethodHandle COMM_EXPORT = h("node4j_comm_export", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    static final MethodHandle COMM_INIT = h("node4j_comm_init", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS, JAVA_LONG));
    // the network around the block (SURVEY.md section 8, row f3): what ComputationGraph.fit / output do for examples.mnist.OdeNetModel
    static final MethodHandle NET_CREATE = h("node4j_net_create", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS));
    static final MethodHandle NET_DESTROY = h("node4j_net_destroy", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    static final MethodHandle NET_NUM_PARAMS = h("node4j_net_num_params", FunctionDescriptor.of(JAVA_LONG, ADDRESS));
    static final MethodHandle NET_ODE_SLICE = h("node4j_net_ode_slice", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
    static final MethodHandle NET_TRAIN_STEP = h("node4j_net_train_step",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, JAVA_FLOAT, JAVA_FLOAT, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    static final MethodHandle NET_OUTPUT = h("node4j_net_output", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, JAVA_INT));

    static {
        try {
            final int v = (int) ABI.invokeExact();
            if (v != ABI_VERSION) {
                throw new UnsatisfiedLinkError("libnode4j_b200.so has ABI version " + v + ", this binding needs " + ABI_VERSION);
            }
        } catch (Throwable t) {
            throw new ExceptionInInitializerError(t);
        }
    }

    /** Library present and at least one CUDA device visible: what the conf classes ask before they select the native helper. */
    static boolean deviceAvailable() {
        try {
            return (int) DEVICE_COUNT.invokeExact() > 0;
        } catch (Throwable t) {
            return false;
        }
    }

    /** Maps N4J_ERR_* onto the exception types the reference itself throws for the same conditions (node4j.h). */
    static void check(int rc) {
        if (rc == OK) return;
        String msg;
        try {
            msg = ((MemorySegment) LAST_ERROR.invokeExact()).reinterpret(4096).getString(0);
        } catch (Throwable t) {
            msg = "node4j error " + rc;
        }
        switch (rc) {
            case ERR_INVALID_ARGUMENT:
                throw new IllegalArgumentException(msg);       // AdaptiveRungeKuttaSolver.java:108, SolverConfig.java:22, MultiStepAdjoint.java:43
            case ERR_UNSUPPORTED:
                throw new UnsupportedOperationException(msg);  // MultiStep.java:57-59, GradientViewSelectionFromBlacklisted.java:87
            default:
                throw new IllegalStateException(msg);          // NanWatchSolver.java:22; CUDA / step guard / comm failures
        }
    }

    private Node4j() {
    }
}
