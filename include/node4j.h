/* node4j.h — C ABI of the B200-native ODE-block engine (libnode4j_b200.so).
 *
 * Drop-in boundary for neuralODE4j's ODE-block hot path.  The reference (100 % Java, arithmetic in
 * un-vendored ND4J/DL4J) has no FFI of its own; its plugin seam for this path is the pair of helper
 * interfaces installed by OdeVertex.Builder.odeConf/odeForward/odeBackward
 * (src/main/java/ode/vertex/conf/OdeVertex.java:271-296):
 *
 *   INDArray   OdeHelperForward.solve (ComputationGraph, LayerWorkspaceMgr, GraphInput)
 *                  src/main/java/ode/vertex/impl/helper/forward/OdeHelperForward.java:22
 *   INDArray[] OdeHelperBackward.solve(ComputationGraph, InputArrays, MiscPar)
 *                  src/main/java/ode/vertex/impl/helper/backward/OdeHelperBackward.java:55
 *
 * A Java shim (Panama downcalls or JNI) implements those two interfaces on top of
 * the entry points below; FixedStep / InputStep / DormandPrince54Solver / SolverConfig keep their
 * constructors and JSON shape.  Everything here is plain C: pointers, sizes, POD structs, int status.
 *
 * Conventions
 *  - All tensors are fp32.  Every data pointer may be a HOST pointer or a CUDA DEVICE pointer of the
 *    engine's device; the engine detects which (cudaPointerGetAttributes) and stages host buffers
 *    itself.  Outputs are caller-allocated and written in place (SingleStep.java:42, MultiStep.java:42).
 *  - Tensors use the reference's layouts: state c-order [B,F] or NCHW [B,C,H,W]
 *    (NoTimeInput.java:32-42); multi-time outputs [B,F,T] / [B,T,C,H,W] (MultiStep.java:49-60);
 *    parameters and gradients in DL4J's flat view order (dense W [nIn,nOut] f-order then b; conv W
 *    [Cout,Cin,3,3] c-order; batchnorm gamma,beta,mean,var).
 *  - Calls are synchronous and a handle is not thread-safe (one caller thread per vertex, like the
 *    reference's helpers).  Functions return N4J_OK or an error code; node4j_last_error() gives the
 *    message of the calling thread's last failure.  The library never aborts and has no CPU fallback:
 *    without a usable CUDA device every compute entry point fails with N4J_ERR_CUDA.
 */
#ifndef NODE4J_H
#define NODE4J_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define N4J_ABI_VERSION 2

/* ---- status codes; the Java shim maps them onto the reference's exception types ---------------- */
enum {
    N4J_OK = 0,
    N4J_ERR_INVALID_ARGUMENT = 1, /* IllegalArgumentException: t.length()!=2 (AdaptiveRungeKuttaSolver.java:108),
                                     minStep>=maxStep (SolverConfig.java:22), unsorted time (MultiStepAdjoint.java:43) */
    N4J_ERR_ILLEGAL_STATE = 2,    /* IllegalStateException: NaN detected (NanWatchSolver.java:22), backward before forward */
    N4J_ERR_UNSUPPORTED = 3,      /* UnsupportedOperationException: layer/vertex outside the native set, rank not 2/4
                                     (MultiStep.java:57-59) */
    N4J_ERR_CUDA = 4,             /* CUDA runtime failure or no device */
    N4J_ERR_MAX_STEPS = 5,        /* step guard tripped (the reference would loop forever) */
    N4J_ERR_COMM = 6              /* multi-GPU exchange: a peer did not show up within the spin limit, or the IPC setup failed */
};

/* ---- inner graph description (what OdeVertex.Builder assembled; chain of DL4J layers) ---------- */
enum {
    N4J_LAYER_DENSE = 1,
    N4J_LAYER_CONV3X3 = 2,
    N4J_LAYER_BATCHNORM = 3,
    N4J_LAYER_ACTIVATION = 4,
    /* Time-dependent f: the reference's "timeConc" first vertex, ShapeMatchVertex(MergeVertex, {1: n_out}) fed by
     * TimeInput (ode/vertex/conf/OdeVertex.java:251, ode/vertex/impl/ShapeMatchVertex.java:50-79,
     * ode/vertex/impl/helper/TimeInput.java:27-50): n_out channels (features) holding the solver's current time are
     * appended to the state along dimension 1.  Only valid as layer 0.  The adjoint's time slot then carries
     * -a^T df/dt (TimeInput.update), and the reference's stage-time quirk (every stage is evaluated at t+h,
     * FirstOrderEquationWithState.java:45,82) becomes observable; N4J_FLAG_EXACT_STAGE_TIMES selects t+c_s*h instead. */
    N4J_LAYER_TIME_CONCAT = 5
};
enum { N4J_ACT_IDENTITY = 0, N4J_ACT_RELU = 1, N4J_ACT_ELU = 2, N4J_ACT_TANH = 3 };

typedef struct n4j_layer_desc {
    int32_t kind;       /* N4J_LAYER_* */
    int32_t activation; /* N4J_ACT_* (every DL4J layer carries one) */
    int64_t n_in;       /* dense: nIn; conv/batchnorm: channels; 0 = infer from the previous layer */
    int64_t n_out;      /* dense: nOut; conv: kernels; batchnorm: channels */
    int32_t has_bias;   /* dense only; conv3x3 is hasBias(false) (examples/mnist/LayerUtil.java:65-72) */
    float eps;          /* batchnorm epsilon */
} n4j_layer_desc;

typedef struct n4j_graph_desc {
    int32_t n_layers;
    int32_t rank;     /* 2: [B,F]   4: [B,C,H,W] */
    int64_t shape[4]; /* state shape, batch first (LOCAL batch when sharded, see node4j_comm_init) */
    const n4j_layer_desc* layers;
} n4j_graph_desc;

/* ---- solver configuration: SolverConfig.java:17-30 + AdaptiveRungeKuttaStepPolicy.java:43-45 ---- */
enum {
    N4J_FLAG_EXACT_STAGE_TIMES = 1u << 0, /* evaluate stage s at t+c_s*h instead of the reference's t+h */
    N4J_FLAG_NO_GRAPH = 1u << 1,          /* debug: drive the step loop from the host (one sync per trial) instead
                                             of the device-side CUDA-graph while loop */
    N4J_FLAG_FAST_TF32 = 1u << 2,         /* single-pass TF32 tensor-core contractions (default: 3xTF32, fp32 parity) */
    N4J_FLAG_NO_TENSOR = 1u << 4,         /* keep convolutions on the fp32 SIMT kernels (used to cross-check the tcgen05 path) */
    N4J_FLAG_FORCE_TENSOR = 1u << 6,      /* take the tcgen05 dense path even for layers below the size threshold (tests) */
    N4J_FLAG_NO_PROLOGUE_FUSION = 1u << 7, /* run BatchNorm apply / backward dx as separate kernels instead of inside the conv operand prologue */
    N4J_FLAG_NO_MLP_FUSION = 1u << 10,     /* small dense inner graphs layer by layer instead of one fused launch per evaluation */
    N4J_FLAG_NO_DEFERRED_STATS = 1u << 9,  /* BatchNorm statistics finalised by the last block of the producing kernel (always so on multi-GPU) instead of by the consumers */
    N4J_FLAG_NO_BWD_STATS_FUSION = 1u << 8, /* reduce BatchNorm backward statistics in their own kernel instead of the dgrad conv epilogue */
    N4J_FLAG_SIMT_WGRAD = 1u << 11,        /* conv weight gradients on the fp32 FMA kernel (k_conv_wgrad64) instead of the tcgen05 one (A/B, cross-check) */
    N4J_FLAG_NO_PDL = 1u << 5,            /* launch kernels without programmatic dependent launch (debug / A-B timing) */
    N4J_FLAG_EXACT_CONTRACTIONS = 1u << 3 /* parity mode: every dense/conv contraction accumulates exact fp32 products in
                                             double and rounds once (order-independent, bit-comparable with the CPU oracle);
                                             CUDA-core kernels on the FP64 pipe: about 7x the default mode's time on the MNIST block */
};

typedef struct n4j_solver_cfg {
    double abs_tol, rel_tol, min_step, max_step;
    double safety, max_growth, min_reduction; /* 0.9, 10, 0.2 */
    int32_t order;                            /* 5 (DormandPrince54Solver.java:100) */
    uint32_t flags;                           /* N4J_FLAG_* */
    int64_t max_trials;                       /* per solve; 0 = default guard (100000) */
} n4j_solver_cfg;

/* ---- observable side channels: nfe (ForwardPass.java:42), accepted steps (StepListener.step) ---- */
typedef struct n4j_stats {
    int64_t nfe;      /* RHS evaluations as the reference counts them: 3 + 6*trials per solve (+1 per time-grad dot) */
    int64_t accepted; /* accepted steps == StepListener.step callbacks */
    int64_t rejected;
    int64_t solves;   /* integrate() calls (segments) */
    float last_h, last_err;
    float gpu_ms;      /* device time of the call (CUDA events on the engine's stream) */
    int32_t gpu_launches; /* kernels this call launched (graph bodies counted per executed iteration) */
} n4j_stats;

typedef struct n4j_step_rec { /* one entry per TRIAL step, in order */
    float t;  /* solver time after the trial (advanced only if accepted) */
    float h;  /* step size that was tried */
    float err;
    int32_t accepted;
    int32_t solve; /* which integrate() call (segment) of this API call the trial belongs to, 0-based: listener begin/done boundaries */
} n4j_step_rec;

enum { N4J_TIMEGRAD_NONE = 0, /* FixedStep: NoTimeGrad.java:26-28 */
       N4J_TIMEGRAD_ZERO = 1, /* InputStep(needTimeGradient=false): ZeroTimeGrad.java:45-55 */
       N4J_TIMEGRAD_CALC = 2  /* InputStep(needTimeGradient=true):  CalcTimeGrad.java:52-74 */ };

typedef struct n4j_handle n4j_handle;

/* Library / device probes (no compute). */
int32_t node4j_abi_version(void);
const char* node4j_last_error(void);
int32_t node4j_device_count(void);

/* Replaces OdeHelperForward.instantiate()/OdeHelperBackward.instantiate()
 * (conf/helper/forward/OdeHelperForward.java:17, conf/helper/backward/OdeHelperBackward.java:17): builds the
 * native helper pair for one OdeVertex on CUDA device `device`. */
int32_t node4j_create(const n4j_graph_desc* graph, const n4j_solver_cfg* cfg, int32_t device, n4j_handle** out);
int32_t node4j_destroy(n4j_handle* h);

int64_t node4j_num_params(const n4j_handle* h);      /* ComputationGraph.numParams() of the inner graph */
int64_t node4j_num_real_params(const n4j_handle* h); /* realGradientView length (GradientViewSelectionFromBlacklisted.java:50-61) */
int64_t node4j_state_len(const n4j_handle* h);

/* Binds the caller-owned flat parameter buffer shared with DL4J (OdeVertex.java:67-85, :124-129).  The buffer
 * is re-read at every forward/backward call (the updater changes it between iterations). */
int32_t node4j_bind_params(n4j_handle* h, const float* params, int64_t n_params);

/* Replaces OdeHelperForward.solve (impl/helper/forward/FixedStep.java:20-29, InputStep.java:29-36):
 *   nt == 2: SingleStep.solve  — out has the state's shape.
 *   nt  > 2: MultiStep.solve   — out is [B,F,nt] or [B,nt,C,H,W] and includes z(t0); `interpolate` selects
 *            InterpolatingMultiStepSolver (one solve + dense output) vs SingleSteppingMultiStepSolver.
 * The result is also kept on the device as `lastOutput` for node4j_backward (OdeGraphHelper.java:129). */
int32_t node4j_forward(n4j_handle* h, const float* y0, const float* t, int32_t nt, int32_t interpolate,
                       float* out, n4j_stats* stats);

/* Replaces OdeHelperBackward.solve (impl/helper/backward/FixedStepAdjoint.java:19-23, InputStepAdjoint.java:29-53,
 * SingleStepAdjoint.java:37-88, MultiStepAdjoint.java:49-90).
 *   last_output: z(t1) / all z(t_k) in node4j_forward's output layout, or NULL to use the device-resident copy
 *                of the previous node4j_forward on this handle.
 *   dL_dz:       loss gradient (epsilon) in the same layout as the forward output.
 *   grad_inout:  flat gradient view [num_params]; the theta-adjoint is SEEDED from its current content
 *                (SingleStepAdjoint.java:59-60) and the result is written back into the real-gradient slots
 *                (:85); batchnorm mean/var slots are left untouched.
 *   dL_dy0:      [state shape] epsilon for the vertex input.
 *   dL_dt:       [nt] time gradient for N4J_TIMEGRAD_ZERO/CALC, may be NULL for N4J_TIMEGRAD_NONE. */
int32_t node4j_backward(n4j_handle* h, const float* last_output, const float* dL_dz, const float* t, int32_t nt,
                        int32_t time_grad_mode, float* grad_inout, float* dL_dy0, float* dL_dt, n4j_stats* stats);

/* Replay of StepListener.step (solve/api/StepListener.java:26) for the last forward or backward call: copies up
 * to `cap` trial records and returns the total number recorded. */
int64_t node4j_get_step_log(n4j_handle* h, n4j_step_rec* buf, int64_t cap);

/* f(z,t) and its vector-Jacobian products on their own (ForwardPass.calculateDerivative, ForwardPass.java:41-48;
 * BackpropagateAdjoint.backPropagate, BackpropagateAdjoint.java:87-143).  dreal has num_real_params entries in
 * realGradientView order. */
int32_t node4j_rhs(n4j_handle* h, const float* z, float t, float* fz);
int32_t node4j_rhs_vjp(n4j_handle* h, const float* z, float t, const float* g, float* fz, float* dz, float* dreal);
/* Same, and also g^T df/dt, the epsilon of the time input (TimeInput.java:46-50); 0 for graphs without a time vertex. */
int32_t node4j_rhs_vjp_time(n4j_handle* h, const float* z, float t, const float* g, float* fz, float* dz, float* dreal, float* dt);

/* DormandPrince54Mse.estimateMse (solve/impl/DormandPrince54Solver.java:78-93) on caller data: K is [7,n]
 * row-major.  Exposed for the unit pins of the solver kernels and for bandwidth measurements. */
int32_t node4j_error_norm(n4j_handle* h, const float* K, const float* y0, const float* y1, float step, int64_t n,
                          float* err_out);

/* Multi-GPU (new work; the reference is single-device): the minibatch is sharded by rows, one process per GPU of one
 * NVLink/NVSwitch box.  The couplings of the path (error-norm scalar, BatchNorm statistics, parameter gradients) are
 * exchanged by one-shot all-reduces over peer-mapped memory written into the kernels themselves, so there is no NCCL
 * call on the data path.  Setup: every rank calls node4j_comm_export (64-byte cudaIpcMemHandle_t of its exchange region),
 * the host framework all-gathers the handles (rank order), every rank calls node4j_comm_init with all of them.
 * Afterwards the step controller, BatchNorm statistics and the parameter gradients are global: `graph.shape[0]` is the
 * LOCAL batch, global_batch = world * local batch. */
int32_t node4j_comm_export(n4j_handle* h, void* handle64);
int32_t node4j_comm_init(n4j_handle* h, int32_t rank, int32_t world, const void* all_handles, int64_t global_batch);

/* Microbenchmark hook used by bench.py for the roofline figure: runs `iters` back-to-back trial-step solver passes
 * (6 stage combinations + error norm, no RHS) over a state of n floats and returns the average device time of one
 * pass in microseconds. */
int32_t node4j_bench_solver_pass(n4j_handle* h, int64_t n, int32_t iters, float* us_per_pass);

/* Same idea for pieces of the RHS on the handle's own shapes: average device time (CUDA events on the engine's
 * stream) of one f(z,t) evaluation, one augmented evaluation, or one launch of a conv kernel. */
enum {
    N4J_BENCH_RHS_FWD = 1,
    N4J_BENCH_RHS_AUG = 2,
    N4J_BENCH_CONV_FWD = 3,         /* bare conv kernel on a prepared operand image */
    N4J_BENCH_CONV_DGRAD = 4,
    N4J_BENCH_CONV_WGRAD = 5,
    N4J_BENCH_CONV_FWD_FUSED = 6,   /* the forward conv launch exactly as an evaluation issues it: BatchNorm-apply prologue + statistics epilogue */
    N4J_BENCH_CONV_DGRAD_FUSED = 7  /* the dgrad launch as issued: BatchNorm-backward prologue + backward-statistics epilogue */
};
int32_t node4j_bench_piece(n4j_handle* h, int32_t what, int32_t iters, float* us_per_iter);

/* The CUDA stream (cudaStream_t) all work of this handle is enqueued on, so that a host framework can record
 * events around calls or order its own work after them. */
int32_t node4j_get_stream(n4j_handle* h, void** stream);

/* ---- the network around the ODE block -------------------------------------------------------------
 * One full training step of the reference's MNIST ODE-net (examples/mnist/OdeNetModel.java:60-88) on the device: ResBlockStem
 * (examples/mnist/ResBlockStem.java:25-48), the OdeVertex above, Output (examples/mnist/Output.java:27-35: BN-ReLU, global average
 * pooling, Dense(10) softmax + LossMCXENT), gradients divided by the minibatch size, Nesterovs(momentum) and the BatchNorm
 * running-statistics update — what ComputationGraph.fit(DataSet) does for that model.  `params` / `updater_state` are the model's
 * flat vectors in DL4J order (what ModelSerializer stores as coefficients.bin / updaterState.bin; see neuralode4j_b200/dl4j_zip.py),
 * host or device pointers, updated in place.  images [B,1,28,28], labels one-hot [B,10]. */
typedef struct n4j_net n4j_net;
typedef struct n4j_net_desc {
    int32_t batch;
    int32_t channels;      /* nrofKernels: 64 in the reference (examples/mnist/OdeNetModel.java:40) */
    n4j_solver_cfg solver; /* of the ODE block */
    float t0, t1;          /* FixedStep time of the block: 0, 1 (OdeNetModel.java:86) */
} n4j_net_desc;
int32_t node4j_net_create(const n4j_net_desc* desc, int32_t device, n4j_net** out);
int32_t node4j_net_destroy(n4j_net* net);
int64_t node4j_net_num_params(const n4j_net* net);
/* offset and length of the ODE block's slice inside the flat vectors */
int32_t node4j_net_ode_slice(const n4j_net* net, int64_t* offset, int64_t* length);
/* score = LossMCXENT / B of the minibatch BEFORE the update; grad_out (optional, [num_params]) receives the gradient view as the
 * updater saw it (sums over the minibatch; BatchNorm mean/var slots = (running - batch) * (1 - decay)). */
int32_t node4j_net_train_step(n4j_net* net, float* params, float* updater_state, const float* images, const float* labels, float lr,
                              float momentum, float* score, float* grad_out, n4j_stats* ode_forward, n4j_stats* ode_backward);
/* ComputationGraph.output: probabilities [B,10].  train != 0 uses batch statistics in the outer BatchNorm layers (what fit sees),
 * train == 0 their running statistics; the ODE block always runs in training mode (SingleStep.java:37). */
int32_t node4j_net_output(n4j_net* net, const float* params, const float* images, float* probs, int32_t train);

#ifdef __cplusplus
}
#endif
#endif /* NODE4J_H */
