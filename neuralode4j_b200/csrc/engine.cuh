// engine.cuh — host side of the native OdeHelperForward/OdeHelperBackward pair: plans the inner graph, owns the
// device buffers, enqueues RHS evaluations and builds the device-driven solve (CUDA graph with a WHILE node whose
// condition the step controller sets on the GPU).
#pragma once

#include <map>
#include <memory>
#include <string>
#include <vector>

#include "common.cuh"
#include "conv_tc.cuh"
#include "conv_wgrad_tc.cuh"
#include "gemm_tc.cuh"
#include "layer_kernels.cuh"
#include "mlp_small.cuh"
#include "solver_kernels.cuh"

namespace n4j {

struct LayerPlan {
    n4j_layer_desc d;
    long long rows;          // B*H*W or B
    int c_in, c_out;         // channels / features in and out
    int H, W;
    long long in_len, out_len;
    long long p_off, p_len;  // slot in the flat DL4J parameter vector
    long long g_off, g_len;  // slot in the real-gradient vector (== slot inside the theta block of the adjoint state)
    bool exact;              // dense: double-accumulating exact kernels
    float* act = nullptr;    // output of this layer (scratch), unused for the last layer
    float* gin = nullptr;    // gradient w.r.t. this layer's input (scratch), unused for the first layer
    float* act2[2] = {nullptr, nullptr};   // the two buffer sets act / gin point into
    float* gin2[2] = {nullptr, nullptr};
    float* wf = nullptr;     // conv: [tap][ci][co]
    float* wb = nullptr;     // conv: [tap][co][ci]
    // dense layer on the tcgen05 path (gemm_tc.cuh)
    bool tcd = false;
    long long img_rows = 0;          // batch rows padded to a multiple of 128
    float* ximg = nullptr;           // split image of the layer input  [2*nIn/4][img_rows][4]
    float* dimg = nullptr;           // split image of the output delta [2*nOut/4][img_rows][4]
    float* dTimg = nullptr;          // split image of delta^T [2*img_rows/4][nOut padded][4]      (wgrad A operand)
    float* xTimg = nullptr;          // stacked image of x^T   [nIn/128 blocks][img_rows/4][256][4] (wgrad B operand)
    float* wsf = nullptr;            // stacked weight image, forward  (rows = out features, K = in features)
    float* wsb = nullptr;            // stacked weight image, dgrad    (rows = in features,  K = out features)
    CUtensorMap tm_x_k, tm_d_k, tm_dT, tm_xT, tm_wf, tm_wb;
    bool tc = false;         // conv runs on the tcgen05 path
    float* img = nullptr;    // conv: operand image of the layer input (hi/lo, padded, chunk-major)
    float* gimg = nullptr;   // conv: operand image of the gradient w.r.t. the layer output
    float* wimg_f = nullptr; // conv: per-tap weight images, forward
    float* wimg_b = nullptr; // conv: per-tap weight images, dgrad (flipped taps, transposed)
    BnScratch bn{};
    double* bn_pool = nullptr;     // deferred statistics: this layer's four accumulator sets inside Engine::bn_pool_
    bool bn_deferred_ok = false;
    unsigned int* bn_seq = nullptr; // sharded deferred statistics: sequence counters [forward|backward][set 0|1] of this layer (device)
    long long bn_mb = -1;          // ... and the byte offset of its mailboxes inside the statistics area of a comm region
};

struct DevBuf {
    float* p = nullptr;
    long long n = 0;
};

struct SolveGraph {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    int prologue_launches = 0, body_launches = 0, epilogue_launches = 0;
};

class Engine {
public:
    Engine(const n4j_graph_desc& g, const n4j_solver_cfg& cfg, int device);
    ~Engine();

    long long num_params() const { return n_params_; }
    long long num_real() const { return n_real_; }
    long long state_len() const { return nz_; }

    void bind_params(const float* p, long long n);
    void forward(const float* y0, const float* t, int nt, int interpolate, float* out, n4j_stats* st);
    void backward(const float* last_output, const float* dLdz, const float* t, int nt, int tg_mode, float* grad_inout,
                  float* dLdy0, float* dLdt, n4j_stats* st);
    long long get_step_log(n4j_step_rec* buf, long long cap);
    void rhs(const float* z, float t, float* fz);
    void rhs_vjp(const float* z, float t, const float* g, float* fz, float* dz, float* dreal, float* dt);
    void error_norm(const float* K, const float* y0, const float* y1, float step, long long n, float* err_out);
    void bench_solver_pass(long long n, int iters, float* us);
    void bench_piece(int what, int iters, float* us);
    void comm_export(void* handle64);
    void comm_init(int rank, int world, const void* handles, long long global_batch);
    cudaStream_t stream() const { return stream_; }
    // rank-4 states cross forward / backward in the engine's internal NHWC layout instead of the reference's NCHW (callers inside
    // the library whose own tensors are NHWC: outer_net.cuh's tensor-core path)
    void set_nhwc_io(bool on) { nhwc_io_ = on; }

private:
    // ---- plan ------------------------------------------------------------------------------------
    int device_;
    bool nhwc_io_ = false;
    n4j_solver_cfg cfg_;
    bool exact_ = false;       // N4J_FLAG_EXACT_CONTRACTIONS
    bool pdl_ = true;          // programmatic dependent launch on every kernel (N4J_FLAG_NO_PDL turns it off)
    ConvGeom geom_{};          // padded geometry shared by all conv layers of the block
    size_t conv_smem_ = 0;
    std::vector<LayerPlan> L_;
    int rank_;
    long long B_;
    int C_, H_, W_;            // state geometry (rank 4) or C_=F (rank 2), H_=W_=1
    long long nz_, nz_pad_;    // state length, padded to a multiple of 4
    long long n_params_, n_real_, real_pad_;
    long long n_alloc_;        // floats per solver buffer (adjoint layout: z | a | theta | a_t | 0-tail)
    long long off_a_, off_theta_, off_t_;
    long long max_len_;        // largest activation

    // ---- device state ----------------------------------------------------------------------------
    cudaStream_t stream_ = nullptr, stream2_ = nullptr;
    cudaStream_t side_ = nullptr;          // side branch: weight-gradient kernels overlap the dgrad/BatchNorm chain
    static constexpr int EV_POOL = 16;
    cudaEvent_t ev_pool_[EV_POOL] = {};
    int ev_next_ = 0;
    cudaEvent_t slot_event_[2] = {};
    bool slot_event_valid_[2] = {false, false};
    bool side_busy_ = false;
    bool side_capable_ = false;
    size_t wgrad_smem_ = 0;
    int wgrad_nbuf_ = 1;
    WgradTcGeom wgeom_{};            // tensor-core weight gradient (conv_wgrad_tc.cuh)
    size_t wgrad_tc_smem_ = 0;
    bool wgrad_tc_ = false;
    // k_stage_combine_last: the evaluation of stages 1-5 leaves its last elementwise kernel (forward: the last BatchNorm's apply,
    // adjoint: the first BatchNorm's backward dx) to the stage combination that follows
    bool leave_last_ = false;        // request to the evaluation being enqueued
    FuseLast left_{};                // what it left (mode 0: nothing)
    bool prearm_next_ = false;       // the next PRO 1 convolution re-arms the first BatchNorm's statistics set (see ConvTcArgs::prearm)
    bool combine_fusion_ = getenv("N4J_NO_COMBINE_FUSION") == nullptr;
    bool fuse_last_bn_ok_ = false;   // set while enqueueing the forward half of an augmented evaluation
    bool fwd_left_last_bn_ = false;  // ... and the forward half left the last BatchNorm apply to the backward statistics kernel
    cudaEvent_t ev0_ = nullptr, ev1_ = nullptr;
    float* Y_ = nullptr;       // [2][n_alloc]
    float* K_ = nullptr;       // [7][n_alloc]
    Ctrl* ctrl_ = nullptr;
    double* red_acc_ = nullptr;
    // multi-GPU (row-sharded minibatch)
    int world_ = 1, rank_id_ = 0;
    void* comm_region_ = nullptr;
    size_t comm_bytes_ = 0;
    long long comm_xcap_ = 0, comm_xll_off_ = 0, comm_xll_cap_ = 0;
    long long comm_stat_off_ = 0;      // BatchNorm statistics mailboxes: [layer][forward|backward][set][source rank][128 cells of 16 bytes]
    bool mgpu_deferred_ = false;       // sharded path keeps the deferred statistics (and the combine fusion) — N4J_MGPU_DEFERRED=0 switches back to the last-block exchange
    bool vec_allreduce_ll_ = getenv("N4J_NO_LL_VEC_ALLREDUCE") == nullptr;        // A/B switch: pull exchange through xbuf instead
    bool skip_vec_allreduce_ = getenv("N4J_DBG_SKIP_VEC_ALLREDUCE") != nullptr;   // timing experiments only
    void enqueue_vec_allreduce(cudaStream_t s, Ref slot, long long n);
    CommDev* comm_dev_ = nullptr;
    std::vector<void*> peer_ptrs_;
    int comm_grid(long long n) const;   // strided accumulators of the scalar reductions (error norm, initial step)
    Ctrl* ctrl_host_ = nullptr;   // pinned
    n4j_step_rec* log_ = nullptr;
    int log_cap_ = 1 << 16;
    std::vector<n4j_step_rec> log_host_;
    long long log_count_ = 0;
    float* tpair_ = nullptr;      // [max_pairs][2]
    float* tpair_host_ = nullptr; // pinned
    int max_pairs_ = 0;
    float* wanted_ = nullptr;     // requested times (device)
    int wanted_cap_ = 0;
    double* dot_acc_ = nullptr;
    // small dense inner graphs: one launch per evaluation (mlp_small.cuh)
    bool mlp_fused_ = false;
    MlpSmallArgs mlp_args_{};
    size_t mlp_smem_ = 0;
    double* mlp_part_ = nullptr;
    unsigned int* mlp_ticket_ = nullptr;
    void enqueue_mlp_small(cudaStream_t s, Ref in, Ref out, bool aug);
    // deferred BatchNorm statistics (layer_kernels.cuh: BnScratch)
    double* bn_pool_ = nullptr;
    size_t bn_pool_len_ = 0;
    int bn_slot_ = 0;
    bool eval_begun_ = false, in_aug_ = false;
    void set_bn_slot(int slot);
    void begin_eval();
    void bn_rearm_all(cudaStream_t s);
    // time input (N4J_LAYER_TIME_CONCAT)
    bool time_input_ = false;
    bool aug_tslot_ = false;      // the adjoint solve being enqueued carries a time slot (CalcTimeGrad)
    TimeSpec tspec_{0, 0.f, 0.f}; // how the RHS evaluation being enqueued derives its time from the control block
    double* tcat_acc_ = nullptr;
    float* tcat_misc_ = nullptr;  // [0] ticket (as unsigned), [1] last g^T df/dt
    float* tg_ = nullptr;         // [0]=lastTime (CalcMultiStepTimeGrad.java:18), [1]=current d
    DevBuf params_, gradflat_, zt_, gt_, io_, io2_, gb_[3], theta_tmp_, wpart_, dldt_;
    const float* user_params_ = nullptr;
    bool params_fresh_ = false;   // device copy of the parameters is current (set by forward, consumed by backward)
    bool have_last_output_ = false;
    int last_nt_ = 0;
    std::map<long long, SolveGraph> graphs_;
    int launches_ = 0;            // launches enqueued since the counter was reset (capture-time accounting)

    // ---- helpers -------------------------------------------------------------------------------------
    void ensure(DevBuf& b, long long n);
    void ensure_zt(long long n);
    static bool is_device_ptr(const void* p);
    void to_device(float* dst, const float* src, long long n);     // host or device source
    void from_device(float* dst, const float* src, long long n);   // host or device destination
    int grid_for(long long n, int per_thread = 1) const;
    int grid_red(long long n4) const;
    SolverParams solver_params(long long n_norm) const;

    void upload_params();
    void state_in(const float* user, float* internal);             // reference layout -> internal
    void state_out(const float* internal, float* user_dev);        // internal -> reference layout (device buffer)
    void theta_flat_to_internal(const float* flat, float* theta);
    void theta_internal_to_flat(const float* theta, float* flat);
    void theta_internal_to_real(const float* theta, float* real);

    void enqueue_rhs_fwd(cudaStream_t s, Ref in, Ref out, bool need_compact = true, bool bn1_ready = false);
    void enqueue_rhs_aug(cudaStream_t s, Ref in, Ref out, bool bn1_ready = false, int slot = 0, bool defer_join = false, bool last_of_trial = false);
    int first_conv_ = -1;
    void join_side(cudaStream_t s);
    void launch_gemm_tc(cudaStream_t s, const CUtensorMap& ta, const CUtensorMap& tb, const GemmTcArgs& a);
    void launch_conv_tc(cudaStream_t s, const float* img, const float* wimg, Ref out, int act, const LayerPlan* next_bn = nullptr,
                        const ConvTcArgs* extra = nullptr, int pro_kind = 0, int epi_kind = 0);
    struct ConvRec { ConvTcArgs a; int pro, epi; };
    std::vector<ConvRec>* conv_rec_ = nullptr;   // bench_piece: records the conv launches of one evaluation
    void launch_conv_args(cudaStream_t s, const ConvTcArgs& a, int pro_kind, int epi_kind);
    void enqueue_rhs(cudaStream_t s, bool aug, Ref in, Ref out, bool bn1_ready = false) {
        aug ? enqueue_rhs_aug(s, in, out, bn1_ready) : enqueue_rhs_fwd(s, in, out, false, bn1_ready);
    }
    BnFuse bn_fuse() const;
    void enqueue_prologue(cudaStream_t s, bool aug, long long n4, const SolverParams& sp, bool dense_out, int nt);
    void enqueue_trial(cudaStream_t s, bool aug, long long n4, const SolverParams& sp, bool dense_out, int nt,
                       cudaGraphConditionalHandle cond, int use_cond);
    void enqueue_epilogue(cudaStream_t s, bool dense_out, int nt);
    SolveGraph& get_graph(bool aug, int tslot, bool dense_out, int nt);
    void run_solve(bool aug, int tslot, bool dense_out, int nt);
    void reset_ctrl();
    void finish_call(n4j_stats* st, int launches_outside);
    void set_time_pairs(const std::vector<float>& pairs);

    int calls_launches_ = 0;       // launches of the current API call (incl. executed graph bodies)
    std::vector<std::pair<SolveGraph*, int>> pending_graph_runs_;
};

}  // namespace n4j
