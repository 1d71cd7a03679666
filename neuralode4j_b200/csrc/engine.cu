// engine.cu — implementation of the native ODE-block engine and its C ABI (include/node4j.h).
//
// Reference call stacks replaced here (all under /root/reference/src/main/java/ode):
//   vertex/impl/helper/forward/{SingleStep,MultiStep,FixedStep,InputStep}.java     -> Engine::forward
//   vertex/impl/helper/backward/{SingleStepAdjoint,MultiStepAdjoint,...}.java      -> Engine::backward
//   vertex/impl/helper/forward/ForwardPass.java:41-102                             -> Engine::enqueue_rhs_fwd
//   vertex/impl/helper/backward/BackpropagateAdjoint.java:63-143                   -> Engine::enqueue_rhs_aug
//   solve/impl/AdaptiveRungeKuttaSolver.java:107-181                               -> Engine::run_solve (+ kernels)
#include "engine.cuh"
#include "outer_net.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

namespace n4j {

static thread_local bool t_pdl = true;   // engine-wide switch (N4J_FLAG_NO_PDL), refreshed at the start of every API call

// Every kernel goes out through cudaLaunchKernelEx with programmatic stream serialization (see pdl_begin in common.cuh).
template <typename... KArgs, typename... Args>
static inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool pdl, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = (pdl && t_pdl) ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
    if (e != cudaSuccess) throw Error{N4J_ERR_CUDA, std::string("kernel launch failed: ") + cudaGetErrorString(e)};
}

#define LAUNCH(kernel, grid, block, stream, ...)                                              \
    do {                                                                                      \
        launch_k(kernel, dim3(grid), dim3(block), 0, (stream), pdl_, __VA_ARGS__);            \
        ++launches_;                                                                          \
    } while (0)

static inline long long round_up4(long long n) { return (n + 3) & ~3LL; }
static void set_conv_smem_all(size_t smem);

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda needed)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p)
            throw Error{N4J_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver"};
        fn = (EncodeTiledFn)p;
    }
    return fn;
}
// activation image [2*nch][rows][4 floats] with a {4, box_rows, box_chunks} box
static CUtensorMap make_image_map(float* img, long long rows, long long nch2, int box_rows, int box_chunks) {
    CUtensorMap tm;
    const cuuint64_t dims[3] = {4, (cuuint64_t)rows, (cuuint64_t)nch2};
    const cuuint64_t strides[2] = {16, (cuuint64_t)rows * 16};
    const cuuint32_t box[3] = {4, (cuuint32_t)box_rows, (cuuint32_t)box_chunks};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode_tiled_fn()(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, img, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error{N4J_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")"};
    return tm;
}
// stacked weight image [nblocks][nch][256][4 floats] with a {4, 256, 8, 1} box
static CUtensorMap make_weight_map(float* img, long long nch, long long nblocks) {
    CUtensorMap tm;
    const cuuint64_t dims[4] = {4, 256, (cuuint64_t)nch, (cuuint64_t)nblocks};
    const cuuint64_t strides[3] = {16, 256 * 16, (cuuint64_t)nch * 256 * 16};
    const cuuint32_t box[4] = {4, 256, 8, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode_tiled_fn()(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, img, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error{N4J_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")"};
    return tm;
}

// =====================================================================================================
// construction / planning
// =====================================================================================================
Engine::Engine(const n4j_graph_desc& g, const n4j_solver_cfg& cfg, int device) : device_(device), cfg_(cfg) {
    if (!(cfg.min_step < cfg.max_step))   // SolverConfig.java:22-24
        throw Error{N4J_ERR_INVALID_ARGUMENT, "Max step smaller than min step! Swapped arguments? max: " + std::to_string(cfg.max_step) +
                                                  " min " + std::to_string(cfg.min_step)};
    if (cfg_.order <= 0) cfg_.order = 5;
    if (cfg_.safety == 0) cfg_.safety = 0.9;
    if (cfg_.max_growth == 0) cfg_.max_growth = 10;
    if (cfg_.min_reduction == 0) cfg_.min_reduction = 0.2;
    if (cfg_.max_trials <= 0) cfg_.max_trials = 100000;
    exact_ = (cfg_.flags & N4J_FLAG_EXACT_CONTRACTIONS) != 0;
    pdl_ = (cfg_.flags & N4J_FLAG_NO_PDL) == 0;
    if (g.rank != 2 && g.rank != 4) throw Error{N4J_ERR_UNSUPPORTED, "Rank not supported: " + std::to_string(g.rank)};
    rank_ = g.rank;
    B_ = g.shape[0];
    if (rank_ == 2) { C_ = (int)g.shape[1]; H_ = 1; W_ = 1; }
    else { C_ = (int)g.shape[1]; H_ = (int)g.shape[2]; W_ = (int)g.shape[3]; }
    if (B_ <= 0 || C_ <= 0 || H_ <= 0 || W_ <= 0) throw Error{N4J_ERR_INVALID_ARGUMENT, "state shape must be positive"};
    nz_ = B_ * C_ * H_ * W_;
    nz_pad_ = round_up4(nz_);

    int c = C_;
    long long po = 0, go = 0;
    max_len_ = nz_;
    const long long rows = B_ * H_ * W_;
    for (int i = 0; i < g.n_layers; ++i) {
        LayerPlan lp{};
        lp.d = g.layers[i];
        lp.rows = rows; lp.H = H_; lp.W = W_;
        lp.c_in = c;
        switch (lp.d.kind) {
            case N4J_LAYER_DENSE:
                if (rank_ != 2) throw Error{N4J_ERR_UNSUPPORTED, "dense layer inside a convolutional ODE block is not supported"};
                if (lp.d.n_in == 0) lp.d.n_in = c;
                if (lp.d.n_in != c) throw Error{N4J_ERR_INVALID_ARGUMENT, "dense nIn does not match the previous layer"};
                lp.c_out = (int)lp.d.n_out;
                lp.p_len = lp.d.n_in * lp.d.n_out + (lp.d.has_bias ? lp.d.n_out : 0);
                lp.g_len = lp.p_len;
                lp.exact = lp.d.n_in <= 128 && lp.d.n_out <= 128;
                break;
            case N4J_LAYER_CONV3X3:
                if (rank_ != 4) throw Error{N4J_ERR_UNSUPPORTED, "conv3x3 needs a rank-4 state"};
                if (lp.d.n_in == 0) lp.d.n_in = c;
                if (lp.d.n_in != c) throw Error{N4J_ERR_INVALID_ARGUMENT, "conv nIn does not match the previous layer"};
                if (lp.d.has_bias) throw Error{N4J_ERR_UNSUPPORTED, "conv3x3 with bias is outside the native set"};
                lp.c_out = (int)lp.d.n_out;
                lp.p_len = lp.d.n_out * lp.d.n_in * 9;
                lp.g_len = lp.p_len;
                break;
            case N4J_LAYER_BATCHNORM:
                if (lp.d.n_out == 0) lp.d.n_out = c;
                if (lp.d.n_out != c) throw Error{N4J_ERR_INVALID_ARGUMENT, "batchnorm size does not match the previous layer"};
                lp.d.n_in = c;
                lp.c_out = c;
                lp.p_len = 4LL * c;    // gamma, beta, mean, var
                lp.g_len = 2LL * c;    // GradientViewSelectionFromBlacklisted.java:33-37
                if (lp.d.eps == 0.f) lp.d.eps = 1e-5f;
                break;
            case N4J_LAYER_ACTIVATION:
                lp.c_out = c; lp.p_len = 0; lp.g_len = 0;
                break;
            case N4J_LAYER_TIME_CONCAT:
                // OdeVertex.Builder(conf, name, vertex, addTimeAsInput = true, ...): only the first vertex is fed the time
                if (i != 0) throw Error{N4J_ERR_UNSUPPORTED, "time concat is only supported as the first vertex of the inner graph"};
                if (lp.d.n_out <= 0) lp.d.n_out = 1;       // ShapeMatchVertex(MergeVertex): overrideSizeDims {1: 1}
                lp.d.n_in = c; lp.d.activation = N4J_ACT_IDENTITY;
                lp.c_out = c + (int)lp.d.n_out; lp.p_len = 0; lp.g_len = 0;
                time_input_ = true;
                break;
            default:
                throw Error{N4J_ERR_UNSUPPORTED, "layer kind outside the native set {Dense, Convolution2D 3x3 same, BatchNormalization, Activation, time concat}"};
        }
        lp.in_len = rows * lp.c_in;
        lp.out_len = rows * lp.c_out;
        lp.p_off = po; lp.g_off = go;
        po += lp.p_len; go += lp.g_len;
        c = lp.c_out;
        max_len_ = std::max(max_len_, lp.out_len);
        L_.push_back(lp);
    }
    if (L_.empty()) throw Error{N4J_ERR_INVALID_ARGUMENT, "inner graph has no layers"};
    for (int i = (int)L_.size() - 1; i >= 0; --i) if (L_[i].d.kind == N4J_LAYER_CONV3X3) first_conv_ = i;
    if (c != C_) throw Error{N4J_ERR_INVALID_ARGUMENT, "inner graph must map the state shape onto itself"};
    n_params_ = po; n_real_ = go;
    real_pad_ = round_up4(n_real_);
    off_a_ = nz_pad_;
    off_theta_ = 2 * nz_pad_;
    off_t_ = off_theta_ + real_pad_;
    n_alloc_ = off_t_ + 4;

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        throw Error{N4J_ERR_CUDA, "no CUDA device available: the node4j engine has no CPU fallback"};
    if (device < 0 || device >= ndev) throw Error{N4J_ERR_INVALID_ARGUMENT, "bad device index"};
    N4J_CUDA(cudaSetDevice(device_));
    N4J_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    N4J_CUDA(cudaStreamCreateWithFlags(&stream2_, cudaStreamNonBlocking));
    {   // the side branch (weight gradients) yields to the main chain when both have blocks pending (N4J_SIDE_PRIO=0 switches it off, A/B)
        int lo = 0, hi = 0;
        N4J_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        const char* sp = getenv("N4J_SIDE_PRIO");
        if (sp && atoi(sp) == 0) N4J_CUDA(cudaStreamCreateWithFlags(&side_, cudaStreamNonBlocking));
        else N4J_CUDA(cudaStreamCreateWithPriority(&side_, cudaStreamNonBlocking, lo));
    }
    for (auto& e : ev_pool_) N4J_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : slot_event_) N4J_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    N4J_CUDA(cudaEventCreate(&ev0_));
    N4J_CUDA(cudaEventCreate(&ev1_));
    N4J_CUDA(cudaMalloc(&Y_, sizeof(float) * 2 * n_alloc_));
    N4J_CUDA(cudaMalloc(&K_, sizeof(float) * 7 * n_alloc_));
    N4J_CUDA(cudaMemsetAsync(Y_, 0, sizeof(float) * 2 * n_alloc_, stream_));
    N4J_CUDA(cudaMemsetAsync(K_, 0, sizeof(float) * 7 * n_alloc_, stream_));
    N4J_CUDA(cudaMalloc(&ctrl_, sizeof(Ctrl)));
    N4J_CUDA(cudaMemsetAsync(ctrl_, 0, sizeof(Ctrl), stream_));
    N4J_CUDA(cudaMalloc(&red_acc_, sizeof(double) * 4 * ACC_SLOTS * ACC_STRIDE));
    N4J_CUDA(cudaMemsetAsync(red_acc_, 0, sizeof(double) * 4 * ACC_SLOTS * ACC_STRIDE, stream_));
    N4J_CUDA(cudaMemcpyAsync(&ctrl_->red_acc, &red_acc_, sizeof(double*), cudaMemcpyHostToDevice, stream_));
    N4J_CUDA(cudaHostAlloc(&ctrl_host_, sizeof(Ctrl), cudaHostAllocDefault));
    N4J_CUDA(cudaMalloc(&log_, sizeof(n4j_step_rec) * log_cap_));
    max_pairs_ = 4096;
    N4J_CUDA(cudaMalloc(&tpair_, sizeof(float) * 2 * max_pairs_));
    N4J_CUDA(cudaHostAlloc(&tpair_host_, sizeof(float) * 2 * max_pairs_, cudaHostAllocDefault));
    wanted_cap_ = max_pairs_;
    N4J_CUDA(cudaMalloc(&wanted_, sizeof(float) * wanted_cap_));
    N4J_CUDA(cudaMalloc(&dot_acc_, sizeof(double)));
    N4J_CUDA(cudaMemsetAsync(dot_acc_, 0, sizeof(double), stream_));
    N4J_CUDA(cudaMalloc(&tg_, sizeof(float) * 4));
    N4J_CUDA(cudaMalloc(&tcat_acc_, sizeof(double) * ACC_SLOTS * ACC_STRIDE));
    N4J_CUDA(cudaMemsetAsync(tcat_acc_, 0, sizeof(double) * ACC_SLOTS * ACC_STRIDE, stream_));
    N4J_CUDA(cudaMalloc(&tcat_misc_, 16));      // [0]: ticket, [1]: last g^T df/dt
    N4J_CUDA(cudaMemsetAsync(tcat_misc_, 0, 16, stream_));
    ensure(params_, std::max<long long>(n_params_, 1));
    ensure(gradflat_, std::max<long long>(n_params_, 1));
    ensure(theta_tmp_, std::max<long long>(real_pad_, 4));
    for (auto& b : gb_) ensure(b, max_len_);
    int n_def = 0;
    for (size_t i = 0; i < L_.size(); ++i) {
        LayerPlan& l = L_[i];
        for (int sl = 0; sl < 2; ++sl) {   // two sets: consecutive stage evaluations alternate, so a weight-gradient kernel of
                                           // stage s may still read its inputs while stage s+1 runs
            if (i + 1 < L_.size()) N4J_CUDA(cudaMalloc(&l.act2[sl], sizeof(float) * l.out_len));
            if (i > 0) N4J_CUDA(cudaMalloc(&l.gin2[sl], sizeof(float) * l.in_len));
        }
        l.act = l.act2[0]; l.gin = l.gin2[0];
        if (l.d.kind == N4J_LAYER_CONV3X3) {
            N4J_CUDA(cudaMalloc(&l.wf, sizeof(float) * l.p_len));
            N4J_CUDA(cudaMalloc(&l.wb, sizeof(float) * l.p_len));
            ensure(wpart_, 16LL * l.p_len);
            // tensor-core path: 64->64 channels, identity activation, default (non-parity) mode
            l.tc = !exact_ && !(cfg_.flags & N4J_FLAG_NO_TENSOR) && l.c_in == 64 && l.c_out == 64 && l.d.activation == N4J_ACT_IDENTITY;
            if (l.tc && (long long)B_ * (H_ + 2) * (W_ + 2) * 64 >= (1LL << 31)) l.tc = false;   // the tensor kernels index padded pixels x channels in 32 bits
            if (l.tc) {
                // large feature maps: the staged tile (+halo) or the per-image weight-gradient staging would not fit one SM's shared
                // memory -> the fp32 SIMT kernels take the layer
                const ConvGeom gg = make_conv_geom((int)B_, H_, W_);
                if (conv_tc_smem_bytes(gg) + 4096 > 227 * 1024 || conv_wgrad_smem_bytes(gg) > 200 * 1024) l.tc = false;
            }
            if (l.tc) {
                geom_ = make_conv_geom((int)B_, H_, W_);
                conv_smem_ = conv_tc_smem_bytes(geom_);
                const size_t img_bytes = sizeof(float) * conv_image_floats(geom_);
                N4J_CUDA(cudaMalloc(&l.img, img_bytes));
                N4J_CUDA(cudaMalloc(&l.gimg, img_bytes));
                N4J_CUDA(cudaMemsetAsync(l.img, 0, img_bytes, stream_));    // zero padding is written once and never touched
                N4J_CUDA(cudaMemsetAsync(l.gimg, 0, img_bytes, stream_));
                N4J_CUDA(cudaMalloc(&l.wimg_f, sizeof(float) * 9 * 2 * 4096));
                N4J_CUDA(cudaMalloc(&l.wimg_b, sizeof(float) * 9 * 2 * 4096));
                set_conv_smem_all(conv_smem_);
                side_capable_ = true;
                wgrad_smem_ = conv_wgrad_smem_bytes(geom_);
                wgrad_nbuf_ = 2 * wgrad_smem_ <= 200 * 1024 ? 2 : 1;      // second staging buffer when it fits
                wgrad_smem_ *= wgrad_nbuf_;
                N4J_CUDA(cudaFuncSetAttribute(k_conv_wgrad64, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wgrad_smem_));
                ensure(wpart_, (long long)B_ * 9 * 4096);
                // tensor-core weight gradient (conv_wgrad_tc.cuh): per-tile partials, tiles of 128 shared-border padded pixels
                wgeom_ = make_wgrad_tc_geom((int)B_, H_, W_);
                wgrad_tc_smem_ = conv_wgrad_tc_smem_bytes(wgeom_);
                wgrad_tc_ = !(cfg_.flags & N4J_FLAG_SIMT_WGRAD) && wgrad_tc_smem_ <= 227 * 1024;
                if (wgrad_tc_) {
                    N4J_CUDA(cudaFuncSetAttribute(k_conv_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wgrad_tc_smem_));
                    ensure(wpart_, (long long)wgeom_.tiles * 9 * 4096);
                }
            }
        }
        if (l.d.kind == N4J_LAYER_DENSE && !l.exact && !exact_ && !(cfg_.flags & N4J_FLAG_NO_TENSOR) && l.c_in % 32 == 0 && l.c_out % 32 == 0 &&
            (l.rows * l.c_in * l.c_out >= (1LL << 27) || (cfg_.flags & N4J_FLAG_FORCE_TENSOR))) {   // small layers: launch latency beats the tensor pipe
            // tensor-core dense path: split images of the layer input / output delta and stacked weight images, moved by TMA
            l.tcd = true;
            l.img_rows = (l.rows + 127) / 128 * 128;
            const size_t xb = sizeof(float) * 2 * l.c_in * l.img_rows, db = sizeof(float) * 2 * l.c_out * l.img_rows;
            const long long nbf = (l.c_out + 127) / 128, nbb = (l.c_in + 127) / 128;
            const size_t wfb = sizeof(float) * nbf * (l.c_in / 4) * 256 * 4, wbb = sizeof(float) * nbb * (l.c_out / 4) * 256 * 4;
            N4J_CUDA(cudaMalloc(&l.ximg, xb)); N4J_CUDA(cudaMemsetAsync(l.ximg, 0, xb, stream_));
            N4J_CUDA(cudaMalloc(&l.dimg, db)); N4J_CUDA(cudaMemsetAsync(l.dimg, 0, db, stream_));
            N4J_CUDA(cudaMalloc(&l.wsf, wfb)); N4J_CUDA(cudaMemsetAsync(l.wsf, 0, wfb, stream_));
            N4J_CUDA(cudaMalloc(&l.wsb, wbb)); N4J_CUDA(cudaMemsetAsync(l.wsb, 0, wbb, stream_));
            const long long oc_pad = (l.c_out + 127) / 128 * 128;
            const size_t dtb = sizeof(float) * 2 * l.img_rows * oc_pad, xtb = sizeof(float) * nbb * (l.img_rows / 4) * 256 * 4;
            N4J_CUDA(cudaMalloc(&l.dTimg, dtb)); N4J_CUDA(cudaMemsetAsync(l.dTimg, 0, dtb, stream_));
            N4J_CUDA(cudaMalloc(&l.xTimg, xtb)); N4J_CUDA(cudaMemsetAsync(l.xTimg, 0, xtb, stream_));
            l.tm_x_k = make_image_map(l.ximg, l.img_rows, 2 * (l.c_in / 4), 128, 8);
            l.tm_d_k = make_image_map(l.dimg, l.img_rows, 2 * (l.c_out / 4), 128, 8);
            l.tm_dT = make_image_map(l.dTimg, oc_pad, 2 * (l.img_rows / 4), 128, 8);
            l.tm_xT = make_weight_map(l.xTimg, l.img_rows / 4, nbb);
            l.tm_wf = make_weight_map(l.wsf, l.c_in / 4, nbf);
            l.tm_wb = make_weight_map(l.wsb, l.c_out / 4, nbb);
            N4J_CUDA(cudaFuncSetAttribute(k_gemm_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_TC_SMEM));
            N4J_CUDA(cudaFuncSetAttribute(k_gemm_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_TC2_SMEM));
        }
        if (l.d.kind == N4J_LAYER_DENSE && !l.exact && !l.tcd && (long long)l.c_in * l.c_out <= (1LL << 22)) ensure(wpart_, 8LL * l.c_in * l.c_out);
        if (l.d.kind == N4J_LAYER_BATCHNORM) {
            const int cc = l.c_out;
            N4J_CUDA(cudaMalloc(&l.bn.acc, sizeof(double) * 2 * cc * ACC_STRIDE));
            N4J_CUDA(cudaMemsetAsync(l.bn.acc, 0, sizeof(double) * 2 * cc * ACC_STRIDE, stream_));
            N4J_CUDA(cudaMalloc(&l.bn.ticket, sizeof(unsigned int)));
            N4J_CUDA(cudaMemsetAsync(l.bn.ticket, 0, sizeof(unsigned int), stream_));
            float* f = nullptr;
            N4J_CUDA(cudaMalloc(&f, sizeof(float) * 4 * cc));
            l.bn.mean = f; l.bn.invstd = f + cc; l.bn.m_dy = f + 2 * cc; l.bn.m_dyxh = f + 3 * cc;
            l.bn.eps = l.d.eps; l.bn.rows = l.rows;
            l.bn.fast_final = (exact_ || getenv("N4J_EXACT_BN_FINAL")) ? 0 : 1;     // parity mode keeps the oracle's divisions (env: A/B switch)
            l.bn.inv_rows = 1.0 / (double)l.rows;
            l.bn_deferred_ok = (cc == 64) && !(cfg_.flags & N4J_FLAG_NO_DEFERRED_STATS);
            if (l.bn_deferred_ok) ++n_def;
        }
    }
    // deferred statistics: [layer][forward|backward][set 0|1][2*64 strided doubles], one allocation so that one memset re-arms it
    if (n_def > 0) {
        bn_pool_len_ = (size_t)n_def * 4 * 128 * ACC_STRIDE;
        N4J_CUDA(cudaMalloc(&bn_pool_, sizeof(double) * bn_pool_len_));
        N4J_CUDA(cudaMemsetAsync(bn_pool_, 0, sizeof(double) * bn_pool_len_, stream_));
        int k = 0;
        for (auto& l : L_) if (l.d.kind == N4J_LAYER_BATCHNORM && l.bn_deferred_ok) { l.bn_pool = bn_pool_ + (size_t)(k++) * 4 * 128 * ACC_STRIDE; l.bn.deferred = 1; }
        set_bn_slot(0);
    }
    // small dense inner graph (optionally behind the time merge): one fused launch per evaluation (mlp_small.cuh)
    {
        bool ok = !(cfg_.flags & N4J_FLAG_NO_MLP_FUSION) && rank_ == 2 && n_params_ == n_real_ && n_params_ <= 12288;
        MlpSmallArgs m{};
        int feat = C_;
        for (size_t i = 0; ok && i < L_.size(); ++i) {
            const LayerPlan& l = L_[i];
            if (l.d.kind == N4J_LAYER_TIME_CONCAT) { m.tcat_k = l.c_out - l.c_in; feat = l.c_out; continue; }
            if (l.d.kind != N4J_LAYER_DENSE || !l.exact || m.nl >= MLP_MAX_LAYERS) { ok = false; break; }
            const int k = m.nl++;
            m.nin[k] = l.c_in; m.nout[k] = l.c_out; m.act[k] = l.d.activation; m.has_bias[k] = l.d.has_bias; m.p_off[k] = (int)l.p_off;
            m.a_off[k] = k == 0 ? 0 : m.a_off[k - 1] + m.nin[k - 1];
            feat = l.c_out;
        }
        if (ok && m.nl > 0) {
            m.a_off[m.nl] = m.a_off[m.nl - 1] + m.nin[m.nl - 1];
            m.a_tot = m.a_off[m.nl] + C_;
            m.D = C_; m.n_params = (int)n_params_; m.B = B_;
            m.off_a = off_a_; m.off_theta = off_theta_; m.off_t = off_t_;
            // rows per block: as few blocks as shared memory allows (their partial parameter gradients are summed by ONE block)
            mlp_smem_ = mlp_small_smem_bytes(m);
            if (mlp_smem_ <= 200 * 1024 && feat == C_) {
                const int blocks = (int)((B_ + MLP_ROWS - 1) / MLP_ROWS);
                int group = 1;
                while (group * group < blocks) ++group;                      // ~sqrt(blocks) blocks per group, as many groups
                const int groups = (blocks + group - 1) / group;
                m.group = group;
                N4J_CUDA(cudaMalloc(&mlp_part_, sizeof(double) * (size_t)(blocks + groups) * (n_params_ + 1)));
                N4J_CUDA(cudaMalloc(&mlp_ticket_, sizeof(unsigned int) * (1 + groups)));
                N4J_CUDA(cudaMemsetAsync(mlp_ticket_, 0, sizeof(unsigned int) * (1 + groups), stream_));
                N4J_CUDA(cudaFuncSetAttribute(k_mlp_small<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mlp_smem_));
                N4J_CUDA(cudaFuncSetAttribute(k_mlp_small<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mlp_smem_));
                m.part = mlp_part_; m.ticket = mlp_ticket_; m.dt_out = tcat_misc_ + 1; m.c = ctrl_;
                mlp_args_ = m;
                mlp_fused_ = true;
            }
        }
    }
    N4J_CUDA(cudaStreamSynchronize(stream_));
}

Engine::~Engine() {
    cudaSetDevice(device_);
    if (stream_) cudaStreamSynchronize(stream_);
    for (auto& kv : graphs_) {
        if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
        if (kv.second.graph) cudaGraphDestroy(kv.second.graph);
    }
    for (auto& l : L_) {
        for (int sl = 0; sl < 2; ++sl) { cudaFree(l.act2[sl]); cudaFree(l.gin2[sl]); }
        cudaFree(l.ximg); cudaFree(l.dimg); cudaFree(l.dTimg); cudaFree(l.xTimg); cudaFree(l.wsf); cudaFree(l.wsb);
        cudaFree(l.wf); cudaFree(l.wb); cudaFree(l.img); cudaFree(l.gimg); cudaFree(l.wimg_f); cudaFree(l.wimg_b);
        if (l.d.kind == N4J_LAYER_BATCHNORM) { cudaFree(l.bn.acc); cudaFree(l.bn.ticket); cudaFree(l.bn.mean); cudaFree(l.bn_seq); }
    }
    for (DevBuf* b : {&params_, &gradflat_, &zt_, &gt_, &io_, &io2_, &gb_[0], &gb_[1], &gb_[2], &theta_tmp_, &wpart_, &dldt_}) cudaFree(b->p);
    for (void* pp : peer_ptrs_) cudaIpcCloseMemHandle(pp);
    cudaFree(comm_region_); cudaFree(comm_dev_);
    cudaFree(Y_); cudaFree(K_); cudaFree(ctrl_); cudaFree(red_acc_); cudaFree(log_); cudaFree(tpair_); cudaFree(wanted_); cudaFree(dot_acc_); cudaFree(tg_); cudaFree(tcat_acc_); cudaFree(tcat_misc_); cudaFree(bn_pool_); cudaFree(mlp_part_); cudaFree(mlp_ticket_);
    cudaFreeHost(ctrl_host_); cudaFreeHost(tpair_host_);
    if (ev0_) cudaEventDestroy(ev0_);
    if (ev1_) cudaEventDestroy(ev1_);
    if (stream_) cudaStreamDestroy(stream_);
    if (stream2_) cudaStreamDestroy(stream2_);
    if (side_) cudaStreamDestroy(side_);
    for (auto& e : ev_pool_) if (e) cudaEventDestroy(e);
    for (auto& e : slot_event_) if (e) cudaEventDestroy(e);
}

// =====================================================================================================
// small helpers
// =====================================================================================================
void Engine::ensure(DevBuf& b, long long n) {
    if (b.n >= n) return;
    if (b.p) { N4J_CUDA(cudaStreamSynchronize(stream_)); cudaFree(b.p); }
    N4J_CUDA(cudaMalloc(&b.p, sizeof(float) * n));
    b.n = n;
}
// zt_ is referenced by captured dense-output kernels: drop the cached graphs when it moves
void Engine::ensure_zt(long long n) {
    if (zt_.n >= n) return;
    ensure(zt_, n);
    have_last_output_ = false;
    for (auto& kv : graphs_) {
        if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
        if (kv.second.graph) cudaGraphDestroy(kv.second.graph);
    }
    graphs_.clear();
}
bool Engine::is_device_ptr(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}
void Engine::to_device(float* dst, const float* src, long long n) {
    N4J_CUDA(cudaMemcpyAsync(dst, src, sizeof(float) * n, cudaMemcpyDefault, stream_));
}
void Engine::from_device(float* dst, const float* src, long long n) {
    N4J_CUDA(cudaMemcpyAsync(dst, src, sizeof(float) * n, cudaMemcpyDefault, stream_));
}
int Engine::grid_for(long long n, int per_thread) const {
    long long blocks = (n + 256LL * per_thread - 1) / (256LL * per_thread);
    blocks = std::max<long long>(1, std::min<long long>(blocks, 148LL * 8));
    return (int)blocks;
}
// Grid for kernels that end in a cross-block reduction (double atomics + ticket on a handful of addresses): same-address
// atomics serialise in L2, so the block count is capped at two per SM (measured best on B200, r01) and shrinks with the problem.
int Engine::grid_red(long long n4) const {
    static const int per = getenv("N4J_GRID_RED_PER") ? atoi(getenv("N4J_GRID_RED_PER")) : 2;
    static const int cap = getenv("N4J_GRID_RED_CAP") ? atoi(getenv("N4J_GRID_RED_CAP")) : 296;
    long long blocks = (n4 + 256LL * per - 1) / (256LL * per);
    blocks = std::max<long long>(1, std::min<long long>(blocks, n4 > (4LL << 20) ? 148 * 8 : cap));   // HBM-sized states need the extra loads in flight
    return (int)blocks;
}
// Deferred BatchNorm statistics: which accumulator set the evaluation being enqueued uses.  Consecutive evaluations alternate;
// every solve (prologue: 2 evaluations, trial body: 6) starts on set 1 and ends on set 0, so the captured graphs bake a
// position-independent pattern.  Eager one-off evaluations re-arm everything afterwards (bn_rearm_all).
void Engine::set_bn_slot(int slot) {
    bn_slot_ = slot;
    for (auto& l : L_) {
        if (l.d.kind != N4J_LAYER_BATCHNORM || !l.bn_pool) continue;
        double* base = l.bn_pool;
        const size_t set = 128 * ACC_STRIDE;
        l.bn.accf = base + (size_t)(0 * 2 + slot) * set;
        l.bn.accf_other = base + (size_t)(0 * 2 + (slot ^ 1)) * set;
        l.bn.accb = base + (size_t)(1 * 2 + slot) * set;
        l.bn.accb_other = base + (size_t)(1 * 2 + (slot ^ 1)) * set;
        if (l.bn_seq) {     // sharded: this set's sequence counters and mailboxes
            const long long mset = (long long)COMM_MAX_WORLD * 128 * 16;
            l.bn.seqf = l.bn_seq + (0 * 2 + slot); l.bn.seqb = l.bn_seq + (1 * 2 + slot);
            l.bn.mbf = l.bn_mb + (0 * 2 + slot) * mset; l.bn.mbb = l.bn_mb + (1 * 2 + slot) * mset;
        }
    }
}
void Engine::begin_eval() {
    if (eval_begun_) { eval_begun_ = false; return; }   // the caller already switched sets (stage combination in front of the evaluation)
    set_bn_slot(bn_slot_ ^ 1);
}
void Engine::bn_rearm_all(cudaStream_t s) {
    if (bn_pool_) N4J_CUDA(cudaMemsetAsync(bn_pool_, 0, sizeof(double) * bn_pool_len_, s));
    set_bn_slot(0);
    eval_begun_ = false;
}
// stage combination + statistics of a leading 64-channel BatchNorm in one pass
BnFuse Engine::bn_fuse() const {
    BnFuse bf{};
    const LayerPlan& l = L_[0];
    if (rank_ == 4 && l.d.kind == N4J_LAYER_BATCHNORM && l.c_out == 64 && nz_ % 4 == 0) {
        bf.enabled = 1; bf.sc = l.bn; bf.eps = l.d.eps; bf.rows = l.rows; bf.nz4 = nz_ / 4;
    }
    return bf;
}
SolverParams Engine::solver_params(long long n_norm) const {
    SolverParams p;
    p.atol = (float)cfg_.abs_tol; p.rtol = (float)cfg_.rel_tol;
    p.hmin = (float)cfg_.min_step; p.hmax = (float)cfg_.max_step;
    p.safety = (float)cfg_.safety; p.max_growth = (float)cfg_.max_growth; p.min_reduction = (float)cfg_.min_reduction;
    p.order = cfg_.order; p.flags = cfg_.flags;
    p.n_norm = n_norm; p.max_trials = cfg_.max_trials; p.log_cap = log_cap_;
    p.rep_begin4 = 1LL << 60; p.count_rep = 1;
    return p;
}

void Engine::bind_params(const float* p, long long n) {
    if (n != n_params_) throw Error{N4J_ERR_INVALID_ARGUMENT, "parameter vector has " + std::to_string(n) + " entries, inner graph needs " + std::to_string(n_params_)};
    user_params_ = p;
    params_fresh_ = false;
}

// re-read the caller's flat parameter buffer (the updater changes it between iterations) and derive the
// internal conv layouts
void Engine::upload_params() {
    if (n_params_ == 0) return;
    if (!user_params_) throw Error{N4J_ERR_ILLEGAL_STATE, "parameters not bound (node4j_bind_params)"};
    to_device(params_.p, user_params_, n_params_);
    for (auto& l : L_)
        if (l.tcd) {
            const float* Wt = params_.p + l.p_off;     // [nOut][nIn] (DL4J's f-order [nIn,nOut])
            LAUNCH(k_weights_to_stacked_image, grid_for((long long)l.c_in * l.c_out), 256, stream_, Wt, l.wsf, l.c_in, l.c_out, 0);
            LAUNCH(k_weights_to_stacked_image, grid_for((long long)l.c_in * l.c_out), 256, stream_, Wt, l.wsb, l.c_out, l.c_in, 1);
        }
    for (auto& l : L_)
        if (l.d.kind == N4J_LAYER_CONV3X3) {
            LAUNCH(k_conv_w_to_internal, grid_for(l.p_len), 256, stream_, params_.p + l.p_off, l.wf, l.wb, l.c_out, l.c_in);
            if (l.tc) {
                LAUNCH(k_conv_w_to_tc, 64, 256, stream_, params_.p + l.p_off, l.wimg_f, 0);
                LAUNCH(k_conv_w_to_tc, 64, 256, stream_, params_.p + l.p_off, l.wimg_b, 1);
            }
        }
}

void Engine::state_in(const float* user, float* internal) {
    if (rank_ == 2 || nhwc_io_) { to_device(internal, user, nz_); return; }
    ensure(io2_, nz_);
    to_device(io2_.p, user, nz_);
    LAUNCH(k_nchw_to_nhwc, grid_for(nz_), 256, stream_, io2_.p, internal, (int)B_, C_, H_ * W_);
}
void Engine::state_out(const float* internal, float* user_dev) {
    if (rank_ == 2 || nhwc_io_) { N4J_CUDA(cudaMemcpyAsync(user_dev, internal, sizeof(float) * nz_, cudaMemcpyDeviceToDevice, stream_)); return; }
    LAUNCH(k_nhwc_to_nchw, grid_for(nz_), 256, stream_, internal, user_dev, (int)B_, C_, H_ * W_);
}
void Engine::theta_flat_to_internal(const float* flat, float* theta) {
    for (auto& l : L_) {
        if (l.g_len == 0) continue;
        if (l.d.kind == N4J_LAYER_CONV3X3) LAUNCH(k_conv_g_flat_to_internal, grid_for(l.g_len), 256, stream_, flat + l.p_off, theta + l.g_off, l.c_out, l.c_in);
        else N4J_CUDA(cudaMemcpyAsync(theta + l.g_off, flat + l.p_off, sizeof(float) * l.g_len, cudaMemcpyDeviceToDevice, stream_));
    }
}
void Engine::theta_internal_to_flat(const float* theta, float* flat) {
    for (auto& l : L_) {
        if (l.g_len == 0) continue;
        if (l.d.kind == N4J_LAYER_CONV3X3) LAUNCH(k_conv_g_internal_to_flat, grid_for(l.g_len), 256, stream_, theta + l.g_off, flat + l.p_off, l.c_out, l.c_in);
        else N4J_CUDA(cudaMemcpyAsync(flat + l.p_off, theta + l.g_off, sizeof(float) * l.g_len, cudaMemcpyDeviceToDevice, stream_));
    }
}
void Engine::theta_internal_to_real(const float* theta, float* real) {
    for (auto& l : L_) {
        if (l.g_len == 0) continue;
        if (l.d.kind == N4J_LAYER_CONV3X3) LAUNCH(k_conv_g_internal_to_flat, grid_for(l.g_len), 256, stream_, theta + l.g_off, real + l.g_off, l.c_out, l.c_in);
        else N4J_CUDA(cudaMemcpyAsync(real + l.g_off, theta + l.g_off, sizeof(float) * l.g_len, cudaMemcpyDeviceToDevice, stream_));
    }
}

// =====================================================================================================
// RHS: f(z,t) and the augmented dynamics
// =====================================================================================================
static inline Ref with_off(Ref r, long long off) { r.off += off; return r; }

template <int MODE>
static void launch_gemm(cudaStream_t s, GemmArgs g, int& launches, bool exact) {
    if (exact) g.splits = 1;   // double accumulation is only order-independent inside one accumulator chain
    dim3 grid((g.N + 63) / 64, (g.M + 63) / 64, g.splits);
    // (16-row tiles for parity mode — k_gemm_tiled<MODE, double, 1>, four times the blocks — were measured and are slower: 8.2 vs 5.4 ms
    // for the forward solve; with 1 x 4 outputs per thread the shared-memory loads, not the FP64 pipe, become the bound)
    constexpr bool dense = MODE == G_DENSE_FWD || MODE == G_DENSE_DGRAD || MODE == G_DENSE_WGRAD;
    constexpr bool a_kfast = MODE == G_DENSE_FWD || MODE == G_DENSE_DGRAD || MODE == G_CONV_FWD || MODE == G_CONV_DGRAD;
    const long long b64 = (long long)grid.x * grid.y, b48 = (long long)grid.x * ((g.M + 47) / 48);
    if constexpr (a_kfast) {
        // parity mode, FP64-pipe-bound: 48-row tiles when they still fit one wave and the 64-row tiles leave SMs idle (same sums, same order)
        if (exact && b64 < 148 && b48 <= 148 && b48 > b64) {
            grid.y = (g.M + 47) / 48;
            launch_k(k_gemm_tiled<MODE, double, 3>, grid, dim3(256), 0, s, true, g);
            ++launches;
            return;
        }
    }
    if (exact) launch_k(k_gemm_tiled<MODE, double>, grid, dim3(256), 0, s, true, g);
    else if (dense && g.K / g.splits >= 64) launch_k(k_gemm_tiled<MODE, float, 4, 32>, grid, dim3(256), 0, s, true, g);   // half as many dependent trips to L2
    else launch_k(k_gemm_tiled<MODE, float>, grid, dim3(256), 0, s, true, g);
    ++launches;
}

void Engine::launch_gemm_tc(cudaStream_t s, const CUtensorMap&, const CUtensorMap&, const GemmTcArgs& a) {
    const int tiles_n = (a.N + 127) / 128, tiles_m = (a.M + 127) / 128;
    // enough 128 x 128 tiles to fill the chip: pair the M tiles (256 x 128 per CTA, every B tile feeds two accumulators)
    if ((long long)tiles_m * tiles_n >= 148 && tiles_m >= 2) {
        const dim3 grid(((tiles_m + 1) / 2) * tiles_n), block(GEMM_TC_THREADS);
        launch_k(k_gemm_tc<2>, grid, block, GEMM_TC2_SMEM, s, pdl_, a);
    } else {
        const dim3 grid(tiles_m * tiles_n), block(GEMM_TC_THREADS);
        launch_k(k_gemm_tc<1>, grid, block, GEMM_TC_SMEM, s, pdl_, a);
    }
    ++launches_;
}

template <int PRO, int EPI>
static void launch_conv_variant(const ConvTcArgs& a, dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool pdl, bool fast) {
    const bool sh = a.pbn.cm != nullptr || a.bn.cm != nullptr;      // sharded instantiation: carries the cross-rank statistics exchange
    if (sh) {
        if (fast) launch_k(k_conv3x3_tc<1, EPI, PRO, 1>, grid, block, smem, s, pdl, a);
        else launch_k(k_conv3x3_tc<3, EPI, PRO, 1>, grid, block, smem, s, pdl, a);
    } else {
        if (fast) launch_k(k_conv3x3_tc<1, EPI, PRO, 0>, grid, block, smem, s, pdl, a);
        else launch_k(k_conv3x3_tc<3, EPI, PRO, 0>, grid, block, smem, s, pdl, a);
    }
}
template <int PRO, int EPI>
static void set_conv_smem(size_t smem) {
    N4J_CUDA(cudaFuncSetAttribute(k_conv3x3_tc<3, EPI, PRO, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    N4J_CUDA(cudaFuncSetAttribute(k_conv3x3_tc<1, EPI, PRO, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    N4J_CUDA(cudaFuncSetAttribute(k_conv3x3_tc<3, EPI, PRO, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    N4J_CUDA(cudaFuncSetAttribute(k_conv3x3_tc<1, EPI, PRO, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
}
// the (prologue, epilogue) combinations the engine launches: forward convs take PRO 0/1 with EPI 0/1, dgrad convs PRO 0/2 with EPI 0/2
static void set_conv_smem_all(size_t smem) {
    set_conv_smem<0, 0>(smem); set_conv_smem<0, 1>(smem); set_conv_smem<1, 0>(smem); set_conv_smem<1, 1>(smem);
    set_conv_smem<0, 2>(smem); set_conv_smem<2, 0>(smem); set_conv_smem<2, 2>(smem);
}

// pro_kind: fused operand prologue (1 = BatchNorm apply in front of a forward conv, 2 = BatchNorm backward in front of a dgrad)
// epi_kind: fused epilogue statistics (1 = forward statistics for next_bn, 2 = backward statistics of the BatchNorm in front, fields in *extra)
void Engine::launch_conv_tc(cudaStream_t s, const float* img, const float* wimg, Ref out, int act, const LayerPlan* next_bn,
                            const ConvTcArgs* extra, int pro_kind, int epi_kind) {
    ConvTcArgs a{};
    if (extra) a = *extra;
    a.img = img; a.wimg = wimg; a.out = out; a.g = geom_; a.act = act;
    a.dbg_nowstream = getenv("N4J_DBG_CONV_NOWSTREAM") ? 1 : 0;
    if (next_bn) { epi_kind = 1; a.bn = next_bn->bn; a.bn_eps = next_bn->d.eps; a.bn_rows = next_bn->rows; }
    a.fuse_stats = epi_kind;
    if (conv_rec_) conv_rec_->push_back(ConvRec{a, pro_kind, epi_kind});
    launch_conv_args(s, a, pro_kind, epi_kind);
}
void Engine::launch_conv_args(cudaStream_t s, const ConvTcArgs& a, int pro_kind, int epi_kind) {
    const bool fast = (cfg_.flags & N4J_FLAG_FAST_TF32) != 0;
    const dim3 grid(geom_.tiles), block(conv_tc_threads(pro_kind));
    const int combo = pro_kind * 3 + epi_kind;
    switch (combo) {
        case 0: launch_conv_variant<0, 0>(a, grid, block, conv_smem_, s, pdl_, fast); break;
        case 1: launch_conv_variant<0, 1>(a, grid, block, conv_smem_, s, pdl_, fast); break;
        case 2: launch_conv_variant<0, 2>(a, grid, block, conv_smem_, s, pdl_, fast); break;
        case 3: launch_conv_variant<1, 0>(a, grid, block, conv_smem_, s, pdl_, fast); break;
        case 4: launch_conv_variant<1, 1>(a, grid, block, conv_smem_, s, pdl_, fast); break;
        case 6: launch_conv_variant<2, 0>(a, grid, block, conv_smem_, s, pdl_, fast); break;
        case 8: launch_conv_variant<2, 2>(a, grid, block, conv_smem_, s, pdl_, fast); break;
        default: throw Error{N4J_ERR_ILLEGAL_STATE, "conv kernel variant not instantiated"};
    }
    ++launches_;
}

void Engine::enqueue_rhs_fwd(cudaStream_t s, Ref in, Ref out, bool need_compact, bool bn1_ready) {
    Ref cur = in;
    bool image_ready = false;   // the previous layer already wrote the operand image of this (tensor-core) conv
    bool pro_pending = false;   // ... or left its BatchNorm apply to the conv kernel's operand prologue
    if (mlp_fused_ && !comm_dev_ && !in_aug_) { begin_eval(); enqueue_mlp_small(s, in, out, false); return; }
    if (!in_aug_) begin_eval();
    ConvTcArgs pro_args{};
    bool stats_ready = bn1_ready;   // the previous kernel (conv epilogue / stage combination) already produced this BatchNorm's mean / invstd
    for (size_t i = 0; i < L_.size(); ++i) {
        LayerPlan& l = L_[i];
        Ref dst = (i + 1 == L_.size()) ? out : fixed_ref(l.act);
        switch (l.d.kind) {
            case N4J_LAYER_DENSE: {
                const float* Wt = params_.p + l.p_off;
                const float* bias = l.d.has_bias ? Wt + l.d.n_in * l.d.n_out : nullptr;
                if (l.exact) {
                    LAUNCH(k_dense_fwd_exact, grid_for(l.out_len), 256, s, cur, dst, Wt, bias, l.d.activation, l.rows, l.c_in, l.c_out);
                } else if (l.tcd) {
                    LAUNCH(k_rowmajor_to_image, grid_for(l.in_len / 4), 256, s, cur, 1.f, l.ximg, l.rows, l.c_in, l.img_rows);
                    GemmTcArgs ga{};
                    ga.M = (int)l.rows; ga.N = l.c_out; ga.kblocks = l.c_in / 32; ga.a_nch = l.c_in / 4; ga.out = dst; ga.bias = bias; ga.act = l.d.activation;
                    ga.a_img = l.ximg; ga.a_rows = l.img_rows; ga.b_img = l.wsf; ga.b_nch = l.c_in / 4;
                    launch_gemm_tc(s, l.tm_x_k, l.tm_wf, ga);
                } else {
                    GemmArgs g{}; g.a = cur; g.c = dst; g.w = Wt; g.bias = bias; g.M = (int)l.rows; g.N = l.c_out; g.K = l.c_in;
                    g.act = l.d.activation; g.splits = 1;
                    launch_gemm<G_DENSE_FWD>(s, g, launches_, exact_);
                }
                break;
            }
            case N4J_LAYER_TIME_CONCAT:
                LAUNCH(k_time_concat_fwd, grid_for(l.out_len), 256, s, ctrl_, tspec_, cur, dst, l.rows, l.c_in, l.c_out - l.c_in);
                break;
            case N4J_LAYER_CONV3X3: {
                if (l.tc) {
                    if (!image_ready) LAUNCH(k_compact_to_image, grid_for(l.in_len / 4), 256, s, cur, geom_, l.img);
                    // a BatchNorm right after the conv gets its batch statistics from the conv epilogue
                    const bool bn_next = i + 1 < L_.size() && L_[i + 1].d.kind == N4J_LAYER_BATCHNORM && L_[i + 1].c_out == 64;
                    launch_conv_tc(s, l.img, l.wimg_f, dst, l.d.activation, bn_next ? &L_[i + 1] : nullptr, pro_pending ? &pro_args : nullptr,
                                   pro_pending ? 1 : 0);
                    pro_pending = false;
                    image_ready = false;
                    stats_ready = bn_next;
                    break;
                }
                GemmArgs g{}; g.a = cur; g.c = dst; g.w = l.wf; g.M = (int)l.rows; g.N = l.c_out; g.K = 9 * l.c_in;
                g.act = l.d.activation; g.H = l.H; g.W = l.W; g.Ci = l.c_in; g.Co = l.c_out; g.splits = 1;
                launch_gemm<G_CONV_FWD>(s, g, launches_, exact_);
                break;
            }
            case N4J_LAYER_BATCHNORM: {
                const int cc = l.c_out;
                const bool vec = (cc % 4 == 0) && (256 % (cc / 4) == 0) && (cc / 4 <= 256);
                const bool feeds_tc = i + 1 < L_.size() && L_[i + 1].d.kind == N4J_LAYER_CONV3X3 && L_[i + 1].tc;
                if (stats_ready) {
                    stats_ready = false;
                } else if (vec) {
                    const int rows_per_iter = 256 / (cc / 4);
                    int grid = (int)std::max<long long>(1, std::min<long long>(74, (l.rows + 4 * rows_per_iter - 1) / (4 * rows_per_iter)));
                    LAUNCH(k_bn_stats, grid, 256, s, cur, l.rows, cc, l.d.eps, l.bn);
                } else {
                    LAUNCH(k_bn_stats_generic, cc, 256, s, cur, l.rows, cc, l.d.eps, l.bn);
                }
                if (feeds_tc && !(cfg_.flags & N4J_FLAG_NO_PROLOGUE_FUSION) && (l.d.activation == N4J_ACT_RELU || l.d.activation == N4J_ACT_IDENTITY)) {
                    // the BatchNorm apply (+ReLU, TF32 split, zero padding) runs as the operand prologue of the conv kernel itself
                    pro_pending = true;
                    pro_args = ConvTcArgs{};
                    pro_args.px = cur; pro_args.pgamma = params_.p + l.p_off; pro_args.pbeta = params_.p + l.p_off + cc; pro_args.pbn = l.bn;
                    pro_args.pact = l.d.activation; pro_args.pcompact = need_compact ? l.act : (float*)nullptr;
                    if (i == 0 && prearm_next_) { pro_args.prearm = 1; prearm_next_ = false; }
                    image_ready = true;
                } else if (feeds_tc) {
                    // BatchNorm-apply writes the next conv's operand image directly (hi/lo split, zero-padded, chunk-major);
                    // the compact activation is only needed by the backward pass
                    LAUNCH(k_bn_apply_to_image, grid_for(l.out_len / 4), 256, s, cur, params_.p + l.p_off, params_.p + l.p_off + cc, l.bn,
                           l.d.activation, geom_, L_[i + 1].img, need_compact ? l.act : (float*)nullptr);
                    image_ready = true;
                } else if (need_compact && fuse_last_bn_ok_ && i + 1 == L_.size() && cc == 64) {
                    // augmented evaluation: the last BatchNorm's apply is done by the first backward kernel (k_bn_bwd_stats64)
                    fwd_left_last_bn_ = true;
                } else if (leave_last_ && !need_compact && i + 1 == L_.size() && cc == 64 && l.bn.deferred == 1) {
                    // forward solve, stages 1-5: the stage combination that follows applies this BatchNorm and writes the K row
                    left_ = FuseLast{};
                    left_.mode = 1; left_.act = l.d.activation; left_.x = cur; left_.gamma = params_.p + l.p_off; left_.beta = params_.p + l.p_off + cc;
                    left_.bn = l.bn;
                } else {
                    if (cc == 64) LAUNCH(k_bn_apply64, grid_for(l.out_len / 4), 256, s, cur, dst, params_.p + l.p_off, params_.p + l.p_off + cc, l.bn, l.d.activation, l.rows);
                    else LAUNCH(k_bn_apply, grid_for(l.out_len), 256, s, cur, dst, params_.p + l.p_off, params_.p + l.p_off + cc, l.bn, l.d.activation, l.rows, cc);
                }
                break;
            }
            case N4J_LAYER_ACTIVATION:
                LAUNCH(k_act_fwd, grid_for(l.out_len), 256, s, cur, dst, l.d.activation, l.out_len);
                break;
        }
        cur = dst;
    }
}

// BackpropagateAdjoint.calculateDerivative :63-85 on the layout [z | a | theta | a_t]:
//   out.z <- f(z,t);  out.a <- (-a)^T df/dz;  out.theta <- (-a)^T df/dtheta;  out.a_t stays 0 (NoTimeInput.java:26-29)
void Engine::join_side(cudaStream_t s) {
    if (side_busy_) {
        cudaEvent_t ej = ev_pool_[(ev_next_++) % EV_POOL];
        N4J_CUDA(cudaEventRecord(ej, side_));
        N4J_CUDA(cudaStreamWaitEvent(s, ej, 0));
        side_busy_ = false;
    }
    slot_event_valid_[0] = slot_event_valid_[1] = false;
}

// slot: which activation / gradient buffer set this evaluation uses; defer_join: leave the weight-gradient kernels running on
// the side branch (the caller joins before it touches the theta block of the K rows).
void Engine::enqueue_mlp_small(cudaStream_t s, Ref in, Ref out, bool aug) {
    MlpSmallArgs m = mlp_args_;
    m.params = params_.p; m.in = in; m.out = out; m.ts = tspec_; m.write_t = (aug && aug_tslot_) ? 1 : 0;
    const dim3 grid((unsigned)((B_ + MLP_ROWS - 1) / MLP_ROWS)), block(MLP_THREADS);
    if (aug) launch_k(k_mlp_small<true>, grid, block, mlp_smem_, s, pdl_, m);
    else launch_k(k_mlp_small<false>, grid, block, mlp_smem_, s, pdl_, m);
    ++launches_;
}

void Engine::enqueue_rhs_aug(cudaStream_t s, Ref in, Ref out, bool bn1_ready, int slot, bool defer_join, bool last_of_trial) {
    if (mlp_fused_ && !comm_dev_) { begin_eval(); enqueue_mlp_small(s, in, out, true); return; }
    for (auto& l : L_) { l.act = l.act2[slot]; l.gin = l.gin2[slot]; }
    if (slot_event_valid_[slot]) {   // the evaluation that used this buffer set before must have finished its side work
        N4J_CUDA(cudaStreamWaitEvent(s, slot_event_[slot], 0));
        slot_event_valid_[slot] = false;
    }
    fwd_left_last_bn_ = false;
    fuse_last_bn_ok_ = true;
    begin_eval();
    in_aug_ = true;
    enqueue_rhs_fwd(s, in, out, true, bn1_ready);
    in_aug_ = false;
    fuse_last_bn_ok_ = false;
    Ref g = with_off(in, off_a_);
    float gscale = -1.f;                                 // zAdjoint().negi() :75
    const int last = (int)L_.size() - 1;
    bool gimg_ready = false;     // the BatchNorm above already wrote the operand image of this conv's output gradient
    bool bpro_pending = false;   // ... or left its backward to the dgrad kernel's operand prologue
    ConvTcArgs bpro_args{};
    bool forked = false;
    bool bstats_ready = false;   // the dgrad kernel behind this BatchNorm already reduced its backward statistics (EPI 2)
    for (int i = last; i >= 0; --i) {
        LayerPlan& l = L_[i];
        Ref x = i == 0 ? in : fixed_ref(L_[i - 1].act);
        Ref y = i == last ? out : fixed_ref(l.act);
        // every layer owns the buffer of the gradient w.r.t. its input: nothing is recycled inside one evaluation, so the
        // weight-gradient kernels can run on a side branch while the dgrad/BatchNorm chain moves on
        Ref dx = i == 0 ? with_off(out, off_a_) : fixed_ref(l.gin);
        Ref dth = with_off(out, off_theta_ + l.g_off);
        switch (l.d.kind) {
            case N4J_LAYER_BATCHNORM: {
                const int cc = l.c_out;
                int grid = (cc <= 256 && 256 % cc == 0) ? (int)std::max<long long>(1, std::min<long long>(74, (l.rows + 4 * (256 / cc) - 1) / (4 * (256 / cc)))) : 74;
                if (bstats_ready) {
                    bstats_ready = false;
                } else if (cc == 64) {
                    const int g64 = (int)std::max<long long>(1, std::min<long long>(148, (l.rows + 15) / 16));
                    LAUNCH(k_bn_bwd_stats64, g64, 256, s, g, gscale, y, x, l.bn, l.d.activation, l.rows, dth, with_off(dth, cc),
                           params_.p + l.p_off, params_.p + l.p_off + cc, (i == last && fwd_left_last_bn_) ? 1 : 0);
                } else {
                    LAUNCH(k_bn_bwd_stats, grid, 256, s, g, gscale, y, x, l.bn, l.d.activation, l.rows, cc, dth, with_off(dth, cc));
                }
                const bool feeds_tc = i > 0 && L_[i - 1].d.kind == N4J_LAYER_CONV3X3 && L_[i - 1].tc;
                if (feeds_tc && !(cfg_.flags & N4J_FLAG_NO_PROLOGUE_FUSION) && (l.d.activation == N4J_ACT_RELU || l.d.activation == N4J_ACT_IDENTITY)) {
                    // BatchNorm backward dx is produced by the operand prologue of the dgrad conv (and stored compactly for wgrad)
                    bpro_pending = true;
                    bpro_args = ConvTcArgs{};
                    bpro_args.px = x; bpro_args.pg = g; bpro_args.pgscale = gscale; bpro_args.pgamma = params_.p + l.p_off; bpro_args.pbeta = params_.p + l.p_off + cc;
                    bpro_args.pbn = l.bn; bpro_args.pact = l.d.activation; bpro_args.pcompact = l.gin;
                    bpro_args.pdgamma = dth; bpro_args.pdbeta = with_off(dth, cc);
                    gimg_ready = true;
                } else if (feeds_tc) {
                    LAUNCH(k_bn_bwd_dx_to_image, grid_for(l.in_len / 4), 256, s, g, gscale, y, x, params_.p + l.p_off, l.bn, l.d.activation, geom_,
                           l.gin, L_[i - 1].gimg, dth, with_off(dth, cc));
                    gimg_ready = true;
                } else if (leave_last_ && i == 0 && cc == 64 && l.bn.deferred == 1) {
                    // adjoint solve, stages 1-5: the stage combination that follows computes this dx and writes the a block of the K row
                    left_ = FuseLast{};
                    left_.mode = 2; left_.act = l.d.activation; left_.g = g; left_.gscale = gscale; left_.y = y; left_.gamma = params_.p + l.p_off;
                    left_.beta = params_.p + l.p_off + cc; left_.bn = l.bn; left_.a4 = off_a_ / 4; left_.dgamma_off = off_theta_ + l.g_off;
                } else {
                    if (cc == 64)
                        LAUNCH(k_bn_bwd_dx64, grid_for(l.in_len / 4), 256, s, g, gscale, y, x, params_.p + l.p_off, l.bn, l.d.activation, l.rows, dx, dth,
                               with_off(dth, cc));
                    else
                        LAUNCH(k_bn_bwd_dx, grid_for(l.in_len), 256, s, g, gscale, y, x, params_.p + l.p_off, l.bn, l.d.activation, l.rows, cc, dx, dth,
                               with_off(dth, cc), params_.p + l.p_off + cc);
                }
                break;
            }
            case N4J_LAYER_ACTIVATION:
                LAUNCH(k_act_bwd, grid_for(l.out_len), 256, s, g, gscale, y, l.d.activation, dx, l.out_len);
                break;
            case N4J_LAYER_TIME_CONCAT:
                // TimeInput.update (TimeInput.java:46-50): the state part is the z-adjoint derivative, the summed time part the
                // derivative of the time slot (dropped when the solve carries none: NoTimeGrad's empty tAdjoint)
                LAUNCH(k_time_concat_bwd, (int)std::min<long long>(grid_red(l.out_len / 4 + 1), 148), 256, s, g, gscale, dx, l.rows, l.c_in,
                       l.c_out - l.c_in, tcat_acc_, reinterpret_cast<unsigned int*>(tcat_misc_), comm_dev_, tcat_misc_ + 1,
                       with_off(out, off_t_), aug_tslot_ ? 1 : 0);
                break;
            case N4J_LAYER_DENSE:
            case N4J_LAYER_CONV3X3: {
                Ref d = g;
                // identity-activation dense layer on the generic GEMM path whose epsilon is -a: the sign rides on the GEMMs' stores
                const bool neg = l.d.kind == N4J_LAYER_DENSE && l.d.activation == N4J_ACT_IDENTITY && gscale == -1.f && !l.exact && !l.tcd;
                if (!neg && (l.d.activation != N4J_ACT_IDENTITY || gscale != 1.f)) {
                    d = fixed_ref(gb_[2].p);
                    LAUNCH(k_act_bwd, grid_for(l.out_len), 256, s, g, gscale, y, l.d.activation, d, l.out_len);
                    gimg_ready = false;
                }
                if (l.d.kind == N4J_LAYER_DENSE) {
                    const float* Wt = params_.p + l.p_off;
                    if (l.exact) {
                        LAUNCH(k_dense_wgrad_exact, (int)std::min<long long>(l.g_len, 148 * 8), 128, s, x, d, dth, l.rows, l.c_in, l.c_out, l.d.has_bias);
                        if (comm_dev_) enqueue_vec_allreduce(s, dth, l.g_len);
                        LAUNCH(k_dense_dgrad_exact, grid_for(l.in_len), 256, s, d, Wt, dx, l.rows, l.c_in, l.c_out);
                    } else if (l.tcd) {
                        LAUNCH(k_rowmajor_to_image, grid_for(l.out_len / 4), 256, s, d, 1.f, l.dimg, l.rows, l.c_out, l.img_rows);
                        // dW^T[o,i] = sum_b delta[b,o] x[b,i]: K-major GEMM over the batch on the transposed images
                        LAUNCH(k_rowmajor_to_image_T, grid_for((l.img_rows / 4) * l.c_out), 256, s, d, l.dTimg, l.rows, l.c_out, l.img_rows,
                               (long long)((l.c_out + 127) / 128 * 128), 0);
                        LAUNCH(k_rowmajor_to_image_T, grid_for((l.img_rows / 4) * l.c_in), 256, s, x, l.xTimg, l.rows, l.c_in, l.img_rows, 0LL, 1);
                        GemmTcArgs wa{};
                        wa.M = l.c_out; wa.N = l.c_in; wa.kblocks = (int)(l.img_rows / 32); wa.a_nch = (int)(l.img_rows / 4); wa.out = dth;
                        wa.a_img = l.dTimg; wa.a_rows = (long long)((l.c_out + 127) / 128 * 128); wa.b_img = l.xTimg; wa.b_nch = (int)(l.img_rows / 4);
                        launch_gemm_tc(s, l.tm_dT, l.tm_xT, wa);
                        if (l.d.has_bias) LAUNCH(k_colsum, (l.c_out + 31) / 32, 32 * COLSUM_RG, s, d, with_off(dth, (long long)l.c_in * l.c_out), l.rows, l.c_out, 0);
                        if (comm_dev_) enqueue_vec_allreduce(s, dth, l.g_len);
                        GemmTcArgs qa{};     // dx[b,i] = sum_o delta[b,o] W(i,o)
                        qa.M = (int)l.rows; qa.N = l.c_in; qa.kblocks = l.c_out / 32; qa.a_nch = l.c_out / 4; qa.out = dx;
                        qa.a_img = l.dimg; qa.a_rows = l.img_rows; qa.b_img = l.wsb; qa.b_nch = l.c_out / 4;
                        launch_gemm_tc(s, l.tm_d_k, l.tm_wb, qa);
                    } else {
                        GemmArgs w{}; w.a = d; w.b = x; w.c = dth; w.M = l.c_out; w.N = l.c_in; w.K = (int)l.rows; w.splits = 1; w.negate = neg ? 1 : 0;
                        const long long tiles = (long long)((l.c_out + 63) / 64) * ((l.c_in + 63) / 64);
                        if (!exact_ && tiles < 148 && l.rows >= 256 && wpart_.n >= 8LL * l.c_in * l.c_out) {   // few output tiles, long K: split the batch
                            w.splits = 8; w.c = fixed_ref(wpart_.p);
                        }
                        launch_gemm<G_DENSE_WGRAD>(s, w, launches_, exact_);
                        if (w.splits > 1) LAUNCH(k_splitk_reduce, grid_for((long long)l.c_in * l.c_out), 256, s, wpart_.p, dth, (long long)l.c_in * l.c_out, 8);
                        if (l.d.has_bias) LAUNCH(k_colsum, (l.c_out + 31) / 32, 32 * COLSUM_RG, s, d, with_off(dth, (long long)l.c_in * l.c_out), l.rows, l.c_out, neg ? 1 : 0);
                        if (comm_dev_) enqueue_vec_allreduce(s, dth, l.g_len);
                        GemmArgs q{}; q.a = d; q.c = dx; q.w = Wt; q.M = (int)l.rows; q.N = l.c_in; q.K = l.c_out; q.splits = 1; q.negate = neg ? 1 : 0;
                        launch_gemm<G_DENSE_DGRAD>(s, q, launches_, exact_);
                    }
                } else if (l.tc) {
                    // dgrad on the main chain.  With the fused prologue the dgrad kernel also materialises `d` (the BatchNorm
                    // backward of the layer behind this conv) compactly, so the weight gradient forks AFTER it.
                    if (!gimg_ready) LAUNCH(k_compact_to_image, grid_for(l.out_len / 4), 256, s, d, geom_, l.gimg);
                    // ... and, when a 64-channel BatchNorm sits in front of the conv, its backward statistics in the epilogue
                    int epi = 0;
                    if (i > 0 && !(cfg_.flags & N4J_FLAG_NO_BWD_STATS_FUSION)) {
                        const LayerPlan& b = L_[i - 1];
                        if (b.d.kind == N4J_LAYER_BATCHNORM && b.c_out == 64 && (b.d.activation == N4J_ACT_RELU || b.d.activation == N4J_ACT_IDENTITY)) {
                            if (!bpro_pending) bpro_args = ConvTcArgs{};
                            bpro_args.ex = i - 1 == 0 ? in : fixed_ref(L_[i - 2].act);
                            bpro_args.egamma = params_.p + b.p_off; bpro_args.ebeta = params_.p + b.p_off + 64;
                            bpro_args.eact = b.d.activation;
                            bpro_args.edgamma = with_off(out, off_theta_ + b.g_off);
                            bpro_args.edbeta = with_off(out, off_theta_ + b.g_off + 64);
                            bpro_args.bn = b.bn; bpro_args.bn_rows = b.rows;
                            epi = 2;
                            bstats_ready = true;
                        }
                    }
                    launch_conv_tc(s, l.gimg, l.wimg_b, dx, N4J_ACT_IDENTITY, nullptr, (bpro_pending || epi) ? &bpro_args : nullptr, bpro_pending ? 2 : 0, epi);
                    bpro_pending = false;
                    // weight gradient on the side branch (fork)
                    cudaEvent_t ef = ev_pool_[(ev_next_++) % EV_POOL];
                    N4J_CUDA(cudaEventRecord(ef, s));
                    N4J_CUDA(cudaStreamWaitEvent(side_, ef, 0));
                    // at most 64 CTAs: they share the chip with the CTAs of the main chain
                    // ... except the very last weight gradient of a trial step: the main chain has nothing heavy left and waits for
                    // the join, so that launch takes one image per CTA (half the latency)
                    const bool wide = last_of_trial && i == first_conv_;
                    const int ipc = wide ? 1 : (int)((B_ + 63) / 64);
                    int wg_ctas = (int)((B_ + ipc - 1) / ipc);
                    if (wgrad_tc_) {      // tcgen05: one CTA per 128 padded pixels
                        wg_ctas = wgeom_.tiles;
                        k_conv_wgrad_tc<<<dim3(wgeom_.tiles), WGRAD_TC_THREADS, wgrad_tc_smem_, side_>>>(x, d, wgeom_, wpart_.p, (long long*)nullptr);
                    } else {
                        k_conv_wgrad64<<<wg_ctas, 256, wgrad_smem_, side_>>>(x, d, geom_, wpart_.p, ipc, wgrad_nbuf_);
                    }
                    ++launches_;
                    // sharded path: the sum over the ranks rides on the reduction (LL exchange of every finished float4)
                    const bool ll = comm_dev_ && vec_allreduce_ll_ && l.g_len <= comm_xll_cap_ && !skip_vec_allreduce_;
                    LAUNCH(k_wgrad_reduce, 144, 256, side_, wpart_.p, dth, wg_ctas, ll ? comm_dev_ : (CommDev*)nullptr);
                    if (comm_dev_ && !ll) enqueue_vec_allreduce(side_, dth, l.g_len);
                    forked = true;
                    side_busy_ = true;
                } else {
                    GemmArgs w{}; w.a = x; w.b = d; w.c = fixed_ref(wpart_.p); w.M = 9 * l.c_in; w.N = l.c_out; w.K = (int)l.rows;
                    w.H = l.H; w.W = l.W; w.Ci = l.c_in; w.Co = l.c_out; w.splits = 16;
                    if (exact_ && l.H * l.W <= 1024) {      // parity mode: same sums in the same order on 16 x 16 output tiles (144 blocks instead of 9)
                        w.c = dth; w.splits = 1;
                        launch_k(k_conv_wgrad_exact, dim3((w.N + 15) / 16, (w.M + 15) / 16), dim3(256), 0, s, true, w);
                        ++launches_;
                    } else {
                        if (exact_) w.c = dth;
                        launch_gemm<G_CONV_WGRAD>(s, w, launches_, exact_);
                    }
                    if (!exact_) LAUNCH(k_splitk_reduce, grid_for(l.g_len), 256, s, wpart_.p, dth, l.g_len, 16);
                    if (comm_dev_) enqueue_vec_allreduce(s, dth, l.g_len);
                    GemmArgs q{}; q.a = d; q.c = dx; q.w = l.wb; q.M = (int)l.rows; q.N = l.c_in; q.K = 9 * l.c_out;
                    q.H = l.H; q.W = l.W; q.Ci = l.c_in; q.Co = l.c_out; q.splits = 1;
                    launch_gemm<G_CONV_DGRAD>(s, q, launches_, exact_);
                }
                gimg_ready = false;
                break;
            }
        }
        g = dx; gscale = 1.f;
    }
    if (forked) {
        if (defer_join) {
            N4J_CUDA(cudaEventRecord(slot_event_[slot], side_));
            slot_event_valid_[slot] = true;
        } else {
            join_side(s);
        }
    }
    for (auto& l : L_) { l.act = l.act2[0]; l.gin = l.gin2[0]; }
}

// =====================================================================================================
// the solve: prologue (K0, initial step), trial body (6 stages + error/controller), epilogue
// =====================================================================================================
void Engine::enqueue_prologue(cudaStream_t s, bool aug, long long n4, const SolverParams& sp, bool dense_out, int nt) {
    Ref ycur{Y_, &ctrl_->y_cur, n_alloc_, 0, 0};
    Ref ywork{Y_, &ctrl_->y_cur, n_alloc_, 0, 1};
    set_bn_slot(0); eval_begun_ = false;                // every solve starts on statistics set 1 (see set_bn_slot)
    LAUNCH(k_solve_begin, 1, 32, s, ctrl_, tpair_);
    if (dense_out) LAUNCH(k_dense_edge, grid_for(nz_), 256, s, ctrl_, Y_, n_alloc_, nz_, wanted_, 0, 0, zt_.p + nz_, nz_);
    tspec_ = TimeSpec{0, 0.f, 0.f};                                                  // time + timeOffset(0)
    enqueue_rhs(s, aug, ycur, Ref{K_, &ctrl_->krow[0], n_alloc_, 0, 0});            // AdaptiveRungeKuttaSolver.java:131
    LAUNCH(k_init_sums, grid_red(n4), 256, s, ctrl_, Y_, K_, n_alloc_, n4, sp);
    set_bn_slot(bn_slot_ ^ 1); eval_begun_ = true;      // the combination below already produces statistics of the NEXT evaluation
    const BnFuse bf = bn_fuse();
    LAUNCH(k_stage_combine<0>, bf.enabled ? grid_red(n4) : grid_for(n4), 256, s, ctrl_, Y_, K_, n_alloc_, 0LL, n4, bf);  // equation.step(ones(1), step)
    tspec_ = TimeSpec{1, 0.f, 0.f};                                                  // timeOffset = the first guess (equation.step above)
    enqueue_rhs(s, aug, ywork, Ref{K_, &ctrl_->krow[1], n_alloc_, 0, 0}, bf.enabled != 0);   // policy :112
    LAUNCH(k_init_final, grid_red(n4), 256, s, ctrl_, Y_, K_, n_alloc_, n4, sp, (cudaGraphConditionalHandle)0, 0);
    (void)nt;
}

template <int S>
static void launch_combine(cudaStream_t s, int grid, Ctrl* c, float* Y, float* K, long long stride, long long i0, long long i1, const BnFuse& bf) {
    launch_k(k_stage_combine<S>, dim3(grid), dim3(256), 0, s, true, c, Y, K, stride, i0, i1, bf);
}

void Engine::enqueue_trial(cudaStream_t s, bool aug, long long n4, const SolverParams& sp, bool dense_out, int nt,
                           cudaGraphConditionalHandle cond, int use_cond) {
    Ref ywork{Y_, &ctrl_->y_cur, n_alloc_, 0, 1};
    set_bn_slot(0); eval_begun_ = false;                // six evaluations: sets 1,0,1,0,1,0 — the body ends where it started
    const BnFuse bf0 = bn_fuse();
    // Adjoint solves: f and the vjp only read the z and a blocks of the stage state, so stages 1-6 combine just those
    // (n4_main); the theta block of y1 is combined once, right before the error norm, after the side branch (weight
    // gradients of all six stages) has been joined.
    const bool split = aug && side_capable_;
    const long long n4_main = split ? (2 * nz_pad_) / 4 : n4;
    const int grid = bf0.enabled ? grid_red(n4_main) : grid_for(n4_main);
    BnFuse nobf{};
    // Stages 1-5 of a 64-channel BatchNorm-first / BatchNorm-last block on one GPU: the evaluation leaves its last elementwise kernel
    // to the combination of the next stage (k_stage_combine_last).  The adjoint form needs the split layout (z | a combined here, theta late).
    const bool can_leave = combine_fusion_ && bf0.enabled && bf0.sc.deferred == 1 && (!comm_dev_ || mgpu_deferred_) && nz_ % 64 == 0 && off_a_ % 4 == 0 &&
                           L_.size() >= 2 && L_.back().d.kind == N4J_LAYER_BATCHNORM && L_.back().c_out == 64 &&
                           (!aug || (split && L_[1].d.kind == N4J_LAYER_CONV3X3 && L_[1].tc && !(cfg_.flags & N4J_FLAG_NO_PROLOGUE_FUSION) &&
                                     (L_[0].d.activation == N4J_ACT_RELU || L_[0].d.activation == N4J_ACT_IDENTITY)));   // the operand prologue of that conv takes the re-arm
    const int grid_last = grid_red(bf0.nz4);
    left_ = FuseLast{};
    for (int st = 1; st <= 6; ++st) {
        set_bn_slot(bn_slot_ ^ 1); eval_begun_ = true;  // the combination produces the first BatchNorm's statistics of this stage's evaluation
        const BnFuse bf = bn_fuse();
        if (left_.mode == 1) {
            switch (st) {
                case 2: launch_k(k_stage_combine_last<2, 1>, dim3(grid_last), dim3(256), 0, s, true, ctrl_, Y_, K_, n_alloc_, bf, left_); break;
                case 3: launch_k(k_stage_combine_last<3, 1>, dim3(grid_last), dim3(256), 0, s, true, ctrl_, Y_, K_, n_alloc_, bf, left_); break;
                case 4: launch_k(k_stage_combine_last<4, 1>, dim3(grid_last), dim3(256), 0, s, true, ctrl_, Y_, K_, n_alloc_, bf, left_); break;
                case 5: launch_k(k_stage_combine_last<5, 1>, dim3(grid_last), dim3(256), 0, s, true, ctrl_, Y_, K_, n_alloc_, bf, left_); break;
                case 6: launch_k(k_stage_combine_last<6, 1>, dim3(grid_last), dim3(256), 0, s, true, ctrl_, Y_, K_, n_alloc_, bf, left_); break;
                default: throw Error{N4J_ERR_ILLEGAL_STATE, "stage 1 has no evaluation in front of it"};
            }
        } else if (left_.mode == 2) {
            switch (st) {
                case 2: launch_k(k_stage_combine_last<2, 2>, dim3(grid_last), dim3(256), 0, s, true, ctrl_, Y_, K_, n_alloc_, bf, left_); break;
                case 3: launch_k(k_stage_combine_last<3, 2>, dim3(grid_last), dim3(256), 0, s, true, ctrl_, Y_, K_, n_alloc_, bf, left_); break;
                case 4: launch_k(k_stage_combine_last<4, 2>, dim3(grid_last), dim3(256), 0, s, true, ctrl_, Y_, K_, n_alloc_, bf, left_); break;
                case 5: launch_k(k_stage_combine_last<5, 2>, dim3(grid_last), dim3(256), 0, s, true, ctrl_, Y_, K_, n_alloc_, bf, left_); break;
                case 6: launch_k(k_stage_combine_last<6, 2>, dim3(grid_last), dim3(256), 0, s, true, ctrl_, Y_, K_, n_alloc_, bf, left_); break;
                default: throw Error{N4J_ERR_ILLEGAL_STATE, "stage 1 has no evaluation in front of it"};
            }
            prearm_next_ = true;     // that kernel still read the set its block 0 would have re-armed
        } else
        switch (st) {
            case 1: launch_combine<1>(s, grid, ctrl_, Y_, K_, n_alloc_, 0, n4_main, bf); break;
            case 2: launch_combine<2>(s, grid, ctrl_, Y_, K_, n_alloc_, 0, n4_main, bf); break;
            case 3: launch_combine<3>(s, grid, ctrl_, Y_, K_, n_alloc_, 0, n4_main, bf); break;
            case 4: launch_combine<4>(s, grid, ctrl_, Y_, K_, n_alloc_, 0, n4_main, bf); break;
            case 5: launch_combine<5>(s, grid, ctrl_, Y_, K_, n_alloc_, 0, n4_main, bf); break;
            case 6: launch_combine<6>(s, grid, ctrl_, Y_, K_, n_alloc_, 0, n4_main, bf); break;
        }
        ++launches_;
        left_ = FuseLast{};
        leave_last_ = can_leave && st < 6;
        // quirk A.3: State.step stores the FULL step in timeOffset, so every stage is evaluated at t + h
        // (FirstOrderEquationWithState.java:45,82); N4J_FLAG_EXACT_STAGE_TIMES uses the tableau's c_s instead
        static const double kC[6] = {1.0 / 5.0, 3.0 / 10.0, 4.0 / 5.0, 8.0 / 9.0, 1.0, 1.0};
        tspec_ = (cfg_.flags & N4J_FLAG_EXACT_STAGE_TIMES) ? TimeSpec{2, (float)kC[st - 1], 0.f} : TimeSpec{1, 0.f, 0.f};
        if (aug) enqueue_rhs_aug(s, ywork, Ref{K_, &ctrl_->krow[st], n_alloc_, 0, 0}, bf.enabled != 0, st & 1, split, st == 6);
        else enqueue_rhs_fwd(s, ywork, Ref{K_, &ctrl_->krow[st], n_alloc_, 0, 0}, false, bf.enabled != 0);
        leave_last_ = false;
        if (prearm_next_) throw Error{N4J_ERR_ILLEGAL_STATE, "no operand prologue took the statistics re-arm"};
    }
    if (left_.mode) throw Error{N4J_ERR_ILLEGAL_STATE, "stage 6 left work behind"};
    if (split) {
        join_side(s);
        launch_combine<6>(s, grid_for(n4 - n4_main), ctrl_, Y_, K_, n_alloc_, n4_main, n4, nobf);   // theta | a_t block of y1
        ++launches_;
    }
    DenseOutArgs dout{dense_out ? wanted_ : nullptr, dense_out ? nt - 1 : 0};
    LAUNCH(k_error_ctrl, grid_red(n4), 256, s, ctrl_, Y_, K_, n_alloc_, n4, sp, log_, dout, cond, use_cond);
    if (dense_out) LAUNCH(k_dense_output, grid_for(nz_), 256, s, ctrl_, Y_, K_, n_alloc_, nz_, wanted_, zt_.p + nz_, nz_);
}

void Engine::enqueue_epilogue(cudaStream_t s, bool dense_out, int nt) {
    if (dense_out) LAUNCH(k_dense_edge, grid_for(nz_), 256, s, ctrl_, Y_, n_alloc_, nz_, wanted_, nt - 2, 1, zt_.p + nz_, nz_);
}

SolveGraph& Engine::get_graph(bool aug, int tslot, bool dense_out, int nt) {
    const long long key = (aug ? 1 : 0) | (tslot ? 2 : 0) | (dense_out ? 4 : 0) | ((long long)(dense_out ? nt : 0) << 8);
    auto it = graphs_.find(key);
    if (it != graphs_.end()) return it->second;
    aug_tslot_ = tslot != 0;
    SolveGraph sg;
    const long long n4 = (aug ? n_alloc_ : nz_pad_) / 4;
    const long long n_norm = aug ? (2 * nz_ * world_ + n_params_ + (tslot ? 1 : 0)) : nz_ * world_;
    SolverParams sp = solver_params(n_norm);
    sp.rep_begin4 = aug ? (2 * nz_pad_) / 4 : n4;
    sp.count_rep = (rank_id_ == 0) ? 1 : 0;
    const int saved = launches_;

    // Measured on B200: inside the captured graph programmatic edges do not pay (6.97 ms vs 6.67 ms per step) — graph
    // replay already hides the launch latency and early-resident dependents only take SM slots — so PDL is used for the
    // eager paths (host-driven loop, RHS probes) and switched off while capturing.
    const bool saved_pdl = t_pdl;
    t_pdl = getenv("N4J_GRAPH_PDL") != nullptr && saved_pdl;      // A/B switch: programmatic edges inside the captured graph (re-measured in r02)
    // A throw inside the capture (an ILLEGAL_STATE check of enqueue_trial, any CUDA failure) must not leave stream2_ capturing,
    // leak the graph or lose the launch counter / PDL switch: every later call on the handle would fail.
    struct CaptureGuard {
        Engine* e; SolveGraph* g; cudaStream_t s; int launches; bool pdl; bool armed = true;
        ~CaptureGuard() {
            if (!armed) return;
            cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
            if (cudaStreamIsCapturing(s, &st) == cudaSuccess && st != cudaStreamCaptureStatusNone) {
                cudaGraph_t dead = nullptr;
                cudaStreamEndCapture(s, &dead);      // ends (and invalidates) the capture; a captured-to-graph body belongs to g->graph
                if (dead && dead != g->graph) cudaGraphDestroy(dead);
            }
            if (g->graph) { cudaGraphDestroy(g->graph); g->graph = nullptr; }
            cudaGetLastError();
            e->launches_ = launches;
            t_pdl = pdl;
        }
    } guard{this, &sg, stream2_, saved, saved_pdl};
    N4J_CUDA(cudaGraphCreate(&sg.graph, 0));
    // prologue as a child graph
    cudaGraph_t pro = nullptr, epi = nullptr;
    launches_ = 0;
    N4J_CUDA(cudaStreamBeginCapture(stream2_, cudaStreamCaptureModeRelaxed));
    enqueue_prologue(stream2_, aug, n4, sp, dense_out, nt);
    N4J_CUDA(cudaStreamEndCapture(stream2_, &pro));
    sg.prologue_launches = launches_;
    cudaGraphNode_t pro_node = nullptr;
    N4J_CUDA(cudaGraphAddChildGraphNode(&pro_node, sg.graph, nullptr, 0, pro));

    // WHILE node: body = one trial step; the controller (last block of k_error_ctrl) sets the condition
    cudaGraphConditionalHandle cond;
    N4J_CUDA(cudaGraphConditionalHandleCreate(&cond, sg.graph, 1, cudaGraphCondAssignDefault));
    cudaGraphNodeParams cp = {cudaGraphNodeTypeConditional};
    cp.type = cudaGraphNodeTypeConditional;
    cp.conditional.handle = cond;
    cp.conditional.type = cudaGraphCondTypeWhile;
    cp.conditional.size = 1;
    cudaGraphNode_t while_node = nullptr;
    N4J_CUDA(cudaGraphAddNode(&while_node, sg.graph, &pro_node, 1, &cp));
    cudaGraph_t body = cp.conditional.phGraph_out[0];
    launches_ = 0;
    N4J_CUDA(cudaStreamBeginCaptureToGraph(stream2_, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
    enqueue_trial(stream2_, aug, n4, sp, dense_out, nt, cond, 1);
    N4J_CUDA(cudaStreamEndCapture(stream2_, nullptr));
    sg.body_launches = launches_;

    launches_ = 0;
    if (dense_out) {
        N4J_CUDA(cudaStreamBeginCapture(stream2_, cudaStreamCaptureModeRelaxed));
        enqueue_epilogue(stream2_, dense_out, nt);
        N4J_CUDA(cudaStreamEndCapture(stream2_, &epi));
        cudaGraphNode_t epi_node = nullptr;
        N4J_CUDA(cudaGraphAddChildGraphNode(&epi_node, sg.graph, &while_node, 1, epi));
    }
    sg.epilogue_launches = launches_;
    launches_ = saved;
    t_pdl = saved_pdl;
    N4J_CUDA(cudaGraphInstantiate(&sg.exec, sg.graph, 0));
    guard.armed = false;
    if (pro) cudaGraphDestroy(pro);
    if (epi) cudaGraphDestroy(epi);
    return graphs_.emplace(key, sg).first->second;
}

// One FirstOrderSolver.integrate(): state in Y[y_cur], time pair taken from tpair_[ctrl.seg_index++].
void Engine::run_solve(bool aug, int tslot, bool dense_out, int nt) {
    if (!(cfg_.flags & N4J_FLAG_NO_GRAPH)) {
        SolveGraph& sg = get_graph(aug, tslot, dense_out, nt);
        N4J_CUDA(cudaGraphLaunch(sg.exec, stream_));
        pending_graph_runs_.push_back({&sg, 0});
        return;
    }
    // debug path: the host drives the loop and reads `done` after every trial
    aug_tslot_ = tslot != 0;
    const long long n4 = (aug ? n_alloc_ : nz_pad_) / 4;
    const long long n_norm = aug ? (2 * nz_ * world_ + n_params_ + (tslot ? 1 : 0)) : nz_ * world_;
    SolverParams sp = solver_params(n_norm);
    sp.rep_begin4 = aug ? (2 * nz_pad_) / 4 : n4;
    sp.count_rep = (rank_id_ == 0) ? 1 : 0;
    enqueue_prologue(stream_, aug, n4, sp, dense_out, nt);
    for (;;) {
        N4J_CUDA(cudaMemcpyAsync(ctrl_host_, ctrl_, sizeof(Ctrl), cudaMemcpyDeviceToHost, stream_));
        N4J_CUDA(cudaStreamSynchronize(stream_));
        if (ctrl_host_->done) break;
        enqueue_trial(stream_, aug, n4, sp, dense_out, nt, (cudaGraphConditionalHandle)0, 0);
    }
    enqueue_epilogue(stream_, dense_out, nt);
}

__global__ void k_reset_ctrl(Ctrl* c) {
    pdl_begin();
    if (threadIdx.x == 0) {
        c->y_cur = 0; c->status = 0; c->log_n = 0; c->seg_index = 0;
        c->nfe = 0; c->accepted = 0; c->rejected = 0; c->solves = 0; c->trials = 0;
        c->ticket[0] = c->ticket[1] = c->ticket[2] = c->ticket[3] = 0u;
        c->done = 0; c->err = 0.f; c->h = 0.f;
    }
}

void Engine::reset_ctrl() {
    t_pdl = pdl_;
    launches_ = 0;
    pending_graph_runs_.clear();
    N4J_CUDA(cudaEventRecord(ev0_, stream_));
    LAUNCH(k_reset_ctrl, 1, 32, stream_, ctrl_);
    bn_rearm_all(stream_);     // every API call starts from clean statistics sets, whatever a failed call may have left behind
}

void Engine::set_time_pairs(const std::vector<float>& pairs) {
    if ((int)pairs.size() / 2 > max_pairs_) throw Error{N4J_ERR_UNSUPPORTED, "too many time points"};
    std::memcpy(tpair_host_, pairs.data(), sizeof(float) * pairs.size());
    N4J_CUDA(cudaMemcpyAsync(tpair_, tpair_host_, sizeof(float) * pairs.size(), cudaMemcpyHostToDevice, stream_));
}

void Engine::finish_call(n4j_stats* st, int) {
    N4J_CUDA(cudaMemcpyAsync(ctrl_host_, ctrl_, sizeof(Ctrl), cudaMemcpyDeviceToHost, stream_));
    N4J_CUDA(cudaEventRecord(ev1_, stream_));
    N4J_CUDA(cudaStreamSynchronize(stream_));
    const Ctrl& c = *ctrl_host_;
    log_count_ = c.log_n;
    int launches = launches_;
    if (!pending_graph_runs_.empty()) {
        SolveGraph* sg = pending_graph_runs_[0].first;
        launches += (int)pending_graph_runs_.size() * (sg->prologue_launches + sg->epilogue_launches) + (int)c.trials * sg->body_launches;
    }
    if (st) {
        st->nfe = c.nfe; st->accepted = c.accepted; st->rejected = c.rejected; st->solves = c.solves;
        st->last_h = c.h; st->last_err = c.err;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev0_, ev1_);
        st->gpu_ms = ms;
        st->gpu_launches = launches;
    }
    if (comm_dev_) {      // a peer timeout anywhere in the call (statistics, weight-gradient or error-norm exchange) is an error, never a silent wrong sum
        int comm_status = 0;
        N4J_CUDA(cudaMemcpy(&comm_status, reinterpret_cast<const char*>(comm_dev_) + offsetof(CommDev, status), sizeof(int), cudaMemcpyDeviceToHost));
        if (comm_status != 0 || c.status == N4J_ERR_COMM)
            throw Error{N4J_ERR_COMM, "multi-GPU exchange timed out: a peer rank did not publish its contribution within the spin limit"};
    }
    if (c.status == N4J_ERR_ILLEGAL_STATE) throw Error{N4J_ERR_ILLEGAL_STATE, "Detected NaN!"};   // NanWatchSolver.java:22
    if (c.status == N4J_ERR_MAX_STEPS) throw Error{N4J_ERR_MAX_STEPS, "step guard tripped: " + std::to_string(cfg_.max_trials) + " trial steps in one solve"};
}

// =====================================================================================================
// multi-GPU setup: exchange buffers in peer-mapped memory (CUDA IPC, one process per GPU)
// =====================================================================================================
// all-reduce (sum over the ranks) of a parameter-gradient slot, in place
void Engine::enqueue_vec_allreduce(cudaStream_t s, Ref slot, long long n) {
    if (skip_vec_allreduce_) return;      // N4J_DBG_SKIP_VEC_ALLREDUCE: timing experiments only (wrong gradients)
    const bool aligned = (slot.off % 4 == 0) && (slot.stride % 4 == 0);
    if (vec_allreduce_ll_ && n % 4 == 0 && n <= comm_xll_cap_ && aligned)
        LAUNCH(k_comm_allreduce_vec_ll, (int)std::max<long long>(1, std::min<long long>(296, (n / 4 + 255) / 256)), 256, s, comm_dev_, slot, n);
    else
        LAUNCH(k_comm_allreduce_vec, comm_grid(n), 256, s, comm_dev_, slot, n);
}
int Engine::comm_grid(long long n) const { return (int)std::max<long long>(1, std::min<long long>(64, (n + 511) / 512)); }

void Engine::comm_export(void* handle64) {
    N4J_CUDA(cudaSetDevice(device_));
    if (!comm_region_) {
        long long xcap = 4;
        for (auto& l : L_) xcap = std::max(xcap, round_up4(l.g_len));
        comm_xcap_ = xcap;
        comm_bytes_ = (size_t)COMM_OFF_XBUF + sizeof(float) * 2 * (size_t)xcap;
        // LL cells for parameter-gradient slots up to 256k floats (8 bytes per float and source rank, two parities); longer
        // slots (config #3's 4096 x 4096 layers) keep the bandwidth-efficient pull exchange through xbuf
        comm_xll_cap_ = std::min<long long>(xcap, 1LL << 18);
        comm_xll_off_ = (long long)comm_bytes_;
        comm_bytes_ += (size_t)2 * COMM_MAX_WORLD * (size_t)comm_xll_cap_ * 8;
        // BatchNorm statistics mailboxes of the sharded deferred scheme: 64 KB per deferred-capable layer
        comm_bytes_ = (comm_bytes_ + 255) & ~(size_t)255;
        comm_stat_off_ = (long long)comm_bytes_;
        long long mb = 0;
        for (auto& l : L_)
            if (l.d.kind == N4J_LAYER_BATCHNORM && l.bn_deferred_ok) { l.bn_mb = mb; mb += 2LL * 2 * COMM_MAX_WORLD * 128 * 16; }
        comm_bytes_ += (size_t)mb;
        N4J_CUDA(cudaMalloc(&comm_region_, comm_bytes_));
        N4J_CUDA(cudaMemset(comm_region_, 0, comm_bytes_));
    }
    cudaIpcMemHandle_t hnd;
    N4J_CUDA(cudaIpcGetMemHandle(&hnd, comm_region_));
    static_assert(sizeof(hnd) == 64, "cudaIpcMemHandle_t is 64 bytes");
    std::memcpy(handle64, &hnd, 64);
}

void Engine::comm_init(int rank, int world, const void* handles, long long global_batch) {
    N4J_CUDA(cudaSetDevice(device_));
    if (world < 1 || world > COMM_MAX_WORLD || rank < 0 || rank >= world) throw Error{N4J_ERR_INVALID_ARGUMENT, "bad rank/world (1..8 ranks)"};
    if (global_batch != B_ * world) throw Error{N4J_ERR_INVALID_ARGUMENT, "global batch must be world x the local batch (equal row shards)"};
    if (!comm_region_) throw Error{N4J_ERR_ILLEGAL_STATE, "node4j_comm_export must be called first"};
    for (auto& l : L_)
        if (l.d.kind == N4J_LAYER_BATCHNORM && (l.c_out > 128 || l.c_out % 4 != 0 || 256 % (l.c_out / 4) != 0))
            throw Error{N4J_ERR_UNSUPPORTED, "sharded BatchNorm needs a channel count that is a multiple of 4, divides 1024 and is <= 128"};
    CommDev cd;
    std::memset(&cd, 0, sizeof(cd));
    cd.rank = rank; cd.world = world; cd.xcap = comm_xcap_; cd.xll_off = comm_xll_off_; cd.xll_cap = comm_xll_cap_; cd.stat_off = comm_stat_off_;
    for (int r = 0; r < world; ++r) {
        if (r == rank) { cd.peer[r] = (char*)comm_region_; continue; }
        cudaIpcMemHandle_t hnd;
        std::memcpy(&hnd, (const char*)handles + 64 * r, 64);
        void* p = nullptr;
        N4J_CUDA(cudaIpcOpenMemHandle(&p, hnd, cudaIpcMemLazyEnablePeerAccess));
        cd.peer[r] = (char*)p;
        peer_ptrs_.push_back(p);
    }
    if (!comm_dev_) N4J_CUDA(cudaMalloc(&comm_dev_, sizeof(CommDev)));
    N4J_CUDA(cudaMemcpy(comm_dev_, &cd, sizeof(CommDev), cudaMemcpyHostToDevice));
    world_ = world; rank_id_ = rank;
    if (world > 1) {
        N4J_CUDA(cudaMemcpy(&ctrl_->comm, &comm_dev_, sizeof(CommDev*), cudaMemcpyHostToDevice));
        const char* md = getenv("N4J_MGPU_DEFERRED");
        mgpu_deferred_ = !(md && atoi(md) == 0);
        for (auto& l : L_) {
            if (l.d.kind != N4J_LAYER_BATCHNORM) continue;
            l.bn.cm = comm_dev_; l.bn.rows_total = l.rows * world;
            l.bn.inv_rows = 1.0 / (double)l.bn.rows_total;
            if (mgpu_deferred_ && l.bn_deferred_ok && l.bn_pool && l.bn_mb >= 0) {
                // sharded deferred statistics: producers accumulate local sums only, consumers exchange and finalise (layer_kernels.cuh: BnScratch)
                l.bn.deferred = 1;
                if (!l.bn_seq) {
                    N4J_CUDA(cudaMalloc(&l.bn_seq, sizeof(unsigned int) * 4));
                    N4J_CUDA(cudaMemset(l.bn_seq, 0, sizeof(unsigned int) * 4));
                }
            } else {
                l.bn.deferred = 0;   // the last block of the producer exchanges the sums and finalises (legacy path)
            }
        }
        set_bn_slot(bn_slot_);
    } else {
        cudaFree(comm_dev_); comm_dev_ = nullptr;
    }
    // kernel arguments are baked into the captured graphs
    for (auto& kv : graphs_) {
        if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
        if (kv.second.graph) cudaGraphDestroy(kv.second.graph);
    }
    graphs_.clear();
}

// =====================================================================================================
// forward: SingleStep.solve / MultiStep.solve
// =====================================================================================================
void Engine::forward(const float* y0, const float* t, int nt, int interpolate, float* out, n4j_stats* st) {
    N4J_CUDA(cudaSetDevice(device_));
    if (nt < 2) throw Error{N4J_ERR_INVALID_ARGUMENT, "time must be size 2! Was: " + std::to_string(nt)};
    if (nt - 1 > max_pairs_) throw Error{N4J_ERR_UNSUPPORTED, "too many time points"};
    std::vector<float> th(nt);
    if (is_device_ptr(t)) N4J_CUDA(cudaMemcpy(th.data(), t, sizeof(float) * nt, cudaMemcpyDeviceToHost));
    else std::memcpy(th.data(), t, sizeof(float) * nt);
    reset_ctrl();
    upload_params();
    params_fresh_ = true;
    ensure_zt((long long)nt * nz_);
    ensure(io_, (long long)nt * nz_);
    // Y[0] <- y0 (internal layout); zt_[0] keeps z(t0) for the multi-time output (MultiStep.java:49-60)
    state_in(y0, zt_.p);
    N4J_CUDA(cudaMemcpyAsync(Y_, zt_.p, sizeof(float) * nz_, cudaMemcpyDeviceToDevice, stream_));
    if (nz_pad_ != nz_) N4J_CUDA(cudaMemsetAsync(Y_ + nz_, 0, sizeof(float) * (nz_pad_ - nz_), stream_));
    const bool dev_out = is_device_ptr(out);
    if (nt == 2) {
        set_time_pairs({th[0], th[1]});
        run_solve(false, 0, false, nt);
        LAUNCH(k_copy_from_state, grid_for(nz_), 256, stream_, ctrl_, Y_, n_alloc_, 0LL, zt_.p, nz_);   // lastOutput (OdeGraphHelper.java:129)
        float* o = dev_out ? out : io_.p;
        state_out(zt_.p, o);
        if (!dev_out) from_device(out, io_.p, nz_);
    } else {
        if (interpolate) {
            // InterpolatingMultiStepSolver.integrate :28-47
            N4J_CUDA(cudaMemcpyAsync(wanted_, th.data() + 1, sizeof(float) * (nt - 1), cudaMemcpyHostToDevice, stream_));
            N4J_CUDA(cudaStreamSynchronize(stream_));   // th is pageable
            set_time_pairs({th[0], th[nt - 1]});
            run_solve(false, 0, true, nt);
        } else {
            // SingleSteppingMultiStepSolver.integrate :26-43
            std::vector<float> pairs;
            for (int k = 1; k < nt; ++k) { pairs.push_back(th[k - 1]); pairs.push_back(th[k]); }
            set_time_pairs(pairs);
            for (int k = 1; k < nt; ++k) {
                run_solve(false, 0, false, nt);
                LAUNCH(k_copy_from_state, grid_for(nz_), 256, stream_, ctrl_, Y_, n_alloc_, 0LL, zt_.p + (long long)k * nz_, nz_);
            }
        }
        float* o = dev_out ? out : io_.p;
        LAUNCH(k_timefirst_to_ref, grid_for((long long)nt * nz_), 256, stream_, zt_.p, o, nt, (int)B_, C_, H_ * W_, rank_ == 4 ? 1 : 0);
        if (!dev_out) from_device(out, io_.p, (long long)nt * nz_);
    }
    have_last_output_ = true;
    last_nt_ = nt;
    finish_call(st, 0);
}

// =====================================================================================================
// backward: SingleStepAdjoint.solve / MultiStepAdjoint.solve with the timegrad.* variants
// =====================================================================================================
__global__ void k_tg_step(Ctrl* c, double* dot_acc, float* tg, float* Ybase, long long stride, long long off_t, float* dLdt, int step, int phase) {
    pdl_begin();
    if (phase == 0 && c->comm) {     // multi-GPU: the dot product runs over all ranks' rows
        __shared__ double v[1];
        if (threadIdx.x == 0) v[0] = *dot_acc;
        __syncthreads();
        comm_allreduce_block(c->comm, v, 1);
        if (threadIdx.x == 0) *dot_acc = v[0];
    }
    if (threadIdx.x != 0) return;
    float* y = Ybase + (long long)c->y_cur * stride;
    if (phase == 0) {            // CalcTimeGrad.calcTimeAdjointT1 :52-58: lastTime -= <dL/dz(t1), f(z(t1))>; a_t := lastTime
        const float d = (float)(*dot_acc);
        *dot_acc = 0.0;
        tg[0] = __fsub_rn(tg[0], d);
        tg[1] = d;
        y[off_t] = tg[0];
        c->nfe += 1;
    } else if (phase == 1) {     // CalcTimeGrad.createLossGradient :66-74 + CalcMultiStepTimeGrad.updateStep :51-53
        dLdt[step - 1] = y[off_t];
        dLdt[step] = tg[1];
    } else if (phase == 2) {     // CalcMultiStepTimeGrad.updateLastStep :56-62
        dLdt[0] = tg[0];
    } else if (phase == 3) {     // start of a backward call
        tg[0] = 0.f; tg[1] = 0.f;
    }
}
// a-slot of the state: a <- g (first segment) or a <- g + a (MultiStepTimeGrad.prepareStep)
__global__ void __launch_bounds__(256) k_adjoint_seed(const Ctrl* __restrict__ c, float* __restrict__ Ybase, long long stride, long long off,
                                                       const float* __restrict__ g, const float* __restrict__ g2, int add_self, long long n) {
    pdl_begin();
    float* __restrict__ a = Ybase + (long long)c->y_cur * stride + off;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        float v = g[i];
        if (add_self) v = __fadd_rn(v, a[i]);
        if (g2) v = __fadd_rn(v, g2[i]);
        a[i] = v;
    }
}
__global__ void __launch_bounds__(256) k_add_out(const Ctrl* __restrict__ c, const float* __restrict__ Ybase, long long stride, long long off,
                                                  const float* __restrict__ add, float* __restrict__ dst, long long n) {
    pdl_begin();
    const float* __restrict__ a = Ybase + (long long)c->y_cur * stride + off;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) dst[i] = add ? __fadd_rn(a[i], add[i]) : a[i];
}

void Engine::backward(const float* last_output, const float* dLdz, const float* t, int nt, int tg_mode, float* grad_inout,
                      float* dLdy0, float* dLdt, n4j_stats* st) {
    N4J_CUDA(cudaSetDevice(device_));
    if (nt < 2) throw Error{N4J_ERR_INVALID_ARGUMENT, "time must be size 2! Was: " + std::to_string(nt)};
    if (nt - 1 > max_pairs_) throw Error{N4J_ERR_UNSUPPORTED, "too many time points"};
    std::vector<float> th(nt);
    if (is_device_ptr(t)) N4J_CUDA(cudaMemcpy(th.data(), t, sizeof(float) * nt, cudaMemcpyDeviceToHost));
    else std::memcpy(th.data(), t, sizeof(float) * nt);
    if (nt > 2) {   // MultiStepAdjoint.assertSorted :41-46
        int sgn = 0;
        for (int i = 0; i + 1 < nt; ++i) sgn += (th[i] > th[i + 1]) - (th[i] < th[i + 1]);
        if (std::abs(sgn) + 1 != nt) throw Error{N4J_ERR_INVALID_ARGUMENT, "Time must be in ascending or descending order!"};
    }
    if (!last_output && (!have_last_output_ || last_nt_ != nt))
        throw Error{N4J_ERR_ILLEGAL_STATE, "backward without last_output needs a preceding forward with the same time vector length"};
    if (tg_mode != N4J_TIMEGRAD_NONE && !dLdt) throw Error{N4J_ERR_INVALID_ARGUMENT, "dL_dt buffer required for this time-gradient mode"};
    const bool tg = tg_mode == N4J_TIMEGRAD_CALC;
    const int tslot = tg ? 1 : 0;
    const long long T = nt == 2 ? 1 : nt;          // stored time slices
    reset_ctrl();
    // The bound buffer is re-read at every call (node4j.h).  One exception: a backward that continues the forward of the same handle
    // (last_output == NULL, the OdeGraphHelper.doForward/doBackward pairing: the updater only runs after the backward pass) reuses
    // the weight images that forward built.  An explicit last_output (gradient checks, inference forwards in between) always re-uploads.
    if (!(params_fresh_ && !last_output)) upload_params();
    params_fresh_ = false;
    ensure_zt(std::max<long long>(T, 2) * nz_);
    ensure(gt_, T * nz_);
    ensure(io_, T * nz_);
    ensure(dldt_, nt);
    // --- inputs to the internal, time-first layout ----------------------------------------------------
    if (nt == 2) {
        if (last_output) state_in(last_output, zt_.p);
        state_in(dLdz, gt_.p);
    } else {
        if (last_output) {
            to_device(io_.p, last_output, T * nz_);
            LAUNCH(k_ref_to_timefirst, grid_for(T * nz_), 256, stream_, io_.p, zt_.p, nt, (int)B_, C_, H_ * W_, rank_ == 4 ? 1 : 0);
        }
        to_device(io_.p, dLdz, T * nz_);
        LAUNCH(k_ref_to_timefirst, grid_for(T * nz_), 256, stream_, io_.p, gt_.p, nt, (int)B_, C_, H_ * W_, rank_ == 4 ? 1 : 0);
    }
    // --- theta-adjoint seeded from the gradient view (SingleStepAdjoint.java:59-60) --------------------
    N4J_CUDA(cudaMemsetAsync(Y_, 0, sizeof(float) * 2 * n_alloc_, stream_));
    if (n_params_ > 0) {
        to_device(gradflat_.p, grad_inout, n_params_);
        theta_flat_to_internal(gradflat_.p, Y_ + off_theta_);     // y_cur == 0 after reset
    }
    LAUNCH(k_tg_step, 1, 32, stream_, ctrl_, dot_acc_, tg_, Y_, n_alloc_, off_t_, dldt_.p, 0, 3);
    if (dLdt) N4J_CUDA(cudaMemsetAsync(dldt_.p, 0, sizeof(float) * nt, stream_));

    // time pairs, reversed (Nd4j.reverse(time) :81), one per segment from the last to the first
    std::vector<float> pairs;
    for (int step = nt - 1; step > 0; --step) { pairs.push_back(th[step]); pairs.push_back(th[step - 1]); }
    set_time_pairs(pairs);

    bool have_prev = false;
    for (int step = nt - 1; step > 0; --step) {
        const long long slice = nt == 2 ? 0 : step;
        const float* z_k = zt_.p + slice * nz_;
        const float* g_k = gt_.p + slice * nz_;
        if (tg) {
            // d = <dL/dz(t_step) raw, f(z(t_step), t_step)>: one extra RHS evaluation
            tspec_ = TimeSpec{3, 0.f, th[step]};      // CalcTimeGrad.java:52: calculateDerivative(zt1, time.getColumn(1), ...)
            enqueue_rhs_fwd(stream_, fixed_ref(const_cast<float*>(z_k)), fixed_ref(gb_[2].p));
            bn_rearm_all(stream_);
            LAUNCH(k_dot, grid_for(nz_), 256, stream_, g_k, gb_[2].p, nz_, dot_acc_);
            LAUNCH(k_tg_step, 1, 32, stream_, ctrl_, dot_acc_, tg_, Y_, n_alloc_, off_t_, dldt_.p, step, 0);
        }
        // z-slot <- z(t_step); a-slot <- dL/dz(t_step) (+ running adjoint)
        LAUNCH(k_copy_to_state, grid_for(nz_), 256, stream_, ctrl_, Y_, n_alloc_, 0LL, z_k, (const float*)nullptr, nz_);
        LAUNCH(k_adjoint_seed, grid_for(nz_), 256, stream_, ctrl_, Y_, n_alloc_, off_a_, g_k, (const float*)nullptr, have_prev ? 1 : 0, nz_);
        run_solve(true, tslot, false, nt);
        if (tg) LAUNCH(k_tg_step, 1, 32, stream_, ctrl_, dot_acc_, tg_, Y_, n_alloc_, off_t_, dldt_.p, step, 1);
        have_prev = true;
    }
    if (tg && nt > 2) LAUNCH(k_tg_step, 1, 32, stream_, ctrl_, dot_acc_, tg_, Y_, n_alloc_, off_t_, dldt_.p, 0, 2);

    // --- outputs ---------------------------------------------------------------------------------------
    // dL/dy0 = a(t0) (+ dL/dz(t0) for multi-time, *MultiStepTimeGrad.updateLastStep)
    LAUNCH(k_add_out, grid_for(nz_), 256, stream_, ctrl_, Y_, n_alloc_, off_a_, nt > 2 ? gt_.p : (const float*)nullptr, gb_[0].p, nz_);
    const bool dev_dy0 = is_device_ptr(dLdy0);
    ensure(io2_, nz_);
    if (rank_ == 2) {
        from_device(dLdy0, gb_[0].p, nz_);
    } else {
        float* o = dev_dy0 ? dLdy0 : io_.p;
        state_out(gb_[0].p, o);
        if (!dev_dy0) from_device(dLdy0, io_.p, nz_);
    }
    if (n_params_ > 0) {
        LAUNCH(k_copy_from_state, grid_for(real_pad_), 256, stream_, ctrl_, Y_, n_alloc_, off_theta_, theta_tmp_.p, real_pad_);
        theta_internal_to_flat(theta_tmp_.p, gradflat_.p);     // realParamGrads.assignFrom :85 (mean/var slots untouched)
        from_device(grad_inout, gradflat_.p, n_params_);
    }
    if (dLdt) from_device(dLdt, dldt_.p, nt);
    finish_call(st, 0);
}

long long Engine::get_step_log(n4j_step_rec* buf, long long cap) {
    N4J_CUDA(cudaSetDevice(device_));
    const long long n = std::min<long long>(log_count_, log_cap_);
    const long long m = std::min(n, cap);
    if (buf && m > 0) N4J_CUDA(cudaMemcpy(buf, log_, sizeof(n4j_step_rec) * m, cudaMemcpyDeviceToHost));
    return log_count_;
}

// =====================================================================================================
// RHS on its own, error norm on caller data, solver-pass microbenchmark
// =====================================================================================================
void Engine::rhs(const float* z, float t, float* fz) {
    N4J_CUDA(cudaSetDevice(device_));
    t_pdl = pdl_;
    launches_ = 0;
    upload_params();
    ensure(io_, nz_);
    state_in(z, gb_[0].p);
    tspec_ = TimeSpec{3, 0.f, t};
    enqueue_rhs_fwd(stream_, fixed_ref(gb_[0].p), fixed_ref(gb_[1].p));
    bn_rearm_all(stream_);
    const bool dev = is_device_ptr(fz);
    float* o = dev ? fz : io_.p;
    state_out(gb_[1].p, o);
    if (!dev) from_device(fz, io_.p, nz_);
    N4J_CUDA(cudaStreamSynchronize(stream_));
}

void Engine::rhs_vjp(const float* z, float t, const float* g, float* fz, float* dz, float* dreal, float* dt) {
    N4J_CUDA(cudaSetDevice(device_));
    reset_ctrl();
    upload_params();
    ensure(io_, std::max<long long>(nz_, n_real_));
    // augmented state with a = -g so that the backpropagated epsilon (-a) equals g
    N4J_CUDA(cudaMemsetAsync(Y_, 0, sizeof(float) * n_alloc_, stream_));
    state_in(z, Y_);
    state_in(g, gb_[0].p);
    LAUNCH(k_act_bwd, grid_for(nz_), 256, stream_, fixed_ref(gb_[0].p), -1.f, fixed_ref(gb_[0].p), (int)N4J_ACT_IDENTITY, fixed_ref(Y_ + off_a_), nz_);
    tspec_ = TimeSpec{3, 0.f, t};
    aug_tslot_ = false;
    enqueue_rhs_aug(stream_, fixed_ref(Y_), fixed_ref(K_));
    bn_rearm_all(stream_);
    auto emit = [&](const float* internal, float* user) {
        const bool dev = is_device_ptr(user);
        float* o = dev ? user : io_.p;
        state_out(internal, o);
        if (!dev) { from_device(user, io_.p, nz_); N4J_CUDA(cudaStreamSynchronize(stream_)); }
    };
    emit(K_, fz);
    emit(K_ + off_a_, dz);
    if (n_real_ > 0) {
        theta_internal_to_real(K_ + off_theta_, theta_tmp_.p);
        from_device(dreal, theta_tmp_.p, n_real_);
    }
    if (dt) {
        if (time_input_) from_device(dt, tcat_misc_ + 1, 1); else *dt = 0.f;
    }
    N4J_CUDA(cudaStreamSynchronize(stream_));
}

void Engine::error_norm(const float* K, const float* y0, const float* y1, float step, long long n, float* err_out) {
    N4J_CUDA(cudaSetDevice(device_));
    DevBuf k, a, b;
    ensure(k, 7 * n); ensure(a, n); ensure(b, n);
    to_device(k.p, K, 7 * n); to_device(a.p, y0, n); to_device(b.p, y1, n);
    N4J_CUDA(cudaMemsetAsync(dot_acc_, 0, sizeof(double), stream_));
    k_error_only<<<grid_for(n), 256, 0, stream_>>>(k.p, a.p, b.p, n, step, (float)cfg_.abs_tol, (float)cfg_.rel_tol, dot_acc_);
    double sum = 0;
    N4J_CUDA(cudaMemcpyAsync(&sum, dot_acc_, sizeof(double), cudaMemcpyDeviceToHost, stream_));
    N4J_CUDA(cudaStreamSynchronize(stream_));
    N4J_CUDA(cudaMemsetAsync(dot_acc_, 0, sizeof(double), stream_));
    *err_out = (float)std::sqrt(sum / (double)n);
    cudaFree(k.p); cudaFree(a.p); cudaFree(b.p);
}

// 6 stage combinations + error norm over n floats, no RHS: the solver's own HBM traffic (160*n bytes per pass,
// SURVEY.md section 8d).  The controller is kept from advancing by forcing a rejection-free, non-terminating setup:
// state and K rows are zero, so err == 0 and h stays clamped at hmax.
void Engine::bench_solver_pass(long long n, int iters, float* us) {
    N4J_CUDA(cudaSetDevice(device_));
    t_pdl = pdl_;
    const long long n4 = round_up4(n) / 4, stride = n4 * 4;
    float *Y = nullptr, *K = nullptr;
    Ctrl* c = nullptr;
    N4J_CUDA(cudaMalloc(&Y, sizeof(float) * 2 * stride));
    N4J_CUDA(cudaMalloc(&K, sizeof(float) * 7 * stride));
    N4J_CUDA(cudaMalloc(&c, sizeof(Ctrl)));
    N4J_CUDA(cudaMemsetAsync(Y, 0, sizeof(float) * 2 * stride, stream_));
    N4J_CUDA(cudaMemsetAsync(K, 0, sizeof(float) * 7 * stride, stream_));
    N4J_CUDA(cudaMemsetAsync(c, 0, sizeof(Ctrl), stream_));
    Ctrl hc;
    std::memset(&hc, 0, sizeof(hc));
    for (int j = 0; j < 7; ++j) { hc.krow[j] = j; hc.krow_prev[j] = j; }
    hc.h = 0.01f; hc.t_end = 1e30f;
    hc.red_acc = red_acc_;
    N4J_CUDA(cudaMemcpyAsync(c, &hc, sizeof(Ctrl), cudaMemcpyHostToDevice, stream_));
    SolverParams sp = solver_params(n);
    sp.max_trials = 1LL << 60; sp.log_cap = 0;
    const int grid = grid_for(n4);
    BnFuse nobf{};
    auto pass = [&]() {
        launch_combine<1>(stream_, grid, c, Y, K, stride, 0, n4, nobf);
        launch_combine<2>(stream_, grid, c, Y, K, stride, 0, n4, nobf);
        launch_combine<3>(stream_, grid, c, Y, K, stride, 0, n4, nobf);
        launch_combine<4>(stream_, grid, c, Y, K, stride, 0, n4, nobf);
        launch_combine<5>(stream_, grid, c, Y, K, stride, 0, n4, nobf);
        launch_combine<6>(stream_, grid, c, Y, K, stride, 0, n4, nobf);
        k_error_ctrl<<<grid_red(n4), 256, 0, stream_>>>(c, Y, K, stride, n4, sp, nullptr, DenseOutArgs{nullptr, 0}, (cudaGraphConditionalHandle)0, 0);
    };
    for (int i = 0; i < 3; ++i) pass();
    N4J_CUDA(cudaEventRecord(ev0_, stream_));
    for (int i = 0; i < iters; ++i) pass();
    N4J_CUDA(cudaEventRecord(ev1_, stream_));
    N4J_CUDA(cudaStreamSynchronize(stream_));
    float ms = 0;
    N4J_CUDA(cudaEventElapsedTime(&ms, ev0_, ev1_));
    *us = ms * 1000.f / (float)iters;
    cudaFree(Y); cudaFree(K); cudaFree(c);
}

// Average device time of one launch sequence of a named piece of the hot path on the engine's real buffers.
void Engine::bench_piece(int what, int iters, float* us) {
    N4J_CUDA(cudaSetDevice(device_));
    t_pdl = pdl_;
    if (params_fresh_ == false && user_params_) { launches_ = 0; upload_params(); }
    Ref yin = fixed_ref(Y_), kout = fixed_ref(K_);
    ConvRec rec{};
    auto piece = [&]() {
        switch (what) {
            case N4J_BENCH_RHS_FWD: enqueue_rhs_fwd(stream_, yin, kout); break;
            case N4J_BENCH_RHS_AUG: enqueue_rhs_aug(stream_, yin, kout); break;
            case N4J_BENCH_CONV_FWD:
            case N4J_BENCH_CONV_DGRAD:
            case N4J_BENCH_CONV_WGRAD: {
                LayerPlan* lp = nullptr;
                for (auto& l : L_) if (l.d.kind == N4J_LAYER_CONV3X3) { lp = &l; break; }
                if (!lp) throw Error{N4J_ERR_INVALID_ARGUMENT, "inner graph has no conv3x3 layer"};
                LayerPlan& l = *lp;
                if (l.tc && what == N4J_BENCH_CONV_FWD) {
                    launch_conv_tc(stream_, l.img, l.wimg_f, fixed_ref(gb_[1].p), N4J_ACT_IDENTITY);
                } else if (l.tc && what == N4J_BENCH_CONV_DGRAD) {
                    launch_conv_tc(stream_, l.gimg, l.wimg_b, fixed_ref(gb_[1].p), N4J_ACT_IDENTITY);
                } else if (what == N4J_BENCH_CONV_FWD) {
                    GemmArgs g{}; g.a = fixed_ref(gb_[0].p); g.c = fixed_ref(gb_[1].p); g.w = l.wf; g.M = (int)l.rows; g.N = l.c_out; g.K = 9 * l.c_in;
                    g.act = 0; g.H = l.H; g.W = l.W; g.Ci = l.c_in; g.Co = l.c_out; g.splits = 1;
                    launch_gemm<G_CONV_FWD>(stream_, g, launches_, exact_);
                } else if (what == N4J_BENCH_CONV_DGRAD) {
                    GemmArgs q{}; q.a = fixed_ref(gb_[0].p); q.c = fixed_ref(gb_[1].p); q.w = l.wb; q.M = (int)l.rows; q.N = l.c_in; q.K = 9 * l.c_out;
                    q.H = l.H; q.W = l.W; q.Ci = l.c_in; q.Co = l.c_out; q.splits = 1;
                    launch_gemm<G_CONV_DGRAD>(stream_, q, launches_, exact_);
                } else if (l.tc) {
                    const int ipc = (int)((B_ + 63) / 64);
                    int wg_ctas = (int)((B_ + ipc - 1) / ipc);
                    if (wgrad_tc_) {
                        wg_ctas = wgeom_.tiles;
                        k_conv_wgrad_tc<<<dim3(wgeom_.tiles), WGRAD_TC_THREADS, wgrad_tc_smem_, stream_>>>(fixed_ref(gb_[0].p), fixed_ref(gb_[1].p), wgeom_, wpart_.p,
                                                                                                              (long long*)nullptr);
                    } else {
                        k_conv_wgrad64<<<wg_ctas, 256, wgrad_smem_, stream_>>>(fixed_ref(gb_[0].p), fixed_ref(gb_[1].p), geom_, wpart_.p, ipc, wgrad_nbuf_);
                    }
                    ++launches_;
                    LAUNCH(k_wgrad_reduce, 144, 256, stream_, wpart_.p, fixed_ref(theta_tmp_.p), wg_ctas, (CommDev*)nullptr);
                } else {
                    GemmArgs w{}; w.a = fixed_ref(gb_[0].p); w.b = fixed_ref(gb_[1].p); w.c = fixed_ref(wpart_.p); w.M = 9 * l.c_in; w.N = l.c_out; w.K = (int)l.rows;
                    w.H = l.H; w.W = l.W; w.Ci = l.c_in; w.Co = l.c_out; w.splits = 16;
                    launch_gemm<G_CONV_WGRAD>(stream_, w, launches_, exact_);
                    if (!exact_) LAUNCH(k_splitk_reduce, grid_for(l.g_len), 256, stream_, wpart_.p, fixed_ref(theta_tmp_.p), l.g_len, 16);
                }
                break;
            }
            case N4J_BENCH_CONV_FWD_FUSED:
            case N4J_BENCH_CONV_DGRAD_FUSED:
                launch_conv_args(stream_, rec.a, rec.pro, rec.epi);
                ++launches_;
                break;
            default: throw Error{N4J_ERR_INVALID_ARGUMENT, "unknown bench piece"};
        }
    };
    if (what == N4J_BENCH_CONV_FWD_FUSED || what == N4J_BENCH_CONV_DGRAD_FUSED) {
        // the conv launches exactly as one augmented evaluation issues them (operand prologue + statistics epilogue included):
        // record them from a real evaluation, then replay the first forward / dgrad one
        std::vector<ConvRec> recs;
        conv_rec_ = &recs;
        tspec_ = TimeSpec{3, 0.f, 0.f};
        aug_tslot_ = false;
        enqueue_rhs_aug(stream_, yin, kout);
        conv_rec_ = nullptr;
        const bool want_dgrad = what == N4J_BENCH_CONV_DGRAD_FUSED;
        bool found = false;
        for (auto& r : recs) {
            const bool is_dgrad = r.a.wimg != nullptr && r.pro != 1 && (r.pro == 2 || r.epi == 2);
            if (is_dgrad == want_dgrad && (want_dgrad || r.pro == 1 || r.epi == 1)) { rec = r; found = true; break; }
        }
        if (!found) throw Error{N4J_ERR_INVALID_ARGUMENT, "no fused tensor-core conv launch in this inner graph"};
    }
    if ((what == N4J_BENCH_CONV_FWD_FUSED || what == N4J_BENCH_CONV_DGRAD_FUSED) && getenv("N4J_CONV_DBG")) {
        // debug aid: SM-clock timestamps of CTA 0's phases (conv_stamp in conv_tc.cuh), printed relative to kernel entry
        long long* d = nullptr;
        N4J_CUDA(cudaMalloc(&d, sizeof(long long) * 16));
        ConvRec r2 = rec;
        r2.a.dbg = d;
        for (int rep = 0; rep < 3; ++rep) {
            N4J_CUDA(cudaMemsetAsync(d, 0, sizeof(long long) * 16, stream_));
            launch_conv_args(stream_, r2.a, r2.pro, r2.epi);
            long long h[16];
            N4J_CUDA(cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, stream_));
            N4J_CUDA(cudaStreamSynchronize(stream_));
            fprintf(stderr, "[conv dbg pro=%d epi=%d] cycles since entry: setup %lld | loads issued %lld | finals barrier %lld | t0 prologue done %lld | "
                            "last warp arrived %lld | mma start %lld | mma issued %lld | d_full seen %lld | out stored %lld | stats done %lld | exit %lld\n",
                    r2.pro, r2.epi, h[1] - h[0], h[2] - h[0], h[3] - h[0], h[4] - h[0], h[5] - h[0], h[6] - h[0], h[7] - h[0], h[8] - h[0], h[9] - h[0],
                    h[10] - h[0], h[11] - h[0]);
        }
        cudaFree(d);
    }
    for (int i = 0; i < 3; ++i) piece();
    N4J_CUDA(cudaEventRecord(ev0_, stream_));
    for (int i = 0; i < iters; ++i) piece();
    N4J_CUDA(cudaEventRecord(ev1_, stream_));
    N4J_CUDA(cudaStreamSynchronize(stream_));
    float ms = 0;
    N4J_CUDA(cudaEventElapsedTime(&ms, ev0_, ev1_));
    *us = ms * 1000.f / (float)iters;
    bn_rearm_all(stream_);
    N4J_CUDA(cudaStreamSynchronize(stream_));
}

}  // namespace n4j

// =====================================================================================================
// C ABI
// =====================================================================================================
struct n4j_handle {
    std::unique_ptr<n4j::Engine> eng;
    std::vector<n4j_layer_desc> layers;
};

static thread_local std::string g_last_error;

template <class F>
static int32_t guarded(F&& f) {
    try {
        f();
        return N4J_OK;
    } catch (const n4j::Error& e) {
        g_last_error = e.msg;
        return e.code;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return N4J_ERR_ILLEGAL_STATE;
    }
}

extern "C" {

int32_t node4j_abi_version(void) { return N4J_ABI_VERSION; }
const char* node4j_last_error(void) { return g_last_error.c_str(); }
int32_t node4j_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int32_t node4j_create(const n4j_graph_desc* graph, const n4j_solver_cfg* cfg, int32_t device, n4j_handle** out) {
    return guarded([&] {
        if (!graph || !cfg || !out) throw n4j::Error{N4J_ERR_INVALID_ARGUMENT, "null argument"};
        auto h = std::make_unique<n4j_handle>();
        h->layers.assign(graph->layers, graph->layers + graph->n_layers);
        n4j_graph_desc g = *graph;
        g.layers = h->layers.data();
        h->eng.reset(new n4j::Engine(g, *cfg, device));
        *out = h.release();
    });
}
int32_t node4j_destroy(n4j_handle* h) {
    return guarded([&] { delete h; });
}
int64_t node4j_num_params(const n4j_handle* h) { return h ? h->eng->num_params() : -1; }
int64_t node4j_num_real_params(const n4j_handle* h) { return h ? h->eng->num_real() : -1; }
int64_t node4j_state_len(const n4j_handle* h) { return h ? h->eng->state_len() : -1; }

int32_t node4j_bind_params(n4j_handle* h, const float* params, int64_t n_params) {
    return guarded([&] { h->eng->bind_params(params, n_params); });
}
int32_t node4j_forward(n4j_handle* h, const float* y0, const float* t, int32_t nt, int32_t interpolate, float* out, n4j_stats* stats) {
    return guarded([&] { h->eng->forward(y0, t, nt, interpolate, out, stats); });
}
int32_t node4j_backward(n4j_handle* h, const float* last_output, const float* dL_dz, const float* t, int32_t nt, int32_t time_grad_mode,
                        float* grad_inout, float* dL_dy0, float* dL_dt, n4j_stats* stats) {
    return guarded([&] { h->eng->backward(last_output, dL_dz, t, nt, time_grad_mode, grad_inout, dL_dy0, dL_dt, stats); });
}
int64_t node4j_get_step_log(n4j_handle* h, n4j_step_rec* buf, int64_t cap) {
    int64_t n = -1;
    guarded([&] { n = h->eng->get_step_log(buf, cap); });
    return n;
}
int32_t node4j_rhs(n4j_handle* h, const float* z, float t, float* fz) {
    return guarded([&] { h->eng->rhs(z, t, fz); });
}
int32_t node4j_rhs_vjp(n4j_handle* h, const float* z, float t, const float* g, float* fz, float* dz, float* dreal) {
    return guarded([&] { h->eng->rhs_vjp(z, t, g, fz, dz, dreal, nullptr); });
}
int32_t node4j_rhs_vjp_time(n4j_handle* h, const float* z, float t, const float* g, float* fz, float* dz, float* dreal, float* dt) {
    return guarded([&] { h->eng->rhs_vjp(z, t, g, fz, dz, dreal, dt); });
}
int32_t node4j_error_norm(n4j_handle* h, const float* K, const float* y0, const float* y1, float step, int64_t n, float* err_out) {
    return guarded([&] { h->eng->error_norm(K, y0, y1, step, n, err_out); });
}
int32_t node4j_comm_export(n4j_handle* h, void* handle64) {
    return guarded([&] { h->eng->comm_export(handle64); });
}
int32_t node4j_comm_init(n4j_handle* h, int32_t rank, int32_t world, const void* all_handles, int64_t global_batch) {
    return guarded([&] { h->eng->comm_init(rank, world, all_handles, global_batch); });
}
int32_t node4j_bench_solver_pass(n4j_handle* h, int64_t n, int32_t iters, float* us_per_pass) {
    return guarded([&] { h->eng->bench_solver_pass(n, iters, us_per_pass); });
}
int32_t node4j_bench_piece(n4j_handle* h, int32_t what, int32_t iters, float* us_per_iter) {
    return guarded([&] { h->eng->bench_piece(what, iters, us_per_iter); });
}
int32_t node4j_get_stream(n4j_handle* h, void** stream) {
    return guarded([&] { *stream = (void*)h->eng->stream(); });
}

// ---- the network around the ODE block (outer_net.cuh) ----
struct n4j_net {
    std::unique_ptr<n4j::OuterNet> net;
};
int32_t node4j_net_create(const n4j_net_desc* d, int32_t device, n4j_net** out) {
    return guarded([&] {
        if (!d || !out) throw n4j::Error{N4J_ERR_INVALID_ARGUMENT, "null argument"};
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) { cudaGetLastError(); throw n4j::Error{N4J_ERR_CUDA, "no usable CUDA device: the node4j engine has no CPU fallback"}; }
        std::unique_ptr<n4j_net> h(new n4j_net);
        h->net.reset(new n4j::OuterNet(d->batch, d->channels, d->solver, d->t0, d->t1, device));
        *out = h.release();
    });
}
int32_t node4j_net_destroy(n4j_net* net) {
    return guarded([&] { delete net; });
}
int64_t node4j_net_num_params(const n4j_net* net) { return net ? net->net->num_params() : -1; }
int32_t node4j_net_ode_slice(const n4j_net* net, int64_t* offset, int64_t* length) {
    return guarded([&] {
        if (!net || !offset || !length) throw n4j::Error{N4J_ERR_INVALID_ARGUMENT, "null argument"};
        *offset = net->net->ode_offset(); *length = net->net->ode_num_params();
    });
}
int32_t node4j_net_train_step(n4j_net* net, float* params, float* updater_state, const float* images, const float* labels, float lr, float momentum,
                              float* score, float* grad_out, n4j_stats* ode_forward, n4j_stats* ode_backward) {
    return guarded([&] {
        if (!net || !params || !updater_state || !images || !labels) throw n4j::Error{N4J_ERR_INVALID_ARGUMENT, "null argument"};
        net->net->train_step(params, updater_state, images, labels, lr, momentum, score, grad_out, ode_forward, ode_backward);
    });
}
int32_t node4j_net_output(n4j_net* net, const float* params, const float* images, float* probs, int32_t train) {
    return guarded([&] {
        if (!net || !params || !images || !probs) throw n4j::Error{N4J_ERR_INVALID_ARGUMENT, "null argument"};
        net->net->output(params, images, probs, train);
    });
}

}  // extern "C"
