// oracle/node_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the neuralODE4j ODE-block hot path (DormandPrince54 adaptive solve of an
// OdeVertex inner graph, forward + adjoint backward).  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load this library.  The product path
// (neuralode4j_b200/csrc) never links, imports or calls it.
//
// The reference is 100% Java on top of un-vendored ND4J/DL4J 1.0.0-beta3 and cannot be built or run
// in this image (no JVM); this file restates its algorithm from the sources under
// /root/reference/src/main/java (abbreviated J/ below), function by function, and is pinned against
// the reference's own golden vectors in tests/test_oracle_golden.py:
//   * MultiStepAdjointTest.java:188-231,252-295 (torchdiffeq vectors: ys, dL/dy0, dL/dtheta, dL/dt)
//   * InterpolationTest.java:37-44 (torchdiffeq quartic interpolation tensor)
//   * CircleODE / exponential closed forms (DormandPrince54SolverTest, SingleStepTest, MultiStepTest)
// Parity UNPINNED by the reference (its tests are smoke tests only): conv3x3 + BatchNorm(train) + ReLU
// numerics; for those layers this file *defines* the semantics.
//
// Arithmetic contract (shared with the CUDA path so that accept/reject sequences can be compared
// exactly): the instantiation R=float performs every *elementwise* operation in fp32 in the order
// written here, and every *reduction* (error norm, init-step sums, BatchNorm statistics, dense/conv
// contractions, dot products) accumulates exact fp32 products in double and rounds once.
// R=double does everything in double (used for the reference's double-precision tests).
//
// Build: see oracle/Makefile (g++ -O3 -march=x86-64-v3 -ffp-contract=off -fopenmp -shared -fPIC).

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ---------------------------------------------------------------------------------------------
// Public POD types (mirrored by oracle/oracle.py)
// ---------------------------------------------------------------------------------------------
// L_TIME_CONCAT: the reference's "timeConc" first vertex, ShapeMatchVertex(MergeVertex) fed by TimeInput
// (J/ode/vertex/impl/ShapeMatchVertex.java:50-79, J/util/preproc/DuplicateScalarToShape.java:41-60, J/ode/vertex/impl/helper/TimeInput.java:27-50):
// n_out channels (features) holding the current time are appended to the state along dimension 1.
enum LayerKind { L_DENSE = 1, L_CONV3X3 = 2, L_BATCHNORM = 3, L_ACTIVATION = 4, L_TIME_CONCAT = 5 };
enum Activation { A_IDENTITY = 0, A_RELU = 1, A_ELU = 2, A_TANH = 3 };
enum TimeGradMode { TG_NONE = 0, TG_ZERO = 1, TG_CALC = 2 };

struct OrcLayerDesc {
    int32_t kind;
    int32_t activation;
    int64_t n_in;   // features (dense) / channels (conv, batchnorm); 0 = infer
    int64_t n_out;
    int32_t has_bias;
    float eps;      // batchnorm epsilon (DL4J default 1e-5)
};

struct OrcGraphDesc {
    int32_t n_layers;
    int32_t rank;        // 2: [B,F]   4: [B,C,H,W]
    int64_t shape[4];
    const OrcLayerDesc* layers;
};

struct OrcSolverCfg {  // J/ode/solve/conf/SolverConfig.java:17-30 + StepConfig defaults :43-45
    double abs_tol, rel_tol, min_step, max_step;
    double safety, max_growth, min_reduction;
    int32_t order;
    int32_t exact_stage_times;  // 0 = reference quirk A.3 (every stage sees t+h)
};

struct OrcStats {
    int64_t nfe, accepted, rejected, solves;
};

struct OrcStepRec {
    double t, h, err;
    int32_t accepted;
    int32_t solve;
};

// ---------------------------------------------------------------------------------------------
// Tableau — J/ode/solve/impl/DormandPrince54Solver.java:27-48
// ---------------------------------------------------------------------------------------------
const double TAB_A[6][6] = {
    {1.0 / 5.0},
    {3.0 / 40.0, 9.0 / 40.0},
    {44.0 / 45.0, -56.0 / 15.0, 32.0 / 9.0},
    {19372.0 / 6561.0, -25360.0 / 2187.0, 64448.0 / 6561.0, -212.0 / 729.0},
    {9017.0 / 3168.0, -355.0 / 33.0, 46732.0 / 5247.0, 49.0 / 176.0, -5103.0 / 18656.0},
    {35.0 / 384.0, 0.0, 500.0 / 1113.0, 125.0 / 192.0, -2187.0 / 6784.0, 11.0 / 84.0}};
const double TAB_B[7] = {35.0 / 384.0, 0.0, 500.0 / 1113.0, 125.0 / 192.0, -2187.0 / 6784.0, 11.0 / 84.0, 0.0};
const double TAB_E[7] = {71.0 / 57600.0, 0.0, -71.0 / 16695.0, 71.0 / 1920.0, -17253.0 / 339200.0, 22.0 / 525.0, -1.0 / 40.0};
const double TAB_C[6] = {1.0 / 5.0, 3.0 / 10.0, 4.0 / 5.0, 8.0 / 9.0, 1.0, 1.0};
// NB: numerators without a 'd' suffix are int literals in the Java source but the first operand of
// each division is promoted by the double denominator, so all seven are plain double quotients.
const double TAB_CMID[7] = {6025192743.0 / 30085553152.0 / 2.0, 0.0, 51252292925.0 / 65400821598.0 / 2.0,
                            -2691868925.0 / 45128329728.0 / 2.0, 187940372067.0 / 1594534317056.0 / 2.0,
                            -1776094331.0 / 19743644256.0 / 2.0, 11237099.0 / 235043384.0 / 2.0};

template <class R> inline R rabs(R x) { return x < 0 ? -x : x; }

// Sensitivity probe (tests only): relative perturbation applied to every contraction result, emulating the
// different summation order of another implementation.  0 = off.
double g_probe_noise = 0.0;
inline double probe(double v, int64_t idx) {
    if (g_probe_noise == 0.0) return v;
    uint64_t x = (uint64_t)idx * 0x9E3779B97F4A7C15ull; x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
    const double u = (double)(x & 0xFFFFF) / (double)0xFFFFF * 2.0 - 1.0;
    return v * (1.0 + g_probe_noise * u);
}

// ---------------------------------------------------------------------------------------------
// Inner graph (the RHS f) — restates what ForwardPass.evaluate (J/ode/vertex/impl/helper/forward/
// ForwardPass.java:62-102) and BackpropagateAdjoint.backPropagate (J/.../backward/
// BackpropagateAdjoint.java:87-143) drive through DL4J for a *chain* of layers.  Layer math is
// third-party DL4J 1.0.0-beta3 (not in /root/reference); semantics per SURVEY.md A.9.
// ---------------------------------------------------------------------------------------------
struct LayerPlan {
    OrcLayerDesc d;
    int64_t in_len, out_len;      // per whole batch
    int64_t C, H, W, F;           // geometry of the layer INPUT (C,H,W for rank 4; F for rank 2)
    int64_t p_off, p_len;         // slot in the flat DL4J parameter vector
    int64_t g_off, g_len;         // slot in the contiguous real-gradient vector (g_len <= p_len)
};

struct NetPlan {
    std::vector<LayerPlan> L;
    int rank = 2;
    int64_t B = 0, n_state = 0, n_params = 0, n_real = 0;
    bool time_input = false;   // TimeInputFactory (J/ode/vertex/conf/OdeVertex.java:251)
    std::string err;
};

bool plan_net(const OrcGraphDesc& g, NetPlan& P) {
    P.rank = g.rank;
    if (g.rank != 2 && g.rank != 4) { P.err = "rank must be 2 or 4"; return false; }
    P.B = g.shape[0];
    int64_t C = 0, H = 0, W = 0, F = 0;
    if (g.rank == 2) F = g.shape[1]; else { C = g.shape[1]; H = g.shape[2]; W = g.shape[3]; }
    P.n_state = g.rank == 2 ? P.B * F : P.B * C * H * W;
    int64_t po = 0, go = 0;
    for (int i = 0; i < g.n_layers; ++i) {
        LayerPlan lp{};
        lp.d = g.layers[i];
        lp.C = C; lp.H = H; lp.W = W; lp.F = F;
        lp.in_len = g.rank == 2 ? P.B * F : P.B * C * H * W;
        switch (lp.d.kind) {
            case L_DENSE:
                if (g.rank != 2) { P.err = "dense layer needs rank-2 input"; return false; }
                if (lp.d.n_in == 0) lp.d.n_in = F;
                if (lp.d.n_in != F) { P.err = "dense n_in mismatch"; return false; }
                lp.p_len = lp.d.n_in * lp.d.n_out + (lp.d.has_bias ? lp.d.n_out : 0);
                lp.g_len = lp.p_len;
                F = lp.d.n_out;
                break;
            case L_CONV3X3:
                if (g.rank != 4) { P.err = "conv layer needs rank-4 input"; return false; }
                if (lp.d.n_in == 0) lp.d.n_in = C;
                if (lp.d.n_in != C) { P.err = "conv n_in mismatch"; return false; }
                if (lp.d.has_bias) { P.err = "conv3x3 with bias is out of scope (mnist/LayerUtil.java:68)"; return false; }
                if (H * W > 1024) { P.err = "conv3x3 feature map larger than 32x32 is out of scope"; return false; }
                lp.p_len = lp.d.n_out * lp.d.n_in * 9;
                lp.g_len = lp.p_len;
                C = lp.d.n_out;
                break;
            case L_BATCHNORM: {
                int64_t ch = g.rank == 2 ? F : C;
                if (lp.d.n_out == 0) lp.d.n_out = ch;
                if (lp.d.n_out != ch) { P.err = "batchnorm size mismatch"; return false; }
                lp.d.n_in = ch;
                lp.p_len = 4 * ch;   // gamma, beta, mean, var (BatchNormalizationParamInitializer order)
                lp.g_len = 2 * ch;   // GradientViewSelectionFromBlacklisted.java:33-37 drops mean/var
                break;
            }
            case L_ACTIVATION:
                lp.p_len = 0; lp.g_len = 0;
                break;
            case L_TIME_CONCAT:
                // OdeVertex.Builder(conf, name, vertex, addTimeAsInput=true, ...): only the FIRST vertex sees the time input
                if (i != 0) { P.err = "time concat is only supported as the first vertex of the inner graph"; return false; }
                if (lp.d.n_out <= 0) lp.d.n_out = 1;            // ShapeMatchVertex(MergeVertex): overrideSizeDims {1: 1}
                lp.d.n_in = g.rank == 2 ? F : C;
                lp.d.activation = A_IDENTITY;
                lp.p_len = 0; lp.g_len = 0;
                if (g.rank == 2) F += lp.d.n_out; else C += lp.d.n_out;
                P.time_input = true;
                break;
            default:
                P.err = "unknown layer kind"; return false;
        }
        lp.out_len = g.rank == 2 ? P.B * F : P.B * C * H * W;
        lp.p_off = po; lp.g_off = go;
        po += lp.p_len; go += lp.g_len;
        P.L.push_back(lp);
    }
    int64_t out = g.rank == 2 ? P.B * F : P.B * C * H * W;
    if (out != P.n_state) { P.err = "inner graph must map the state shape onto itself"; return false; }
    P.n_params = po; P.n_real = go;
    return true;
}

template <class R> inline R act_fwd(int a, R x) {
    switch (a) {
        case A_RELU: return x > 0 ? x : R(0);
        case A_ELU: return x > 0 ? x : R(std::exp(x) - R(1));
        case A_TANH: return R(std::tanh(x));
        default: return x;
    }
}
// derivative expressed through the layer OUTPUT y (what the CUDA path keeps, too)
template <class R> inline R act_bwd(int a, R y, R g) {
    switch (a) {
        case A_RELU: return y > 0 ? g : R(0);
        case A_ELU: return y > 0 ? g : R(g * R(y + R(1)));
        case A_TANH: { R yy = R(y * y); R d = R(R(1) - yy); return R(g * d); }
        default: return g;
    }
}

template <class R> struct Net {
    NetPlan P;
    const R* theta = nullptr;              // flat DL4J parameter vector
    std::vector<std::vector<R>> act;       // act[i] = output of layer i (post activation); input of i+1
    std::vector<std::vector<R>> xhat;      // batchnorm: normalised input
    std::vector<std::vector<R>> invstd;    // batchnorm: 1/sqrt(var+eps) per channel
    std::vector<R> x0;                     // copy of the graph input (needed by backward)
    std::vector<R> gbuf[2];
    int64_t nfe = 0;
    R dt_last = 0;                         // gradient w.r.t. the time input of the last backward() (TimeInput.update)

    void init(const NetPlan& p) {
        P = p;
        size_t n = P.L.size();
        act.resize(n); xhat.resize(n); invstd.resize(n);
        for (size_t i = 0; i < n; ++i) {
            act[i].assign(P.L[i].out_len, 0);
            if (P.L[i].d.kind == L_BATCHNORM) { xhat[i].assign(P.L[i].in_len, 0); invstd[i].assign(P.L[i].d.n_out, 0); }
        }
        x0.assign(P.n_state, 0);
    }

    // ---- forward of one layer -----------------------------------------------------------------
    void dense_fwd(const LayerPlan& l, const R* x, R* y) const {
        const int64_t B = P.B, nin = l.d.n_in, nout = l.d.n_out;
        const R* Wt = theta + l.p_off;            // W[nIn,nOut] F-order: W(i,o) = Wt[i + o*nIn]
        const R* bias = l.d.has_bias ? Wt + nin * nout : nullptr;
#pragma omp parallel for collapse(2) if (B * nout * nin > 65536)
        for (int64_t b = 0; b < B; ++b)
            for (int64_t o = 0; o < nout; ++o) {
                double acc = bias ? (double)bias[o] : 0.0;
                const R* xr = x + b * nin; const R* wc = Wt + o * nin;
                for (int64_t i = 0; i < nin; ++i) acc += (double)xr[i] * (double)wc[i];
                y[b * nout + o] = act_fwd<R>(l.d.activation, (R)acc);
            }
    }
    // The three convolution loops keep, for every output element, exactly the order of additions of the plain nested-loop form
    // ((ci, kh, kw) for the forward, (co, kh, kw) for dgrad, images then pixels for wgrad) — exact fp32 products accumulated in
    // double — but run the innermost loop over the channel dimension that is NOT reduced, on a transposed copy of the weights /
    // gradient, so that it vectorises.  Same bits, several times faster.
    void conv_fwd(const LayerPlan& l, const R* x, R* y) const {
        const int64_t B = P.B, Ci = l.d.n_in, Co = l.d.n_out, H = l.H, W = l.W;
        const R* Wt = theta + l.p_off;            // W[Cout,Cin,3,3] c-order, cross-correlation, pad 1
        std::vector<double> wT((size_t)Ci * 9 * Co);          // [ci][tap][co]
        for (int64_t co = 0; co < Co; ++co)
            for (int64_t ci = 0; ci < Ci; ++ci)
                for (int k = 0; k < 9; ++k) wT[((size_t)ci * 9 + k) * Co + co] = (double)Wt[(co * Ci + ci) * 9 + k];
#pragma omp parallel for collapse(2)
        for (int64_t b = 0; b < B; ++b)
            for (int64_t p = 0; p < H * W; ++p) {
                const int64_t h = p / W, ww = p - h * W;
                std::vector<double> acc((size_t)Co, 0.0);
                double* __restrict__ ac = acc.data();
                for (int64_t ci = 0; ci < Ci; ++ci) {
                    const R* xp = x + (b * Ci + ci) * H * W;
                    for (int kh = 0; kh < 3; ++kh) {
                        const int64_t hh = h + kh - 1;
                        if (hh < 0 || hh >= H) continue;
                        for (int kw = 0; kw < 3; ++kw) {
                            const int64_t wc = ww + kw - 1;
                            if (wc < 0 || wc >= W) continue;
                            const double xv = (double)xp[hh * W + wc];
                            const double* __restrict__ wr = wT.data() + ((size_t)ci * 9 + kh * 3 + kw) * Co;
                            for (int64_t co = 0; co < Co; ++co) ac[co] += wr[co] * xv;
                        }
                    }
                }
                for (int64_t co = 0; co < Co; ++co) {
                    const int64_t idx = (b * Co + co) * H * W + p;
                    y[idx] = act_fwd<R>(l.d.activation, (R)probe(ac[co], idx));
                }
            }
    }
    void bn_fwd(size_t li, const LayerPlan& l, const R* x, R* y) {
        const int64_t B = P.B, C = l.d.n_out, S = P.rank == 4 ? l.H * l.W : 1, M = B * S;
        const R* gamma = theta + l.p_off; const R* beta = gamma + C;
        R* xh = xhat[li].data(); R* is = invstd[li].data();
#pragma omp parallel for if (M * C > 65536)
        for (int64_t c = 0; c < C; ++c) {
            double s = 0, ss = 0;
            for (int64_t b = 0; b < B; ++b) {
                const R* xp = x + (b * C + c) * S;
                for (int64_t p = 0; p < S; ++p) { double v = (double)xp[p]; s += v; ss += v * v; }
            }
            // batch mean and *biased* variance over (N,H,W), training mode always (SingleStep.java:37)
            const double mean_d = s / (double)M;
            double var_d = ss / (double)M - mean_d * mean_d;
            if (var_d < 0) var_d = 0;
            const R mean = (R)mean_d;
            const R inv = (R)(1.0 / std::sqrt(var_d + (double)l.d.eps));
            is[c] = inv;
            const R g = gamma[c], bt = beta[c];
            for (int64_t b = 0; b < B; ++b) {
                const R* xp = x + (b * C + c) * S; R* hp = xh + (b * C + c) * S; R* yp = y + (b * C + c) * S;
                for (int64_t p = 0; p < S; ++p) {
                    const R d = R(xp[p] - mean);
                    const R h = R(d * inv);
                    hp[p] = h;
                    const R gh = R(g * h);
                    yp[p] = act_fwd<R>(l.d.activation, R(gh + bt));
                }
            }
        }
    }

    // f(z, t): ForwardPass.calculateDerivative, J/.../forward/ForwardPass.java:41-48 (nfe++ at :42)
    void forward(const R* z, R t, R* out) {
        ++nfe;
        std::memcpy(x0.data(), z, sizeof(R) * P.n_state);
        const R* cur = x0.data();
        for (size_t i = 0; i < P.L.size(); ++i) {
            const LayerPlan& l = P.L[i];
            R* y = act[i].data();
            switch (l.d.kind) {
                case L_DENSE: dense_fwd(l, cur, y); break;
                case L_CONV3X3: conv_fwd(l, cur, y); break;
                case L_BATCHNORM: bn_fwd(i, l, cur, y); break;
                case L_ACTIVATION:
                    for (int64_t k = 0; k < l.out_len; ++k) y[k] = act_fwd<R>(l.d.activation, cur[k]);
                    break;
                case L_TIME_CONCAT: {
                    // MergeVertex along dimension 1 of [state, t duplicated to the state's shape with n_out channels]
                    const int64_t Cs = l.d.n_in, k = l.d.n_out, S = P.rank == 4 ? l.H * l.W : 1;
                    for (int64_t b = 0; b < P.B; ++b) {
                        std::memcpy(y + b * (Cs + k) * S, cur + b * Cs * S, sizeof(R) * Cs * S);
                        for (int64_t q = 0; q < k * S; ++q) y[(b * (Cs + k) + Cs) * S + q] = t;
                    }
                    break;
                }
            }
            cur = y;
        }
        std::memcpy(out, cur, sizeof(R) * P.n_state);
    }

    // ---- backward (vjp) -----------------------------------------------------------------------
    // g_out: epsilon at the graph output; dx: epsilon wrt the graph input; dreal: contiguous real
    // gradient vector (overwritten — BackpropagateAdjoint.java:73 zeroes the flat gradients first).
    // Layer gradients are sums over the batch (no 1/B).
    void backward(const R* g_out, R* dx, R* dreal) {
        gbuf[0].assign(g_out, g_out + P.n_state);
        dt_last = 0;
        int cur = 0;
        for (int64_t i = (int64_t)P.L.size() - 1; i >= 0; --i) {
            const LayerPlan& l = P.L[i];
            const R* xin = i == 0 ? x0.data() : act[i - 1].data();
            const R* y = act[i].data();
            std::vector<R>& gin = gbuf[cur];
            std::vector<R>& gout = gbuf[cur ^ 1];
            gout.assign(l.in_len, 0);
            // delta = g * act'(y)
            for (int64_t k = 0; k < l.out_len; ++k) gin[k] = act_bwd<R>(l.d.activation, y[k], gin[k]);
            switch (l.d.kind) {
                case L_DENSE: {
                    const int64_t B = P.B, nin = l.d.n_in, nout = l.d.n_out;
                    const R* Wt = theta + l.p_off;
                    R* dW = dreal + l.g_off; R* db = dW + nin * nout;
#pragma omp parallel for collapse(2) if (B * nout * nin > 65536)
                    for (int64_t o = 0; o < nout; ++o)
                        for (int64_t ii = 0; ii < nin; ++ii) {
                            double acc = 0;
                            for (int64_t b = 0; b < B; ++b) acc += (double)xin[b * nin + ii] * (double)gin[b * nout + o];
                            dW[ii + o * nin] = (R)acc;
                        }
                    if (l.d.has_bias)
                        for (int64_t o = 0; o < nout; ++o) {
                            double acc = 0;
                            for (int64_t b = 0; b < B; ++b) acc += (double)gin[b * nout + o];
                            db[o] = (R)acc;
                        }
#pragma omp parallel for collapse(2) if (B * nout * nin > 65536)
                    for (int64_t b = 0; b < B; ++b)
                        for (int64_t ii = 0; ii < nin; ++ii) {
                            double acc = 0;
                            for (int64_t o = 0; o < nout; ++o) acc += (double)gin[b * nout + o] * (double)Wt[ii + o * nin];
                            gout[b * nin + ii] = (R)acc;
                        }
                    break;
                }
                case L_CONV3X3: {
                    const int64_t B = P.B, Ci = l.d.n_in, Co = l.d.n_out, H = l.H, W = l.W;
                    const R* Wt = theta + l.p_off;
                    R* dW = dreal + l.g_off;
                    // wgrad: dW[co,ci,kh,kw] = sum_b ( sum_{h,w} x[b,ci,h+kh-1,w+kw-1] * delta[b,co,h,w] )   (per-image sums, then images)
                    std::vector<double> gT((size_t)B * H * W * Co);      // delta as [b][pixel][co]
#pragma omp parallel for
                    for (int64_t b = 0; b < B; ++b)
                        for (int64_t co = 0; co < Co; ++co)
                            for (int64_t q = 0; q < H * W; ++q) gT[((size_t)b * H * W + q) * Co + co] = (double)gin[(b * Co + co) * H * W + q];
#pragma omp parallel for collapse(2)
                    for (int64_t ci = 0; ci < Ci; ++ci)
                        for (int k = 0; k < 9; ++k) {
                            const int kh = k / 3, kw = k % 3;
                            const int64_t h0 = std::max<int64_t>(0, 1 - kh), h1 = std::min<int64_t>(H, H + 1 - kh);
                            const int64_t w0 = std::max<int64_t>(0, 1 - kw), w1 = std::min<int64_t>(W, W + 1 - kw);
                            std::vector<double> accv((size_t)Co, 0.0), av((size_t)Co);
                            double* __restrict__ acc = accv.data();
                            double* __restrict__ a = av.data();
                            for (int64_t b = 0; b < B; ++b) {
                                const R* xp = xin + (b * Ci + ci) * H * W;
                                for (int64_t co = 0; co < Co; ++co) a[co] = 0.0;
                                for (int64_t h = h0; h < h1; ++h)
                                    for (int64_t ww = w0; ww < w1; ++ww) {
                                        const double xv = (double)xp[(h + kh - 1) * W + (ww + kw - 1)];
                                        const double* __restrict__ gr = gT.data() + ((size_t)b * H * W + h * W + ww) * Co;
                                        for (int64_t co = 0; co < Co; ++co) a[co] += xv * gr[co];
                                    }
                                for (int64_t co = 0; co < Co; ++co) acc[co] += a[co];
                            }
                            for (int64_t co = 0; co < Co; ++co) dW[(co * Ci + ci) * 9 + k] = (R)probe(acc[co], (co * Ci + ci) * 9 + k);
                        }
                    // dgrad: dx[b,ci,y,x] = sum_{co,kh,kw} delta[b,co,y-kh+1,x-kw+1] * W[co,ci,kh,kw]
                    std::vector<double> wT2((size_t)Co * 9 * Ci);        // [co][tap][ci]
                    for (int64_t co = 0; co < Co; ++co)
                        for (int64_t ci = 0; ci < Ci; ++ci)
                            for (int k = 0; k < 9; ++k) wT2[((size_t)co * 9 + k) * Ci + ci] = (double)Wt[(co * Ci + ci) * 9 + k];
#pragma omp parallel for collapse(2)
                    for (int64_t b = 0; b < B; ++b)
                        for (int64_t q = 0; q < H * W; ++q) {
                            const int64_t qh = q / W, qw = q - qh * W;
                            std::vector<double> accv((size_t)Ci, 0.0);
                            double* __restrict__ acc = accv.data();
                            for (int64_t co = 0; co < Co; ++co) {
                                const R* gp = gin.data() + (b * Co + co) * H * W;
                                for (int kh = 0; kh < 3; ++kh) {
                                    const int64_t h = qh - kh + 1;      // the output pixel (h,w) whose tap (kh,kw) reads input pixel q
                                    if (h < 0 || h >= H) continue;
                                    for (int kw = 0; kw < 3; ++kw) {
                                        const int64_t ww = qw - kw + 1;
                                        if (ww < 0 || ww >= W) continue;
                                        const double gv = (double)gp[h * W + ww];
                                        const double* __restrict__ wr = wT2.data() + ((size_t)co * 9 + kh * 3 + kw) * Ci;
                                        for (int64_t ci = 0; ci < Ci; ++ci) acc[ci] += wr[ci] * gv;
                                    }
                                }
                            }
                            for (int64_t ci = 0; ci < Ci; ++ci) {
                                const int64_t idx = (b * Ci + ci) * H * W + q;
                                gout[idx] = (R)probe(acc[ci], idx);
                            }
                        }
                    break;
                }
                case L_BATCHNORM: {
                    const int64_t B = P.B, C = l.d.n_out, S = P.rank == 4 ? l.H * l.W : 1, M = B * S;
                    const R* gamma = theta + l.p_off;
                    const R* xh = xhat[i].data(); const R* is = invstd[i].data();
                    R* dgamma = dreal + l.g_off; R* dbeta = dgamma + C;
#pragma omp parallel for if (M * C > 65536)
                    for (int64_t c = 0; c < C; ++c) {
                        double sdy = 0, sdyxh = 0;
                        for (int64_t b = 0; b < B; ++b) {
                            const R* gp = gin.data() + (b * C + c) * S; const R* hp = xh + (b * C + c) * S;
                            for (int64_t p = 0; p < S; ++p) { sdy += (double)gp[p]; sdyxh += (double)gp[p] * (double)hp[p]; }
                        }
                        dgamma[c] = (R)sdyxh; dbeta[c] = (R)sdy;
                        const R m_dy = (R)(sdy / (double)M), m_dyxh = (R)(sdyxh / (double)M);
                        const R k = R(gamma[c] * is[c]);
                        for (int64_t b = 0; b < B; ++b) {
                            const R* gp = gin.data() + (b * C + c) * S; const R* hp = xh + (b * C + c) * S;
                            R* op = gout.data() + (b * C + c) * S;
                            for (int64_t p = 0; p < S; ++p) {
                                const R t1 = R(gp[p] - m_dy);
                                const R t2 = R(hp[p] * m_dyxh);
                                op[p] = R(k * R(t1 - t2));
                            }
                        }
                    }
                    break;
                }
                case L_ACTIVATION:
                    std::copy(gin.begin(), gin.begin() + l.in_len, gout.begin());
                    break;
                case L_TIME_CONCAT: {
                    // MergeVertex.doBackward splits the epsilon; DuplicateScalarToShape.backprop sums the time part (:60-62)
                    const int64_t Cs = l.d.n_in, k = l.d.n_out, S = P.rank == 4 ? l.H * l.W : 1;
                    double acc = 0;
                    for (int64_t b = 0; b < P.B; ++b) {
                        std::memcpy(gout.data() + b * Cs * S, gin.data() + b * (Cs + k) * S, sizeof(R) * Cs * S);
                        for (int64_t q = 0; q < k * S; ++q) acc += (double)gin[(b * (Cs + k) + Cs) * S + q];
                    }
                    dt_last = (R)acc;
                    break;
                }
            }
            cur ^= 1;
        }
        std::memcpy(dx, gbuf[cur].data(), sizeof(R) * P.n_state);
    }
};

// ---------------------------------------------------------------------------------------------
// Solver — AdaptiveRungeKuttaSolver.java:107-181, FirstOrderEquationWithState.java:29-158,
// DormandPrince54Solver.java:78-93, AdaptiveRungeKuttaStepPolicy.java:90-153
// ---------------------------------------------------------------------------------------------
template <class R> struct Equation {   // FirstOrderEquation.calculateDerivative(y, t, fy)
    virtual void eval(const R* y, R t, R* fy) = 0;
    virtual ~Equation() {}
};

template <class R> struct StepListenerI {   // J/ode/solve/api/StepListener.java:18-31
    virtual void begin(R t0, R t1, const R* y0) = 0;
    // called after update() and before shiftDerivative() (AdaptiveRungeKuttaSolver.java:172-176)
    virtual void step(const R* const* K, const R* y, R t, R h, R err) = 0;
    virtual void done() = 0;
    virtual ~StepListenerI() {}
};

// DormandPrince54Mse.estimateMse, DormandPrince54Solver.java:86-91
template <class R> R error_norm(const R* const* K, const R* y0, const R* y1, R h, int64_t N, R atol, R rtol) {
    R e[7];
    for (int j = 0; j < 7; ++j) e[j] = (R)TAB_E[j];
    double sum = 0;
#pragma omp parallel for reduction(+ : sum) if (N > 65536)
    for (int64_t i = 0; i < N; ++i) {
        // errSum = e . K  (row 1 has a zero coefficient and is skipped)
        R es = R(e[0] * K[0][i]);
        es = std::fma(e[2], K[2][i], es);
        es = std::fma(e[3], K[3][i], es);
        es = std::fma(e[4], K[4][i], es);
        es = std::fma(e[5], K[5][i], es);
        es = std::fma(e[6], K[6][i], es);
        const R ys = std::max(rabs(y0[i]), rabs(y1[i]));
        const R tol = R(R(ys * rtol) + atol);
        const R ratio = R(R(es / tol) * h);
        sum += (double)ratio * (double)ratio;
    }
    return (R)std::sqrt(sum / (double)N);
}

// AdaptiveRungeKuttaStepPolicy.step, :142-153
template <class R> R next_step(R h, R err, const OrcSolverCfg& c) {
    const R exponent_base = (R)std::pow((double)err, -1.0 / (double)c.order);
    R fac = R(exponent_base * (R)c.safety);
    fac = std::min((R)c.max_growth, std::max((R)c.min_reduction, fac));
    const R sign = h > 0 ? R(1) : (h < 0 ? R(-1) : R(0));
    R v = R(R(fac * h) * sign);
    v = std::min((R)c.max_step, std::max((R)c.min_step, v));
    return R(v * sign);
}

template <class R> struct Solver {
    OrcSolverCfg cfg;
    OrcStats* stats = nullptr;
    std::vector<OrcStepRec>* log = nullptr;
    std::vector<StepListenerI<R>*> listeners;
    std::vector<R> yW, Kbuf;
    int64_t max_trials = 0;  // 0 = unbounded like the reference

    // AdaptiveRungeKuttaStepPolicy.initializeStep, :90-139.  K0 valid on entry; recomputes K0
    // (the reference's redundant eval :92) and leaves the probe derivative in K1.
    R initialize_step(Equation<R>& eq, R t0, bool backward, const R* y, R* K0, R* K1, int64_t N) {
        eq.eval(y, t0, K0);
        const R atol = (R)cfg.abs_tol, rtol = (R)cfg.rel_tol;
        double Y = 0, D = 0;
#pragma omp parallel for reduction(+ : Y, D) if (N > 65536)
        for (int64_t i = 0; i < N; ++i) {
            const R scal = R(R(rabs(y[i]) * rtol) + atol);
            const R r1 = R(y[i] / scal), r2 = R(K0[i] / scal);
            Y += (double)r1 * (double)r1;
            D += (double)r2 * (double)r2;
        }
        const R Yr = (R)Y, Dr = (R)D;
        R h = ((double)Yr < 1.0e-10 || (double)Dr < 1.0e-10) ? (R)1e-6 : R(std::sqrt(R(Yr / Dr)) * (R)0.01);
        if (backward) h = -h;
        // equation.step(ones(1), step); calculateDerivative(1): probe at y + h*K0, time t0 + h
        for (int64_t i = 0; i < N; ++i) yW[i] = R(y[i] + R(h * K0[i]));
        eq.eval(yW.data(), R(t0 + h), K1);
        double DDs = 0;
#pragma omp parallel for reduction(+ : DDs) if (N > 65536)
        for (int64_t i = 0; i < N; ++i) {
            const R scal = R(R(rabs(y[i]) * rtol) + atol);
            const R r = R(R(K1[i] - K0[i]) / scal);
            DDs += (double)r * (double)r;
        }
        const R DD = R(std::sqrt((R)DDs) / h);   // signed: negative when integrating backwards (:118,122)
        const R m = std::max(std::sqrt(Dr), DD);
        R h1;
        if ((double)m < 1e-15) h1 = std::max((R)1e-6, R(rabs(h) * (R)0.001));
        else h1 = (R)std::pow((double)R((R)0.01 / m), 1.0 / (double)cfg.order);
        h = std::min(R(rabs(h) * (R)100), h1);
        h = std::max(h, R(rabs(t0) * (R)1e-12));
        h = std::max((R)cfg.min_step, h);
        h = std::min((R)cfg.max_step, h);
        if (backward) h = -h;
        return h;
    }

    // AdaptiveRungeKuttaSolver.integrate + solve; y is the caller's yOut already holding y0 (:117)
    void integrate(Equation<R>& eq, R t0, R t1, R* y, int64_t N, int solve_idx) {
        yW.assign(N, 0);
        Kbuf.assign(7 * N, 0);
        R* K[7];
        for (int j = 0; j < 7; ++j) K[j] = Kbuf.data() + j * N;
        for (auto* l : listeners) l->begin(t0, t1, y);

        R t = t0;
        eq.eval(y, t, K[0]);                                            // :131
        if (stats) stats->nfe += 1;
        const bool backward = !(t1 > t0);                               // t.argMax()==0 (:139, policy :106)
        R h = initialize_step(eq, t0, backward, y, K[0], K[1], N);      // :136
        if (stats) { stats->nfe += 2; stats->solves += 1; }
        const double tLast = (double)t1;
        R a[6][6], b[7];
        for (int s = 0; s < 6; ++s) for (int j = 0; j <= s; ++j) a[s][j] = (R)TAB_A[s][j];
        for (int j = 0; j < 7; ++j) b[j] = (R)TAB_B[j];

        bool last;
        int64_t trials = 0;
        do {
            // TimeLimitForwards/Backwards.isLastStep, :76-103
            last = false;
            const R tph = R(t + h);
            if (backward ? ((double)tph <= tLast) : ((double)tph >= tLast)) {
                h = (R)(tLast - (double)t);
                last = true;
            }
            for (int s = 1; s <= 6; ++s) {
                // State.step(a[s-1], h): yW = y + (a*h) . K[0:s]   (FirstOrderEquationWithState.java:42-51)
                R c[6];
                for (int j = 0; j < s; ++j) c[j] = R(a[s - 1][j] * h);
                const bool skip1 = (s == 6);   // a[5][1] == 0
#pragma omp parallel for if (N > 65536)
                for (int64_t i = 0; i < N; ++i) {
                    R acc = R(c[0] * K[0][i]);
                    for (int j = 1; j < s; ++j) {
                        if (skip1 && j == 1) continue;
                        acc = std::fma(c[j], K[j][i], acc);
                    }
                    yW[i] = R(y[i] + acc);
                }
                // quirk A.3: timeOffset holds the FULL step, every stage is evaluated at t + h (:45,:82)
                const R ts = cfg.exact_stage_times ? R(t + R((R)TAB_C[s - 1] * h)) : R(t + h);
                eq.eval(yW.data(), ts, K[s]);
            }
            if (stats) stats->nfe += 6;
            // State.step(b, h) (:156): b[0:6] == a[5][0:6] and b[6] == 0, so yW is already y1; the
            // reference recomputes the identical combination.
            (void)b;
            const R err = error_norm<R>(K, y, yW.data(), h, N, (R)cfg.abs_tol, (R)cfg.rel_tol);   // :159
            const bool ok = (double)err < 1.0;                                                    // :170
            if (ok) {
                t = R(t + h);                                              // update(): time.addi(timeOffset)
                std::memcpy(y, yW.data(), sizeof(R) * N);                  // y.assign(yWorking)
                for (auto* l : listeners) l->step(K, y, t, h, err);        // :174
                std::memcpy(K[0], K[6], sizeof(R) * N);                    // shiftDerivative() FSAL :176
                if (stats) stats->accepted += 1;
            } else if (stats) stats->rejected += 1;
            if (log) log->push_back(OrcStepRec{(double)t, (double)h, (double)err, ok ? 1 : 0, solve_idx});
            last = last && ok;                                              // :161
            h = next_step<R>(h, err, cfg);                                  // :164
            ++trials;
            if (max_trials && trials >= max_trials) break;
        } while (!last);
        for (auto* l : listeners) l->done();
    }
};

// ---------------------------------------------------------------------------------------------
// Dense output — Interpolation.java:33-84 and InterpolatingStepListener.java:59-150
// ---------------------------------------------------------------------------------------------
template <class R> struct Interp {
    std::vector<R> c[5];
    void fit(const R* y0, const R* y1, const R* ymid, const R* f0, const R* f1, R dt, int64_t N) {
        for (auto& v : c) v.resize(N);
        const double dtd = (double)dt;
        const R F0[5] = {(R)(-2 * dtd), (R)(2 * dtd), (R)-8, (R)-8, (R)16};
        const R F1[5] = {(R)(5 * dtd), (R)(-3 * dtd), (R)18, (R)14, (R)-32};
        const R F2[5] = {(R)(-4 * dtd), (R)dtd, (R)-11, (R)-5, (R)16};
        for (int64_t i = 0; i < N; ++i) {
            const R in[5] = {f0[i], f1[i], y0[i], y1[i], ymid[i]};
            const R* Fs[3] = {F0, F1, F2};
            for (int k = 0; k < 3; ++k) {
                R o = R(in[0] * Fs[k][0]);                       // dotProduct :51-57: assign, then addi
                for (int j = 1; j < 5; ++j) o = R(o + R(in[j] * Fs[k][j]));
                c[k][i] = o;
            }
            c[3][i] = R(f0[i] * dt);
            c[4][i] = y0[i];
        }
    }
    void eval(double t0, double t1, double t, R* out, int64_t N) const {
        const double x = (t - t0) / (t1 - t0);
        double xs[5]; xs[4] = 1;
        for (int i = 3; i >= 0; --i) xs[i] = xs[i + 1] * x;
        R xr[5]; for (int i = 0; i < 5; ++i) xr[i] = (R)xs[i];
        for (int64_t i = 0; i < N; ++i) {
            R o = R(c[0][i] * xr[0]);
            for (int j = 1; j < 5; ++j) o = R(o + R(c[j][i] * xr[j]));
            out[i] = o;
        }
    }
};

template <class R> struct InterpListener : StepListenerI<R> {
    const R* wanted; int64_t nw; R* yout; int64_t N;
    R t0s = 0; const R* y0p = nullptr; std::vector<R> y0copy, ymid;
    Interp<R> ip;
    InterpListener(const R* w, int64_t n, R* out, int64_t len) : wanted(w), nw(n), yout(out), N(len) {}
    void begin(R t0, R, const R* y0) override {
        t0s = t0; y0p = y0;     // NOT a copy: aliases the solver state (InterpolatingMultiStepSolver.java:41)
        if (std::fabs((double)t0 - (double)wanted[0]) < 1e-10) std::memcpy(yout, y0, sizeof(R) * N);
    }
    void step(const R* const* K, const R* y, R t, R h, R) override {
        const double lo = h > 0 ? (double)t0s : (double)t, hi = h > 0 ? (double)t : (double)t0s;
        int64_t first = -1, cnt = 0;
        for (int64_t i = 0; i < nw; ++i)
            if ((double)wanted[i] > lo && (double)wanted[i] < hi) { if (first < 0) first = i; ++cnt; }
        if (cnt > 0) {
            ymid.resize(N);
            R cm[7]; for (int j = 0; j < 7; ++j) cm[j] = (R)TAB_CMID[j];
            for (int64_t i = 0; i < N; ++i) {
                R o = R(R(K[0][i] * cm[0]) * h);                            // scaledDotProduct :122-128
                for (int j = 1; j < 7; ++j) o = R(o + R(R(K[j][i] * cm[j]) * h));
                ymid[i] = R(y0p[i] + o);
            }
            ip.fit(y0p, y, ymid.data(), K[0], K[6], h, N);
            for (int64_t i = first; i < first + cnt; ++i) ip.eval((double)t0s, (double)t, (double)wanted[i], yout + i * N, N);
        }
        t0s = t;
        y0copy.assign(y, y + N);
        y0p = y0copy.data();
    }
    void done() override {
        if (std::fabs((double)t0s - (double)wanted[nw - 1]) < 1e-10) std::memcpy(yout + (nw - 1) * N, y0p, sizeof(R) * N);
    }
};

// ---------------------------------------------------------------------------------------------
// Handle: forward/backward helpers
// ---------------------------------------------------------------------------------------------
template <class R> struct FwdEq : Equation<R> {
    Net<R>* net;
    explicit FwdEq(Net<R>* n) : net(n) {}
    void eval(const R* y, R t, R* fy) override { net->forward(y, t, fy); }
};

// BackpropagateAdjoint.calculateDerivative, J/.../backward/BackpropagateAdjoint.java:63-85 with the
// layout of AugmentedDynamics.java:23-30 and SingleStepAdjoint.java:58-66:
//   [ z | a | theta_real | a_t (0 or 1) | pad up to numParams ]
template <class R> struct AdjEq : Equation<R> {
    Net<R>* net; int64_t nz, nreal, nt, ntot;
    std::vector<R> nega, fz, da, dth;
    AdjEq(Net<R>* n, int64_t nt_) : net(n), nz(n->P.n_state), nreal(n->P.n_real), nt(nt_) {
        ntot = 2 * nz + n->P.n_params + nt;
        nega.resize(nz); fz.resize(nz); da.resize(nz); dth.resize(std::max<int64_t>(nreal, 1));
    }
    void eval(const R* zaug, R t, R* out) override {
        std::memcpy(out, zaug, sizeof(R) * ntot);              // updateFrom + transferTo: pad = state
        net->forward(zaug, t, fz.data());                      // :69  z-slot <- f(z,t)
        for (int64_t i = 0; i < nz; ++i) nega[i] = -zaug[nz + i];   // :75  zAdjoint().negi()
        net->backward(nega.data(), da.data(), dth.data());     // :75/:78
        std::memcpy(out, fz.data(), sizeof(R) * nz);
        std::memcpy(out + nz, da.data(), sizeof(R) * nz);      // :79 updateZAdjoint
        std::memcpy(out + 2 * nz, dth.data(), sizeof(R) * nreal);
        // NoTimeInput.update (NoTimeInput.java:26-29): tAdjoint := 0;  TimeInput.update (TimeInput.java:46-50): tAdjoint := the
        // epsilon of the time input (an assign into the empty array of NoTimeGrad when nt == 0, i.e. dropped)
        for (int64_t i = 0; i < nt; ++i) out[2 * nz + nreal + i] = net->P.time_input ? net->dt_last : R(0);
    }
};

template <class R> struct Handle {
    NetPlan plan;
    Net<R> net;
    OrcSolverCfg cfg;
    std::vector<OrcStepRec> log;
    std::string err;

    void gather_real(const R* flat, R* real) const {
        for (auto& l : plan.L) std::memcpy(real + l.g_off, flat + l.p_off, sizeof(R) * l.g_len);
    }
    void scatter_real(const R* real, R* flat) const {
        for (auto& l : plan.L) std::memcpy(flat + l.p_off, real + l.g_off, sizeof(R) * l.g_len);
    }

    // FixedStep/InputStep forward: SingleStep.java:32-46 (nt==2) / MultiStep.java:27-61 (nt>2)
    // out: nt==2 -> [state];  nt>2 -> time-first [nt, state] incl. z(t0) (alignOutShape's permute is
    // applied by the caller-facing wrapper in oracle.py)
    int forward(const R* params, const R* y0, const R* t, int nt, int interpolate, R* out, OrcStats* st) {
        if (nt < 2) { err = "time must have at least two elements"; return 1; }
        net.theta = params; net.nfe = 0; log.clear();
        const int64_t N = plan.n_state;
        Solver<R> s; s.cfg = cfg; s.stats = st; s.log = &log;
        FwdEq<R> eq(&net);
        if (nt == 2) {
            std::memcpy(out, y0, sizeof(R) * N);
            s.integrate(eq, t[0], t[1], out, N, 0);
            return 0;
        }
        std::memcpy(out, y0, sizeof(R) * N);     // z(t0) prepended (MultiStep.java:49-60)
        if (interpolate) {
            // InterpolatingMultiStepSolver.integrate :28-47: one solve t0 -> t_last on a dup of y0
            std::vector<R> y(y0, y0 + N);
            InterpListener<R> il(t + 1, nt - 1, out + N, N);
            s.listeners.push_back(&il);
            s.integrate(eq, t[0], t[nt - 1], y.data(), N, 0);
        } else {
            // SingleSteppingMultiStepSolver.integrate :26-43
            for (int k = 1; k < nt; ++k) {
                std::memcpy(out + k * N, out + (k - 1) * N, sizeof(R) * N);
                s.integrate(eq, t[k - 1], t[k], out + k * N, N, k - 1);
            }
        }
        return 0;
    }

    // SingleStepAdjoint.solve :37-88.  a_inout: in dL/dz(t1), out dL/dz(t0).  grad: flat DL4J gradient
    // view (theta-adjoint is seeded from its current content, :59-60).  a_t: in/out scalar or null.
    void single_adjoint(Solver<R>& s, const R* z_t1, R* a_inout, R t0, R t1, R* grad_flat, R* a_t, int seg) {
        const int64_t nz = plan.n_state, nreal = plan.n_real, nt = a_t ? 1 : 0;
        AdjEq<R> eq(&net, nt);
        std::vector<R> zaug(eq.ntot, 0);
        std::memcpy(zaug.data(), z_t1, sizeof(R) * nz);
        std::memcpy(zaug.data() + nz, a_inout, sizeof(R) * nz);
        gather_real(grad_flat, zaug.data() + 2 * nz);
        if (a_t) zaug[2 * nz + nreal] = *a_t;
        s.integrate(eq, t1, t0, zaug.data(), eq.ntot, seg);          // Nd4j.reverse(time) :81
        std::memcpy(a_inout, zaug.data() + nz, sizeof(R) * nz);
        scatter_real(zaug.data() + 2 * nz, grad_flat);               // :85
        if (a_t) *a_t = zaug[2 * nz + nreal];
    }

    // FixedStepAdjoint / InputStepAdjoint -> SingleStepAdjoint (nt==2) or MultiStepAdjoint :49-90.
    // zt, dLdz: time-first [nt, state] when nt>2 (MultiStepAdjoint.alignInShapeToTimeFirst), else [state].
    // dLdt: [nt] (TG_ZERO/TG_CALC) or null.
    int backward(const R* params, const R* zt, const R* dLdz, const R* t, int nt, int tg_mode,
                 R* grad_flat, R* dLdy0, R* dLdt, OrcStats* st) {
        if (nt < 2) { err = "time must have at least two elements"; return 1; }
        net.theta = params; net.nfe = 0; log.clear();
        const int64_t N = plan.n_state;
        Solver<R> s; s.cfg = cfg; s.stats = st; s.log = &log;
        if (nt > 2) {   // MultiStepAdjoint.assertSorted :41-46
            int sgn = 0;
            for (int i = 0; i + 1 < nt; ++i) sgn += (t[i] > t[i + 1]) - (t[i] < t[i + 1]);
            if (std::abs(sgn) + 1 != nt) { err = "Time must be in ascending or descending order!"; return 2; }
        }
        std::vector<R> a(N), fz(N);
        if (nt == 2) {
            std::memcpy(a.data(), dLdz, sizeof(R) * N);
            R at = 0, d = 0;
            if (tg_mode == TG_CALC) {           // CalcTimeGrad.calcTimeAdjointT1 :52-58
                net.forward(zt, t[1], fz.data()); if (st) st->nfe += 1;
                double acc = 0; for (int64_t i = 0; i < N; ++i) acc += (double)dLdz[i] * (double)fz[i];
                d = (R)acc; at = R(R(0) - d);
            }
            single_adjoint(s, zt, a.data(), t[0], t[1], grad_flat, tg_mode == TG_CALC ? &at : nullptr, 0);
            std::memcpy(dLdy0, a.data(), sizeof(R) * N);
            if (dLdt) {
                if (tg_mode == TG_CALC) { dLdt[0] = at; dLdt[1] = d; }     // CalcTimeGrad.createLossGradient :66-74
                else { dLdt[0] = 0; dLdt[1] = 0; }                          // ZeroTimeGrad :45-55
            }
            return 0;
        }
        R last_time = 0;                      // CalcMultiStepTimeGrad.lastTime :18 (shared accumulator)
        if (dLdt) for (int i = 0; i < nt; ++i) dLdt[i] = 0;
        bool have_prev = false;
        for (int step = nt - 1; step > 0; --step) {
            const R* ztk = zt + (int64_t)step * N;
            const R* graw = dLdz + (int64_t)step * N;
            R d = 0;
            if (tg_mode == TG_CALC) {
                net.forward(ztk, t[step], fz.data()); if (st) st->nfe += 1;
                double acc = 0; for (int64_t i = 0; i < N; ++i) acc += (double)graw[i] * (double)fz[i];
                d = (R)acc;
                last_time = R(last_time - d);                             // startTime.subi(dL_dt1)
            }
            // prepareStep: dL_dzt += previous segment's a  (first segment: nothing to add)
            if (have_prev) for (int64_t i = 0; i < N; ++i) a[i] = R(graw[i] + a[i]);
            else std::memcpy(a.data(), graw, sizeof(R) * N);
            R at = last_time;
            single_adjoint(s, ztk, a.data(), t[step - 1], t[step], grad_flat, tg_mode == TG_CALC ? &at : nullptr, nt - 1 - step);
            if (tg_mode == TG_CALC && dLdt) { dLdt[step - 1] = at; dLdt[step] = d; }   // updateStep :51-53
            have_prev = true;
        }
        if (tg_mode == TG_CALC && dLdt) dLdt[0] = last_time;             // updateLastStep :56-62
        for (int64_t i = 0; i < N; ++i) dLdy0[i] = R(a[i] + dLdz[i]);    // += dL/dz(t0)
        return 0;
    }
};

struct AnyHandle {
    int use_double = 0;
    std::unique_ptr<Handle<float>> f;
    std::unique_ptr<Handle<double>> d;
    std::vector<OrcLayerDesc> layers;
};

thread_local std::string g_err;

}  // namespace

// ---------------------------------------------------------------------------------------------
// C interface (ctypes; see oracle/oracle.py)
// ---------------------------------------------------------------------------------------------
extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

void* orc_create(const OrcGraphDesc* g, const OrcSolverCfg* c, int use_double) {
    if (!(c->min_step < c->max_step)) {   // SolverConfig.java:22-24
        g_err = "Max step smaller than min step! Swapped arguments?";
        return nullptr;
    }
    auto* h = new AnyHandle();
    h->use_double = use_double;
    h->layers.assign(g->layers, g->layers + g->n_layers);
    OrcGraphDesc gg = *g; gg.layers = h->layers.data();
    NetPlan p;
    if (!plan_net(gg, p)) { g_err = p.err; delete h; return nullptr; }
    if (use_double) { h->d.reset(new Handle<double>()); h->d->plan = p; h->d->net.init(p); h->d->cfg = *c; }
    else { h->f.reset(new Handle<float>()); h->f->plan = p; h->f->net.init(p); h->f->cfg = *c; }
    return h;
}
void orc_destroy(void* hh) { delete (AnyHandle*)hh; }

int64_t orc_num_params(void* hh) { auto* h = (AnyHandle*)hh; return h->use_double ? h->d->plan.n_params : h->f->plan.n_params; }
int64_t orc_num_real_params(void* hh) { auto* h = (AnyHandle*)hh; return h->use_double ? h->d->plan.n_real : h->f->plan.n_real; }
// gradient w.r.t. the time input left by the last rhs_vjp / augmented evaluation (0 without a time-concat vertex)
double orc_last_dt(void* hh) { auto* h = (AnyHandle*)hh; return h->use_double ? (double)h->d->net.dt_last : (double)h->f->net.dt_last; }
int64_t orc_state_len(void* hh) { auto* h = (AnyHandle*)hh; return h->use_double ? h->d->plan.n_state : h->f->plan.n_state; }

int orc_forward(void* hh, const void* params, const void* y0, const void* t, int nt, int interpolate, void* out, OrcStats* st) {
    auto* h = (AnyHandle*)hh;
    if (st) *st = OrcStats{0, 0, 0, 0};
    int rc = h->use_double ? h->d->forward((const double*)params, (const double*)y0, (const double*)t, nt, interpolate, (double*)out, st)
                           : h->f->forward((const float*)params, (const float*)y0, (const float*)t, nt, interpolate, (float*)out, st);
    if (rc) g_err = h->use_double ? h->d->err : h->f->err;
    return rc;
}

int orc_backward(void* hh, const void* params, const void* zt, const void* dLdz, const void* t, int nt, int tg_mode,
                 void* grad_flat, void* dLdy0, void* dLdt, OrcStats* st) {
    auto* h = (AnyHandle*)hh;
    if (st) *st = OrcStats{0, 0, 0, 0};
    int rc = h->use_double
                 ? h->d->backward((const double*)params, (const double*)zt, (const double*)dLdz, (const double*)t, nt, tg_mode,
                                  (double*)grad_flat, (double*)dLdy0, (double*)dLdt, st)
                 : h->f->backward((const float*)params, (const float*)zt, (const float*)dLdz, (const float*)t, nt, tg_mode,
                                  (float*)grad_flat, (float*)dLdy0, (float*)dLdt, st);
    if (rc) g_err = h->use_double ? h->d->err : h->f->err;
    return rc;
}

int64_t orc_get_step_log(void* hh, OrcStepRec* buf, int64_t cap) {
    auto* h = (AnyHandle*)hh;
    auto& lg = h->use_double ? h->d->log : h->f->log;
    int64_t n = std::min<int64_t>(cap, (int64_t)lg.size());
    if (buf && n > 0) std::memcpy(buf, lg.data(), sizeof(OrcStepRec) * n);
    return (int64_t)lg.size();
}

// RHS and its vjp on their own (ForwardPassTest analogue; also used for layer parity of the CUDA path)
int orc_rhs(void* hh, const void* params, const void* z, double t, void* fz) {
    auto* h = (AnyHandle*)hh;
    if (h->use_double) { h->d->net.theta = (const double*)params; h->d->net.forward((const double*)z, t, (double*)fz); }
    else { h->f->net.theta = (const float*)params; h->f->net.forward((const float*)z, (float)t, (float*)fz); }
    return 0;
}
int orc_rhs_vjp(void* hh, const void* params, const void* z, double t, const void* g, void* fz, void* dz, void* dreal) {
    auto* h = (AnyHandle*)hh;
    if (h->use_double) {
        h->d->net.theta = (const double*)params; h->d->net.forward((const double*)z, t, (double*)fz);
        h->d->net.backward((const double*)g, (double*)dz, (double*)dreal);
    } else {
        h->f->net.theta = (const float*)params; h->f->net.forward((const float*)z, (float)t, (float*)fz);
        h->f->net.backward((const float*)g, (float*)dz, (float*)dreal);
    }
    return 0;
}

// --- solver internals exposed for the unit pins (DormandPrince54MseTest, AdaptiveRungeKuttaStepPolicyTest,
//     InterpolationTest) ---
double orc_error_norm(const void* K7xN, const void* y0, const void* y1, double h, int64_t N, double atol, double rtol, int use_double) {
    if (use_double) {
        const double* K[7]; for (int j = 0; j < 7; ++j) K[j] = (const double*)K7xN + j * N;
        return error_norm<double>(K, (const double*)y0, (const double*)y1, h, N, atol, rtol);
    }
    const float* K[7]; for (int j = 0; j < 7; ++j) K[j] = (const float*)K7xN + j * N;
    return error_norm<float>(K, (const float*)y0, (const float*)y1, (float)h, N, (float)atol, (float)rtol);
}
double orc_next_step(double h, double err, const OrcSolverCfg* c, int use_double) {
    return use_double ? next_step<double>(h, err, *c) : (double)next_step<float>((float)h, (float)err, *c);
}
// initial step for the handle's graph at (t0 -> t1), returns h; K0/K1 (optional) receive the two derivatives
double orc_init_step(void* hh, const void* params, const void* y, double t0, double t1, void* K0out, void* K1out) {
    auto* h = (AnyHandle*)hh;
    if (h->use_double) {
        auto& H = *h->d; H.net.theta = (const double*)params; int64_t N = H.plan.n_state;
        Solver<double> s; s.cfg = H.cfg; s.yW.assign(N, 0); std::vector<double> K0(N), K1(N);
        FwdEq<double> eq(&H.net); eq.eval((const double*)y, t0, K0.data());
        double r = s.initialize_step(eq, t0, !(t1 > t0), (const double*)y, K0.data(), K1.data(), N);
        if (K0out) std::memcpy(K0out, K0.data(), 8 * N);
        if (K1out) std::memcpy(K1out, K1.data(), 8 * N);
        return r;
    }
    auto& H = *h->f; H.net.theta = (const float*)params; int64_t N = H.plan.n_state;
    Solver<float> s; s.cfg = H.cfg; s.yW.assign(N, 0); std::vector<float> K0(N), K1(N);
    FwdEq<float> eq(&H.net); eq.eval((const float*)y, (float)t0, K0.data());
    double r = s.initialize_step(eq, (float)t0, !(t1 > t0), (const float*)y, K0.data(), K1.data(), N);
    if (K0out) std::memcpy(K0out, K0.data(), 4 * N);
    if (K1out) std::memcpy(K1out, K1.data(), 4 * N);
    return r;
}
void orc_interpolate(const void* y0, const void* y1, const void* ymid, const void* f0, const void* f1, double dt,
                     double t0, double t1, double t, void* out, int64_t N, int use_double) {
    if (use_double) {
        Interp<double> ip; ip.fit((const double*)y0, (const double*)y1, (const double*)ymid, (const double*)f0, (const double*)f1, dt, N);
        ip.eval(t0, t1, t, (double*)out, N);
    } else {
        Interp<float> ip; ip.fit((const float*)y0, (const float*)y1, (const float*)ymid, (const float*)f0, (const float*)f1, (float)dt, N);
        ip.eval(t0, t1, t, (float*)out, N);
    }
}
void orc_tableau(double* a36, double* b7, double* e7, double* c6, double* cmid7) {
    for (int s = 0; s < 6; ++s) for (int j = 0; j < 6; ++j) a36[s * 6 + j] = TAB_A[s][j];
    for (int j = 0; j < 7; ++j) { b7[j] = TAB_B[j]; e7[j] = TAB_E[j]; cmid7[j] = TAB_CMID[j]; }
    for (int j = 0; j < 6; ++j) c6[j] = TAB_C[j];
}
void orc_set_probe_noise(double rel) { g_probe_noise = rel; }
// launchers such as torchrun export OMP_NUM_THREADS=1 for their children: the CPU baseline asks for the host's cores explicitly
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int orc_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
