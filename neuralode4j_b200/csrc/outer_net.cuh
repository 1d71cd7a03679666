// outer_net.cuh — the network AROUND the ODE block, so that one full training step of the reference's MNIST ODE-net runs on the
// device ("the callers either side of the path").
//
// Replaces, for the model of examples/mnist/OdeNetModel.java:60-88, what DL4J's ComputationGraph.fit does around OdeVertex:
//   stem     examples/mnist/ResBlockStem.java:25-48   inputConv 3x3 (Truncate, bias) -> 2 x { BN-ReLU ; 1x1 stride 2  +  3x3 stride 2
//                                                     Same ReLU -> BN-ReLU -> 3x3 Same }                        28 -> 26 -> 13 -> 7
//   block    ode/vertex/impl/OdeVertex.java:111-140   Engine::forward / Engine::backward (the hot path, engine.cu)
//   output   examples/mnist/Output.java:27-35         BN-ReLU -> GlobalPooling(AVG) -> Dense(10) softmax + LossMCXENT
//   updater  examples/mnist/LayerUtil.java:30-47      Nesterovs(momentum) on the flat gradient view divided by the minibatch size;
//            BatchNorm running mean / var through the gradient view with a NoOp updater (OdeGraphHelper.java:105-119)
// Layer arithmetic is DL4J's (third party in the reference); conventions as restated in oracle/outer_oracle.py, which this file is
// checked against (tests/test_outer_net.py).  Tensors are NCHW fp32; every reduction accumulates exact fp32 products in double and
// rounds once, elementwise expressions keep the oracle's order (no FMA contraction), so the two agree to rounding.
// Two data paths: the EXACT path in this file (any channel count; always in parity mode) — plain direct CUDA-core kernels, every
// reduction in double — and, for the reference's 64 kernels in the default mode, the FAST path of outer_net_tc.cuh (NHWC, the stem's
// 64 -> 64 convolutions on the block's tcgen05 kernels), which OuterNet::train_step / output dispatch to (tc_).
#pragma once

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

#include "engine.cuh"
#include "outer_net_tc.cuh"

namespace n4j {

// ---------------------------------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------------------------------
struct ConvShape {
    int B, Ci, H, W, Co, K, S, Ho, Wo, pt, pl;
};

inline ConvShape make_conv_shape(int B, int Ci, int H, int W, int Co, int K, int S, bool same) {
    ConvShape c{B, Ci, H, W, Co, K, S, 0, 0, 0, 0};
    if (same) {
        c.Ho = (H + S - 1) / S; c.Wo = (W + S - 1) / S;
        const int th = std::max((c.Ho - 1) * S + K - H, 0), tw = std::max((c.Wo - 1) * S + K - W, 0);
        c.pt = th / 2; c.pl = tw / 2;
    } else {
        c.Ho = (H - K) / S + 1; c.Wo = (W - K) / S + 1;
    }
    return c;
}

// y[b,co,oh,ow] = act( (float)sum_{ci,kh,kw} x[b,ci,oh*S+kh-pt,ow*S+kw-pl] * w[co,ci,kh,kw]  + bias[co] )
// One thread = one output pixel x OUTER_CT output channels (the x value is loaded once per (ci, tap) and used OUTER_CT times); the
// weights of the block's channel tile sit in shared memory and are read as warp broadcasts.  blockIdx.y = channel tile.
constexpr int OUTER_CT = 8;
__global__ void __launch_bounds__(256) k_outer_conv_fwd(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                         float* __restrict__ y, ConvShape c, int relu) {
    extern __shared__ float s_w[];                       // [OUTER_CT][Ci*K*K]
    const int co0 = blockIdx.y * OUTER_CT, wlen = c.Ci * c.K * c.K;
    for (int i = threadIdx.x; i < OUTER_CT * wlen; i += blockDim.x) {
        const int t = i / wlen;
        s_w[i] = co0 + t < c.Co ? w[(long long)(co0 + t) * wlen + (i - t * wlen)] : 0.f;
    }
    __syncthreads();
    const long long n = (long long)c.B * c.Ho * c.Wo;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int ow = (int)(i % c.Wo), oh = (int)((i / c.Wo) % c.Ho), b = (int)(i / ((long long)c.Wo * c.Ho));
        double acc[OUTER_CT];
#pragma unroll
        for (int t = 0; t < OUTER_CT; ++t) acc[t] = 0.0;
        for (int ci = 0; ci < c.Ci; ++ci) {
            const float* xb = x + ((long long)b * c.Ci + ci) * c.H * c.W;
            for (int kh = 0; kh < c.K; ++kh) {
                const int ih = oh * c.S + kh - c.pt;
                if (ih < 0 || ih >= c.H) continue;
                for (int kw = 0; kw < c.K; ++kw) {
                    const int iw = ow * c.S + kw - c.pl;
                    if (iw < 0 || iw >= c.W) continue;
                    const double xv = (double)xb[ih * c.W + iw];
                    const float* wp = s_w + (ci * c.K + kh) * c.K + kw;
#pragma unroll
                    for (int t = 0; t < OUTER_CT; ++t) acc[t] += xv * (double)wp[t * wlen];
                }
            }
        }
#pragma unroll
        for (int t = 0; t < OUTER_CT; ++t) {
            if (co0 + t >= c.Co) break;
            float v = (float)acc[t];
            if (bias) v = __fadd_rn(v, bias[co0 + t]);
            if (relu && !(v > 0.f)) v = 0.f;
            y[(((long long)b * c.Co + co0 + t) * c.Ho + oh) * c.Wo + ow] = v;
        }
    }
}

// dx[b,ci,h,w] = (float)sum_{co,kh,kw} dz[b,co,oh,ow] * w[co,ci,kh,kw],  oh*S + kh - pt = h;  dz = relu ? (y > 0 ? dy : 0) : dy
// One thread = one input pixel x OUTER_CT input channels; blockIdx.y = input-channel tile, its weights [Co][OUTER_CT][K*K] in shared memory.
__global__ void __launch_bounds__(256) k_outer_conv_dgrad(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ w,
                                                           float* __restrict__ dx, ConvShape c, int relu) {
    extern __shared__ float s_w[];                       // [Co][OUTER_CT][K*K]
    const int ci0 = blockIdx.y * OUTER_CT, kk = c.K * c.K;
    for (int i = threadIdx.x; i < c.Co * OUTER_CT * kk; i += blockDim.x) {
        const int co = i / (OUTER_CT * kk), r = i - co * OUTER_CT * kk, t = r / kk, k = r - t * kk;
        s_w[i] = ci0 + t < c.Ci ? w[((long long)co * c.Ci + ci0 + t) * kk + k] : 0.f;
    }
    __syncthreads();
    // pixels are enumerated parity class by parity class ((ih % S, iw % S) outermost): the lanes of a warp then share the set of taps that
    // reach them, so a strided convolution's tap tests do not diverge
    const int Hc = (c.H + c.S - 1) / c.S, Wc = (c.W + c.S - 1) / c.S;
    const long long cls = (long long)c.B * Hc * Wc, n = cls * c.S * c.S;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int pc = (int)(i / cls);
        const long long j = i - (long long)pc * cls;
        const int iw = (int)(j % Wc) * c.S + pc % c.S, ih = (int)((j / Wc) % Hc) * c.S + pc / c.S, b = (int)(j / ((long long)Wc * Hc));
        if (ih >= c.H || iw >= c.W) continue;
        double acc[OUTER_CT];
#pragma unroll
        for (int t = 0; t < OUTER_CT; ++t) acc[t] = 0.0;
        for (int kh = 0; kh < c.K; ++kh) {
            const int th = ih + c.pt - kh;
            if (th < 0 || th % c.S) continue;
            const int oh = th / c.S;
            if (oh >= c.Ho) continue;
            for (int kw = 0; kw < c.K; ++kw) {
                const int tw = iw + c.pl - kw;
                if (tw < 0 || tw % c.S) continue;
                const int ow = tw / c.S;
                if (ow >= c.Wo) continue;
                for (int co = 0; co < c.Co; ++co) {
                    const long long o = (((long long)b * c.Co + co) * c.Ho + oh) * c.Wo + ow;
                    float g = dy[o];
                    if (relu && !(y[o] > 0.f)) g = 0.f;
                    if (g == 0.f) continue;
                    const double gd = (double)g;
                    const float* wp = s_w + co * OUTER_CT * kk + kh * c.K + kw;
#pragma unroll
                    for (int t = 0; t < OUTER_CT; ++t) acc[t] += gd * (double)wp[t * kk];
                }
            }
        }
#pragma unroll
        for (int t = 0; t < OUTER_CT; ++t)
            if (ci0 + t < c.Ci) dx[(((long long)b * c.Ci + ci0 + t) * c.H + ih) * c.W + iw] = (float)acc[t];
    }
}

// Weight gradient.  grid = (Co, ceil(Ci / 16), OUTER_WSPLIT), block = 16 warps: one WARP per (co, ci, batch slice).  The lanes stride over
// the slice's output pixels — dz and x are read along their contiguous dimension — with K*K (<= 9) double accumulators each, combined
// by a shuffle butterfly (fixed order: deterministic).  part[slice][co][ci][tap] / part_b[slice][co] hold double partials;
// k_outer_wgrad_finish adds the slices in a fixed order.
constexpr int OUTER_WSPLIT = 8;
__global__ void __launch_bounds__(512) k_outer_conv_wgrad(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ y,
                                                           double* __restrict__ part, double* __restrict__ part_b, ConvShape c, int relu) {
    extern __shared__ int s_tab[];                       // per output pixel of one image: [0] offset of its top-left input pixel, [1] tap validity mask
    const int co = blockIdx.x, slice = blockIdx.z, kk = c.K * c.K, nout = c.Ci * kk;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per = c.Ho * c.Wo;
    for (int r = threadIdx.x; r < per; r += blockDim.x) {
        const int oh = r / c.Wo, ow = r - oh * c.Wo;
        const int ih0 = oh * c.S - c.pt, iw0 = ow * c.S - c.pl;
        int mask = 0;
        for (int kh = 0; kh < c.K; ++kh)
            for (int kw = 0; kw < c.K; ++kw)
                if (ih0 + kh >= 0 && ih0 + kh < c.H && iw0 + kw >= 0 && iw0 + kw < c.W) mask |= 1 << (kh * 3 + kw);
        s_tab[2 * r] = ih0 * c.W + iw0;
        s_tab[2 * r + 1] = mask;
    }
    __syncthreads();
    const int ci = blockIdx.y * 16 + warp;
    if (ci >= c.Ci) return;
    const int b0 = (int)((long long)c.B * slice / OUTER_WSPLIT), b1 = (int)((long long)c.B * (slice + 1) / OUTER_WSPLIT);
    double acc[9], accb = 0.0;
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] = 0.0;
    for (int b = b0; b < b1; ++b) {
        const float* dyb = dy + ((long long)b * c.Co + co) * per;
        const float* yb = y + ((long long)b * c.Co + co) * per;
        const float* xb = x + ((long long)b * c.Ci + ci) * c.H * c.W;
        for (int r = lane; r < per; r += 32) {
            float g = dyb[r];
            if (relu && !(yb[r] > 0.f)) g = 0.f;
            if (g == 0.f) continue;
            const double gd = (double)g;
            accb += gd;
            const int off = s_tab[2 * r], mask = s_tab[2 * r + 1];
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                for (int kw = 0; kw < 3; ++kw)
                    if (mask & (1 << (kh * 3 + kw))) acc[kh * 3 + kw] += gd * (double)xb[off + kh * c.W + kw];
        }
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] = warp_sum(acc[k]);
    accb = warp_sum(accb);
    if (lane == 0) {
        for (int kh = 0; kh < c.K; ++kh)
            for (int kw = 0; kw < c.K; ++kw) part[((long long)slice * c.Co + co) * nout + ci * kk + kh * c.K + kw] = acc[kh * 3 + kw];
        if (part_b && ci == 0) part_b[slice * c.Co + co] = accb;
    }
}
__global__ void __launch_bounds__(256) k_outer_wgrad_finish(const double* __restrict__ part, const double* __restrict__ part_b, float* __restrict__ dw,
                                                             float* __restrict__ db, long long n, int Co) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n + (db ? Co : 0); i += (long long)gridDim.x * blockDim.x) {
        double s = 0.0;
        if (i < n) {
            for (int q = 0; q < OUTER_WSPLIT; ++q) s += part[q * n + i];
            dw[i] = (float)s;
        } else {
            for (int q = 0; q < OUTER_WSPLIT; ++q) s += part_b[q * Co + (i - n)];
            db[i - n] = (float)s;
        }
    }
}

// BatchNorm (training), NCHW, one block per channel.  st: [mean | invstd | batch var] x C
__global__ void __launch_bounds__(256) k_outer_bn_fwd(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       float* __restrict__ y, float* __restrict__ st, int B, int C, int HW, float eps, int relu) {
    const int ch = blockIdx.x;
    double s1 = 0.0, s2 = 0.0;
    const long long n = (long long)B * HW;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const double v = (double)x[((i / HW) * C + ch) * HW + i % HW];
        s1 += v; s2 += v * v;
    }
    double r[2] = {s1, s2};
    block_sum<2>(r);
    __shared__ float s_m, s_i;
    if (threadIdx.x == 0) {
        const double m = r[0] / (double)n;
        double var = r[1] / (double)n - m * m;
        if (var < 0) var = 0;
        s_m = (float)m; s_i = (float)(1.0 / sqrt(var + (double)eps));
        st[ch] = s_m; st[C + ch] = s_i; st[2 * C + ch] = (float)var;
    }
    __syncthreads();
    const float m = s_m, inv = s_i, ga = gamma[ch], be = beta[ch];
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const long long o = ((i / HW) * C + ch) * HW + i % HW;
        float v = __fadd_rn(__fmul_rn(ga, __fmul_rn(__fsub_rn(x[o], m), inv)), be);
        if (relu && !(v > 0.f)) v = 0.f;
        y[o] = v;
    }
}
// inference: y = act(gamma * (x - running mean) / sqrt(running var + eps) + beta)
__global__ void __launch_bounds__(256) k_outer_bn_infer(const float* __restrict__ x, const float* __restrict__ bn /*gamma,beta,mean,var*/, float* __restrict__ y,
                                                         long long n, int C, int HW, float eps, int relu) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)((i / HW) % C);
        const float inv = (float)(1.0 / sqrt((double)bn[3 * C + ch] + (double)eps));
        float v = __fadd_rn(__fmul_rn(bn[ch], __fmul_rn(__fsub_rn(x[i], bn[2 * C + ch]), inv)), bn[C + ch]);
        if (relu && !(v > 0.f)) v = 0.f;
        y[i] = v;
    }
}
// backward; also the running-statistics "gradient" (run - batch) * (1 - decay) into the mean / var slots of the gradient view
__global__ void __launch_bounds__(256) k_outer_bn_bwd(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ dy,
                                                       const float* __restrict__ bn /*gamma,beta,mean,var (params)*/, const float* __restrict__ st,
                                                       float* __restrict__ dx, float* __restrict__ gbn /*gradient view of this layer*/, int B, int C, int HW,
                                                       int relu, float one_minus_decay) {
    const int ch = blockIdx.x;
    const float m = st[ch], inv = st[C + ch];
    const long long n = (long long)B * HW;
    double s1 = 0.0, s2 = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const long long o = ((i / HW) * C + ch) * HW + i % HW;
        float g = dy[o];
        if (relu && !(y[o] > 0.f)) g = 0.f;
        const float xh = __fmul_rn(__fsub_rn(x[o], m), inv);
        s1 += (double)g; s2 += (double)g * (double)xh;
    }
    double r[2] = {s1, s2};
    block_sum<2>(r);
    __shared__ float s_m1, s_m2;
    if (threadIdx.x == 0) {
        gbn[ch] = (float)r[1]; gbn[C + ch] = (float)r[0];
        gbn[2 * C + ch] = __fmul_rn(__fsub_rn(bn[2 * C + ch], m), one_minus_decay);
        gbn[3 * C + ch] = __fmul_rn(__fsub_rn(bn[3 * C + ch], st[2 * C + ch]), one_minus_decay);
        s_m1 = (float)(r[0] / (double)n); s_m2 = (float)(r[1] / (double)n);
    }
    __syncthreads();
    const float m1 = s_m1, m2 = s_m2, k = __fmul_rn(bn[ch], inv);
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const long long o = ((i / HW) * C + ch) * HW + i % HW;
        float g = dy[o];
        if (relu && !(y[o] > 0.f)) g = 0.f;
        const float xh = __fmul_rn(__fsub_rn(x[o], m), inv);
        dx[o] = __fmul_rn(k, __fsub_rn(__fsub_rn(g, m1), __fmul_rn(xh, m2)));
    }
}

__global__ void __launch_bounds__(256) k_outer_add(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) o[i] = __fadd_rn(a[i], b[i]);
}

// global average pooling: one warp per (b, c)
__global__ void __launch_bounds__(256) k_outer_pool_fwd(const float* __restrict__ x, float* __restrict__ p, int BC, int HW) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= BC) return;
    double s = 0.0;
    for (int i = lane; i < HW; i += 32) s += (double)x[(long long)w * HW + i];
    s = warp_sum(s);
    if (lane == 0) p[w] = (float)(s / (double)HW);
}
__global__ void __launch_bounds__(256) k_outer_pool_bwd(const float* __restrict__ dp, float* __restrict__ dx, long long n, int HW) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dx[i] = __fdiv_rn(dp[i / HW], (float)HW);
}

// OutputLayer: logits = pool W + b (W [C,10] f-order), softmax, LossMCXENT score; one thread per example, then one block reduces the score
__global__ void __launch_bounds__(256) k_outer_output_fwd(const float* __restrict__ pool, const float* __restrict__ W, const float* __restrict__ bias,
                                                           const float* __restrict__ labels, float* __restrict__ probs, double* __restrict__ score_parts, int B, int C) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float lg[10];
    float mx = -INFINITY;
    for (int j = 0; j < 10; ++j) {
        double acc = 0.0;
        for (int c = 0; c < C; ++c) acc += (double)pool[(long long)b * C + c] * (double)W[c + (long long)C * j];
        lg[j] = __fadd_rn((float)acc, bias[j]);
        mx = fmaxf(mx, lg[j]);
    }
    double se = 0.0;
    for (int j = 0; j < 10; ++j) { lg[j] = expf(__fsub_rn(lg[j], mx)); se += (double)lg[j]; }
    const float den = (float)se;
    double sc = 0.0;
    for (int j = 0; j < 10; ++j) {
        const float p = __fdiv_rn(lg[j], den);
        probs[b * 10 + j] = p;
        if (labels) {
            double pc = (double)p;
            pc = pc < 1e-10 ? 1e-10 : (pc > 1.0 - 1e-10 ? 1.0 - 1e-10 : pc);
            sc -= (double)labels[b * 10 + j] * log(pc);
        }
    }
    if (score_parts) score_parts[b] = sc;
}
// backward of the output layer: single block (10 x C weights, B examples)
__global__ void __launch_bounds__(256) k_outer_output_bwd(const float* __restrict__ pool, const float* __restrict__ W, const float* __restrict__ probs,
                                                           const float* __restrict__ labels, const double* __restrict__ score_parts, float* __restrict__ dW,
                                                           float* __restrict__ db, float* __restrict__ dpool, float* __restrict__ score, int B, int C) {
    for (int i = threadIdx.x; i < C * 10; i += blockDim.x) {       // dW[c + C*j] = sum_b pool[b,c] * (p - y)[b,j]
        const int c = i % C, j = i / C;
        double acc = 0.0;
        for (int b = 0; b < B; ++b) acc += (double)pool[(long long)b * C + c] * (double)__fsub_rn(probs[b * 10 + j], labels[b * 10 + j]);
        dW[i] = (float)acc;
    }
    for (int j = threadIdx.x; j < 10; j += blockDim.x) {
        double acc = 0.0;
        for (int b = 0; b < B; ++b) acc += (double)__fsub_rn(probs[b * 10 + j], labels[b * 10 + j]);
        db[j] = (float)acc;
    }
    for (int i = threadIdx.x; i < B * C; i += blockDim.x) {        // dpool[b,c] = sum_j (p - y)[b,j] * W[c,j]
        const int b = i / C, c = i % C;
        double acc = 0.0;
        for (int j = 0; j < 10; ++j) acc += (double)__fsub_rn(probs[b * 10 + j], labels[b * 10 + j]) * (double)W[c + (long long)C * j];
        dpool[i] = (float)acc;
    }
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int b = 0; b < B; ++b) s += score_parts[b];
        *score = (float)(s / (double)B);
    }
}

// Updater.  kind 0: Nesterovs on g / B  (v' = mu v - lr g/B;  p += -mu v + (1 + mu) v');  kind 1: NoOp — the "gradient" is subtracted as is
// (BatchNorm running statistics).
__global__ void __launch_bounds__(256) k_outer_update(float* __restrict__ p, float* __restrict__ v, const float* __restrict__ g, const unsigned char* __restrict__ kind,
                                                       long long n, float batch, float lr, float mu, float one_plus_mu) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (kind[i]) { p[i] = __fsub_rn(p[i], g[i]); continue; }
        const float gb = __fdiv_rn(g[i], batch);
        const float vo = v[i];
        const float vn = __fsub_rn(__fmul_rn(mu, vo), __fmul_rn(lr, gb));
        p[i] = __fadd_rn(p[i], __fadd_rn(__fmul_rn(-mu, vo), __fmul_rn(one_plus_mu, vn)));
        v[i] = vn;
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// the network
// ---------------------------------------------------------------------------------------------------------------------------
class OuterNet {
public:
    OuterNet(int batch, int channels, const n4j_solver_cfg& cfg, float t0, float t1, int device);
    ~OuterNet();
    long long num_params() const { return n_; }
    long long ode_offset() const { return off_ode_; }
    long long ode_num_params() const { return n_ode_; }
    // one training step: params / updater state are read and written in place (host or device pointers)
    void train_step(float* params, float* updater_state, const float* images, const float* labels, float lr, float mu, float* score,
                    float* grad_out, n4j_stats* ode_fwd, n4j_stats* ode_bwd);
    // forward only; train != 0: BatchNorm batch statistics everywhere (what fit() sees), else running statistics in the outer layers
    void output(const float* params, const float* images, float* probs, int train);

private:
    struct Block {
        ConvShape ds, c1, c2;
        long long o_bn1, o_ds, o_c1, o_bn2, o_c2;     // parameter offsets
        float *n1, *dsy, *f1, *n2, *f2, *a_out, *st1, *st2;
        int Hin, Hout;
    };
    void forward(const float* p, const float* images, const float* labels, bool train, n4j_stats* ode_fwd);
    void backward(n4j_stats* ode_bwd);
    float* dev_alloc(long long n);
    void conv_fwd(const float* x, const float* w, const float* bias, float* y, const ConvShape& c, int relu);
    void conv_dgrad(const float* dy, const float* y, const float* w, float* dx, const ConvShape& c, int relu);
    void conv_wgrad(const float* x, const float* dy, const float* y, float* dw, float* db, const ConvShape& c, int relu);
    bool is_dev(const void* p) const;
    int grid(long long n) const { return (int)std::max<long long>(1, std::min<long long>((n + 255) / 256, 148LL * 8)); }

    int device_, B_, C_;
    float time_[2];
    std::unique_ptr<Engine> ode_;
    cudaStream_t s_ = nullptr;
    ConvShape cin_;
    Block blk_[2];
    long long n_ = 0, off_in_b_ = 0, off_in_w_ = 0, off_ode_ = 0, n_ode_ = 0, off_bno_ = 0, off_ow_ = 0, off_ob_ = 0;
    float *p_ = nullptr, *v_ = nullptr, *g_ = nullptr;     // device copies: params, updater state, gradient view
    unsigned char* kind_ = nullptr;
    float *img_ = nullptr, *lab_ = nullptr, *a0_ = nullptr, *z1_ = nullptr, *no_ = nullptr, *sto_ = nullptr, *pool_ = nullptr, *probs_ = nullptr;
    float *d1_ = nullptr, *d2_ = nullptr, *d3_ = nullptr, *dpool_ = nullptr, *score_ = nullptr, *tdev_ = nullptr;
    double* score_parts_ = nullptr;
    double *wpart_ = nullptr, *wpart_b_ = nullptr;      // weight-gradient partials of the outer convolutions [OUTER_WSPLIT][Co][Ci*K*K]
    std::vector<void*> allocs_;

    // ---- fast path (outer_net_tc.cuh): 64 kernels, default mode — NHWC activations, the 64 -> 64 convolutions on the tcgen05 kernels ----
    struct TcBlock {
        ConvGeom gi, go;               // operand-image geometries at the block's input / output resolution
        WgradTcGeom wi, wo;
        int sub_off;                   // sampled pixel of the stride-2 firstConv: h = 2 * oh + sub_off
        float *n1, *n1img, *dsy, *f1, *n2, *n2img, *f2, *a_out;
        float *w1f, *w1b, *w2f, *w2b;  // per-tap hi/lo weight images, forward / dgrad
        float *t1, *gup, *gupimg, *dn1, *daimg, *da_in;
        double *acc1, *acc2;
    };
    bool tc_ = false;
    int passes_ = 3;
    TcBlock tb_[2];
    float *a0n_ = nullptr, *z1n_ = nullptr, *non_ = nullptr, *dA_ = nullptr, *dB_ = nullptr, *dC_ = nullptr;
    double *acc_pool_ = nullptr, *acco_ = nullptr, *dspart_ = nullptr, *inpart_ = nullptr;
    size_t acc_pool_bytes_ = 0;
    float *wpart_tc_ = nullptr, *wmid_ = nullptr;
    size_t conv_smem_max_ = 0, wgrad_smem_max_ = 0;
    static constexpr int WMID = 32;
    // side branch: the weight gradients (and downSample's forward) are off the chain of dependent kernels — they run on a second stream
    // beside the dgrad / BatchNorm chain, like the block's own weight gradients (engine.cu: side_)
    cudaStream_t s2_ = nullptr;
    cudaEvent_t ev_fork_[8] = {}, ev_join_ = nullptr, ev_ds_[2] = {};
    int fork_n_ = 0;
    void fork_side();      // side branch continues from here (everything enqueued on s_ so far is visible to it)
    void join_side();
    void init_tc();
    void set_tc_attributes();
    void conv_tc(const float* img, const float* wimg, float* out, const ConvGeom& g, int act, bool sub, int sub_off, int Ho);
    void wgrad_tc(cudaStream_t st, const float* x, const float* g, const WgradTcGeom& wg, float* dw_flat);
    void forward_tc(const float* p, const float* images, const float* labels, bool train, n4j_stats* ode_fwd);
    void backward_tc(const float* images, n4j_stats* ode_bwd);
    int grid_rows(long long rows) const { return (int)std::max<long long>(1, std::min<long long>((rows + 15) / 16, 148LL * 4)); }
};

inline float* OuterNet::dev_alloc(long long n) {
    float* p = nullptr;
    N4J_CUDA(cudaMalloc(&p, sizeof(float) * std::max<long long>(n, 1)));
    allocs_.push_back(p);
    return p;
}
inline void OuterNet::conv_fwd(const float* x, const float* w, const float* bias, float* y, const ConvShape& c, int relu) {
    const long long n = (long long)c.B * c.Ho * c.Wo;
    const dim3 g((unsigned)std::max<long long>(1, std::min<long long>((n + 255) / 256, 148LL * 4)), (unsigned)((c.Co + OUTER_CT - 1) / OUTER_CT));
    k_outer_conv_fwd<<<g, 256, sizeof(float) * OUTER_CT * c.Ci * c.K * c.K, s_>>>(x, w, bias, y, c, relu);
}
inline void OuterNet::conv_dgrad(const float* dy, const float* y, const float* w, float* dx, const ConvShape& c, int relu) {
    const long long n = (long long)c.B * c.H * c.W;
    const dim3 g((unsigned)std::max<long long>(1, std::min<long long>((n + 255) / 256, 148LL * 4)), (unsigned)((c.Ci + OUTER_CT - 1) / OUTER_CT));
    k_outer_conv_dgrad<<<g, 256, sizeof(float) * c.Co * OUTER_CT * c.K * c.K, s_>>>(dy, y, w, dx, c, relu);
}
inline void OuterNet::conv_wgrad(const float* x, const float* dy, const float* y, float* dw, float* db, const ConvShape& c, int relu) {
    if (c.K > 3) throw Error{N4J_ERR_UNSUPPORTED, "outer convolutions: kernel size up to 3"};
    const int nout = c.Ci * c.K * c.K;
    k_outer_conv_wgrad<<<dim3(c.Co, (c.Ci + 15) / 16, OUTER_WSPLIT), 512, sizeof(int) * 2 * c.Ho * c.Wo, s_>>>(x, dy, y, wpart_, db ? wpart_b_ : nullptr, c, relu);
    const long long n = (long long)c.Co * nout;
    k_outer_wgrad_finish<<<grid(n + c.Co), 256, 0, s_>>>(wpart_, wpart_b_, dw, db, n, c.Co);
}
inline bool OuterNet::is_dev(const void* p) const {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

inline OuterNet::OuterNet(int batch, int channels, const n4j_solver_cfg& cfg, float t0, float t1, int device) : device_(device), B_(batch), C_(channels) {
    if (batch < 1 || channels < 1) throw Error{N4J_ERR_INVALID_ARGUMENT, "batch and channels must be positive"};
    N4J_CUDA(cudaSetDevice(device));
    time_[0] = t0; time_[1] = t1;
    const int C = C_;
    // the ODE block: BN-ReLU, conv3x3, BN-ReLU, conv3x3, BN-ReLU on [B,C,7,7] (examples/mnist/OdeNetModel.java:72-88)
    n4j_layer_desc L[5];
    for (int i = 0; i < 5; ++i) {
        L[i] = n4j_layer_desc{};
        L[i].kind = (i % 2 == 0) ? N4J_LAYER_BATCHNORM : N4J_LAYER_CONV3X3;
        L[i].activation = (i % 2 == 0) ? N4J_ACT_RELU : N4J_ACT_IDENTITY;
        L[i].n_in = C; L[i].n_out = C; L[i].has_bias = 0; L[i].eps = 1e-5f;
    }
    n4j_graph_desc g{};
    g.n_layers = 5; g.rank = 4; g.shape[0] = batch; g.shape[1] = C; g.shape[2] = 7; g.shape[3] = 7; g.layers = L;
    ode_.reset(new Engine(g, cfg, device));
    s_ = ode_->stream();
    n_ode_ = ode_->num_params();
    // parameter layout (DL4J order; oracle/outer_oracle.py: OuterLayout)
    long long off = 0;
    off_in_b_ = off; off += C; off_in_w_ = off; off += (long long)C * 9;
    cin_ = make_conv_shape(batch, 1, 28, 28, C, 3, 1, false);
    int H = cin_.Ho;
    for (int r = 0; r < 2; ++r) {
        Block& b = blk_[r];
        b.Hin = H;
        b.o_bn1 = off; off += 4 * C;
        b.o_ds = off; off += (long long)C * C;
        b.o_c1 = off; off += (long long)C * C * 9;
        b.o_bn2 = off; off += 4 * C;
        b.o_c2 = off; off += (long long)C * C * 9;
        b.ds = make_conv_shape(batch, C, H, H, C, 1, 2, false);
        b.c1 = make_conv_shape(batch, C, H, H, C, 3, 2, true);
        b.Hout = b.c1.Ho;
        b.c2 = make_conv_shape(batch, C, b.Hout, b.Hout, C, 3, 1, true);
        if (b.ds.Ho != b.Hout) throw Error{N4J_ERR_ILLEGAL_STATE, "stem branches disagree on the downsampled size"};
        const long long nin = (long long)batch * C * H * H, nout = (long long)batch * C * b.Hout * b.Hout;
        b.n1 = dev_alloc(nin); b.dsy = dev_alloc(nout); b.f1 = dev_alloc(nout); b.n2 = dev_alloc(nout); b.f2 = dev_alloc(nout); b.a_out = dev_alloc(nout);
        b.st1 = dev_alloc(3 * C); b.st2 = dev_alloc(3 * C);
        H = b.Hout;
    }
    if (H != 7) throw Error{N4J_ERR_ILLEGAL_STATE, "stem does not end at 7x7"};
    off_ode_ = off; off += n_ode_;
    off_bno_ = off; off += 4 * C;
    off_ow_ = off; off += (long long)C * 10; off_ob_ = off; off += 10;
    n_ = off;
    p_ = dev_alloc(n_); v_ = dev_alloc(n_); g_ = dev_alloc(n_);
    N4J_CUDA(cudaMalloc(&kind_, n_)); allocs_.push_back(kind_);
    std::vector<unsigned char> kind(n_, 0);
    auto noop = [&](long long o) { for (int i = 2 * C; i < 4 * C; ++i) kind[o + i] = 1; };
    for (int r = 0; r < 2; ++r) { noop(blk_[r].o_bn1); noop(blk_[r].o_bn2); }
    noop(off_bno_);
    N4J_CUDA(cudaMemcpy(kind_, kind.data(), n_, cudaMemcpyHostToDevice));
    const long long n0 = (long long)batch * C * 26 * 26, nz = (long long)batch * C * 49;
    img_ = dev_alloc((long long)batch * 28 * 28); lab_ = dev_alloc((long long)batch * 10);
    a0_ = dev_alloc(n0); z1_ = dev_alloc(nz); no_ = dev_alloc(nz); sto_ = dev_alloc(3 * C);
    pool_ = dev_alloc((long long)batch * C); probs_ = dev_alloc((long long)batch * 10); dpool_ = dev_alloc((long long)batch * C);
    d1_ = dev_alloc(n0); d2_ = dev_alloc(n0); d3_ = dev_alloc(n0);
    score_ = dev_alloc(1); tdev_ = dev_alloc(2);
    N4J_CUDA(cudaMalloc(&score_parts_, sizeof(double) * batch)); allocs_.push_back(score_parts_);
    N4J_CUDA(cudaMalloc(&wpart_, sizeof(double) * OUTER_WSPLIT * (size_t)C * C * 9)); allocs_.push_back(wpart_);
    N4J_CUDA(cudaMalloc(&wpart_b_, sizeof(double) * OUTER_WSPLIT * (size_t)C)); allocs_.push_back(wpart_b_);
    N4J_CUDA(cudaMemcpy(tdev_, time_, sizeof(time_), cudaMemcpyHostToDevice));
    // fast path: the reference's 64 kernels, tensor cores allowed, not the parity mode (which keeps the exact-accumulation kernels above)
    // (N4J_NET_NO_TC=1 / N4J_NET_FORCE_TC=1: A/B switches — the second one keeps the fast stem next to an exact-mode ODE block, which
    // isolates the stem's own deviation from the block's gradient sensitivity)
    tc_ = channels == 64 && !(cfg.flags & (N4J_FLAG_EXACT_CONTRACTIONS | N4J_FLAG_NO_TENSOR)) && getenv("N4J_NET_NO_TC") == nullptr;
    if (channels == 64 && getenv("N4J_NET_FORCE_TC") != nullptr) tc_ = true;
    passes_ = (cfg.flags & N4J_FLAG_FAST_TF32) ? 1 : 3;
    if (tc_) init_tc();
}

inline OuterNet::~OuterNet() {
    cudaSetDevice(device_);
    if (s2_) {
        cudaStreamSynchronize(s2_); cudaStreamDestroy(s2_);
        for (auto& e : ev_fork_) cudaEventDestroy(e);
        for (auto& e : ev_ds_) cudaEventDestroy(e);
        cudaEventDestroy(ev_join_);
    }
    for (void* p : allocs_) cudaFree(p);
}

inline void OuterNet::forward(const float* p, const float* images, const float* labels, bool train, n4j_stats* ode_fwd) {
    const int C = C_, B = B_;
    conv_fwd(images, p + off_in_w_, p + off_in_b_, a0_, cin_, 0);
    const float* a = a0_;
    for (int r = 0; r < 2; ++r) {
        Block& b = blk_[r];
        const int HWi = b.Hin * b.Hin, HWo = b.Hout * b.Hout;
        const long long nin = (long long)B * C * HWi, nout = (long long)B * C * HWo;
        if (train) k_outer_bn_fwd<<<C, 256, 0, s_>>>(a, p + b.o_bn1, p + b.o_bn1 + C, b.n1, b.st1, B, C, HWi, 1e-5f, 1);
        else k_outer_bn_infer<<<grid(nin), 256, 0, s_>>>(a, p + b.o_bn1, b.n1, nin, C, HWi, 1e-5f, 1);
        conv_fwd(b.n1, p + b.o_ds, nullptr, b.dsy, b.ds, 0);
        conv_fwd(b.n1, p + b.o_c1, nullptr, b.f1, b.c1, 1);      // ReLU: examples/mnist/LayerUtil.java:111-116
        if (train) k_outer_bn_fwd<<<C, 256, 0, s_>>>(b.f1, p + b.o_bn2, p + b.o_bn2 + C, b.n2, b.st2, B, C, HWo, 1e-5f, 1);
        else k_outer_bn_infer<<<grid(nout), 256, 0, s_>>>(b.f1, p + b.o_bn2, b.n2, nout, C, HWo, 1e-5f, 1);
        conv_fwd(b.n2, p + b.o_c2, nullptr, b.f2, b.c2, 0);
        k_outer_add<<<grid(nout), 256, 0, s_>>>(b.dsy, b.f2, b.a_out, nout);                          // ElementWiseVertex(Add): stemAdd_r
        a = b.a_out;
    }
    N4J_CUDA(cudaGetLastError());
    // the ODE block (training mode always: SingleStep.java:37)
    ode_->bind_params(p + off_ode_, n_ode_);
    ode_->forward(a, time_, 2, 0, z1_, ode_fwd);
    const long long nz = (long long)B * C * 49;
    if (train) k_outer_bn_fwd<<<C, 256, 0, s_>>>(z1_, p + off_bno_, p + off_bno_ + C, no_, sto_, B, C, 49, 1e-5f, 1);
    else k_outer_bn_infer<<<grid(nz), 256, 0, s_>>>(z1_, p + off_bno_, no_, nz, C, 49, 1e-5f, 1);
    k_outer_pool_fwd<<<(B * C * 32 + 255) / 256, 256, 0, s_>>>(no_, pool_, B * C, 49);
    k_outer_output_fwd<<<(B + 255) / 256, 256, 0, s_>>>(pool_, p + off_ow_, p + off_ob_, labels, probs_, labels ? score_parts_ : nullptr, B, C);
    N4J_CUDA(cudaGetLastError());
}

inline void OuterNet::output(const float* params, const float* images, float* probs, int train) {
    N4J_CUDA(cudaSetDevice(device_));
    N4J_CUDA(cudaMemcpyAsync(p_, params, sizeof(float) * n_, cudaMemcpyDefault, s_));
    N4J_CUDA(cudaMemcpyAsync(img_, images, sizeof(float) * B_ * 28 * 28, cudaMemcpyDefault, s_));
    if (tc_) forward_tc(p_, img_, nullptr, train != 0, nullptr);
    else { ode_->set_nhwc_io(false); forward(p_, img_, nullptr, train != 0, nullptr); }
    N4J_CUDA(cudaMemcpyAsync(probs, probs_, sizeof(float) * B_ * 10, cudaMemcpyDefault, s_));
    N4J_CUDA(cudaStreamSynchronize(s_));
}

// backward of the exact path (NCHW, exact-accumulation CUDA-core kernels): gradient view g_ filled from the activations forward() left behind
inline void OuterNet::backward(n4j_stats* ode_bwd) {
    const int C = C_, B = B_;
    // ---- backward --------------------------------------------------------------------------------------------------------------
    const long long nz = (long long)B * C * 49;
    k_outer_output_bwd<<<1, 256, 0, s_>>>(pool_, p_ + off_ow_, probs_, lab_, score_parts_, g_ + off_ow_, g_ + off_ob_, dpool_, score_, B, C);
    k_outer_pool_bwd<<<grid(nz), 256, 0, s_>>>(dpool_, d1_, nz, 49);
    k_outer_bn_bwd<<<C, 256, 0, s_>>>(z1_, no_, d1_, p_ + off_bno_, sto_, d2_, g_ + off_bno_, B, C, 49, 1, 0.1f);
    N4J_CUDA(cudaGetLastError());
    // the adjoint solve: continues the forward of this handle (lastOutput stays on the device), gradient view slice seeded (zero) and written in place
    ode_->backward(nullptr, d2_, time_, 2, N4J_TIMEGRAD_NONE, g_ + off_ode_, d1_, nullptr, ode_bwd);
    float* da = d1_;              // dL/d(stem output)
    float* t1 = d2_;
    float* t2 = d3_;
    for (int r = 1; r >= 0; --r) {
        Block& b = blk_[r];
        const float* a_in = r == 0 ? a0_ : blk_[0].a_out;
        const int HWi = b.Hin * b.Hin, HWo = b.Hout * b.Hout;
        const long long nin = (long long)B * C * HWi;
        // secondConv (identity): weight gradient, then dL/dn2 into t1
        conv_wgrad(b.n2, da, b.f2, g_ + b.o_c2, nullptr, b.c2, 0);
        conv_dgrad(da, b.f2, p_ + b.o_c2, t1, b.c2, 0);
        // secondNorm: dL/df1 into t2
        k_outer_bn_bwd<<<C, 256, 0, s_>>>(b.f1, b.n2, t1, p_ + b.o_bn2, b.st2, t2, g_ + b.o_bn2, B, C, HWo, 1, 0.1f);
        // firstConv (ReLU) and downSample (identity) both read n1: dL/dn1 = downSample's epsilon + firstConv's epsilon
        conv_wgrad(b.n1, t2, b.f1, g_ + b.o_c1, nullptr, b.c1, 1);
        conv_wgrad(b.n1, da, b.dsy, g_ + b.o_ds, nullptr, b.ds, 0);
        conv_dgrad(t2, b.f1, p_ + b.o_c1, t1, b.c1, 1);             // t1: firstConv's epsilon  [B,C,Hin,Hin]
        conv_dgrad(da, b.dsy, p_ + b.o_ds, t2, b.ds, 0);            // t2: downSample's epsilon
        k_outer_add<<<grid(nin), 256, 0, s_>>>(t2, t1, t1, nin);
        // firstNorm: dL/d(block input) into da's next buffer
        k_outer_bn_bwd<<<C, 256, 0, s_>>>(a_in, b.n1, t1, p_ + b.o_bn1, b.st1, t2, g_ + b.o_bn1, B, C, HWi, 1, 0.1f);
        std::swap(da, t2);
        // keep three distinct buffers: da (new epsilon), t1, t2 (scratch)
    }
    conv_wgrad(img_, da, a0_, g_ + off_in_w_, g_ + off_in_b_, cin_, 0);
}

inline void OuterNet::train_step(float* params, float* updater_state, const float* images, const float* labels, float lr, float mu, float* score,
                                 float* grad_out, n4j_stats* ode_fwd, n4j_stats* ode_bwd) {
    N4J_CUDA(cudaSetDevice(device_));
    const int C = C_, B = B_;
    // N4J_NET_DBG=1: wall-clock phase times of one step on stderr (each phase ends with a stream synchronisation: debugging aid only)
    const bool dbg = getenv("N4J_NET_DBG") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto phase = [&](const char* name) {
        if (!dbg) return;
        cudaStreamSynchronize(s_);
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[node4j_net] %-28s %8.3f ms\n", name, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
    N4J_CUDA(cudaMemcpyAsync(p_, params, sizeof(float) * n_, cudaMemcpyDefault, s_));
    N4J_CUDA(cudaMemcpyAsync(v_, updater_state, sizeof(float) * n_, cudaMemcpyDefault, s_));
    N4J_CUDA(cudaMemcpyAsync(img_, images, sizeof(float) * B * 28 * 28, cudaMemcpyDefault, s_));
    N4J_CUDA(cudaMemcpyAsync(lab_, labels, sizeof(float) * B * 10, cudaMemcpyDefault, s_));
    N4J_CUDA(cudaMemsetAsync(g_, 0, sizeof(float) * n_, s_));          // ZeroGrad (util/listen/training/ZeroGrad.java:14-16): the theta-adjoint is seeded from the view
    phase("copies in");
    if (tc_) {
        forward_tc(p_, img_, lab_, true, ode_fwd);
        phase("forward (stem, block, output)");
        backward_tc(img_, ode_bwd);
        phase("backward");
    } else {
        ode_->set_nhwc_io(false);
        forward(p_, img_, lab_, true, ode_fwd);
        phase("forward (stem, block, output)");
        backward(ode_bwd);
        phase("backward");
    }
    // ---- update ----------------------------------------------------------------------------------------------------------------
    if (grad_out) N4J_CUDA(cudaMemcpyAsync(grad_out, g_, sizeof(float) * n_, cudaMemcpyDefault, s_));
    k_outer_update<<<grid(n_), 256, 0, s_>>>(p_, v_, g_, kind_, n_, (float)B, lr, mu, (float)(1.0 + (double)mu));
    N4J_CUDA(cudaGetLastError());
    N4J_CUDA(cudaMemcpyAsync(params, p_, sizeof(float) * n_, cudaMemcpyDefault, s_));
    N4J_CUDA(cudaMemcpyAsync(updater_state, v_, sizeof(float) * n_, cudaMemcpyDefault, s_));
    if (score) N4J_CUDA(cudaMemcpyAsync(score, score_, sizeof(float), cudaMemcpyDefault, s_));
    N4J_CUDA(cudaStreamSynchronize(s_));
    phase("update + copies out");
}

// ---------------------------------------------------------------------------------------------------------------------------
// fast path
// ---------------------------------------------------------------------------------------------------------------------------
inline void OuterNet::init_tc() {
    const int B = B_;
    auto zeroed = [&](long long n) { float* p = dev_alloc(n); N4J_CUDA(cudaMemset(p, 0, sizeof(float) * n)); return p; };
    acc_pool_bytes_ = sizeof(double) * 5 * ON_BN_ACCS * ACC_STRIDE;
    N4J_CUDA(cudaMalloc(&acc_pool_, acc_pool_bytes_)); allocs_.push_back(acc_pool_);
    int max_tiles = 0;
    for (int r = 0; r < 2; ++r) {
        TcBlock& t = tb_[r];
        const Block& b = blk_[r];
        t.gi = make_conv_geom(B, b.Hin, b.Hin); t.go = make_conv_geom(B, b.Hout, b.Hout);
        t.wi = make_wgrad_tc_geom(B, b.Hin, b.Hin); t.wo = make_wgrad_tc_geom(B, b.Hout, b.Hout);
        t.sub_off = 1 - b.c1.pt;
        const long long nin = (long long)B * 64 * b.Hin * b.Hin, nout = (long long)B * 64 * b.Hout * b.Hout;
        t.n1 = dev_alloc(nin); t.n1img = zeroed(conv_image_floats(t.gi));
        t.dsy = dev_alloc(nout); t.f1 = dev_alloc(nout); t.n2 = dev_alloc(nout); t.n2img = zeroed(conv_image_floats(t.go));
        t.f2 = dev_alloc(nout); t.a_out = dev_alloc(nout);
        t.w1f = dev_alloc(9 * 8192); t.w1b = dev_alloc(9 * 8192); t.w2f = dev_alloc(9 * 8192); t.w2b = dev_alloc(9 * 8192);
        t.t1 = dev_alloc(nout); t.gup = zeroed(nin); t.gupimg = zeroed(conv_image_floats(t.gi)); t.dn1 = dev_alloc(nin);
        t.daimg = zeroed(conv_image_floats(t.go)); t.da_in = dev_alloc(nin);
        t.acc1 = acc_pool_ + (size_t)(2 * r) * ON_BN_ACCS * ACC_STRIDE; t.acc2 = acc_pool_ + (size_t)(2 * r + 1) * ON_BN_ACCS * ACC_STRIDE;
        conv_smem_max_ = std::max(conv_smem_max_, std::max(conv_tc_smem_bytes(t.gi), conv_tc_smem_bytes(t.go)));
        wgrad_smem_max_ = std::max(wgrad_smem_max_, std::max(conv_wgrad_tc_smem_bytes(t.wi), conv_wgrad_tc_smem_bytes(t.wo)));
        max_tiles = std::max(max_tiles, std::max(t.wi.tiles, t.wo.tiles));
    }
    if (conv_smem_max_ > 227 * 1024 || wgrad_smem_max_ > 227 * 1024) throw Error{N4J_ERR_UNSUPPORTED, "stem maps too wide for the tensor-core convolutions"};
    acco_ = acc_pool_ + (size_t)4 * ON_BN_ACCS * ACC_STRIDE;
    const long long n0 = (long long)B * 64 * 26 * 26, nz = (long long)B * 64 * 49;
    a0n_ = dev_alloc(n0); z1n_ = dev_alloc(nz); non_ = dev_alloc(nz); dA_ = dev_alloc(nz); dB_ = dev_alloc(nz); dC_ = dev_alloc(nz);
    wpart_tc_ = dev_alloc((long long)max_tiles * 36864); wmid_ = dev_alloc((long long)WMID * 36864);
    const long long ds_blocks = 148, in_blocks = 148 * 2;
    N4J_CUDA(cudaMalloc(&dspart_, sizeof(double) * ds_blocks * 4096)); allocs_.push_back(dspart_);
    N4J_CUDA(cudaMalloc(&inpart_, sizeof(double) * in_blocks * 640)); allocs_.push_back(inpart_);
    N4J_CUDA(cudaStreamCreateWithFlags(&s2_, cudaStreamNonBlocking));
    for (auto& e : ev_fork_) N4J_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : ev_ds_) N4J_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    N4J_CUDA(cudaEventCreateWithFlags(&ev_join_, cudaEventDisableTiming));
}
inline void OuterNet::fork_side() {
    cudaEvent_t e = ev_fork_[fork_n_++ & 7];
    N4J_CUDA(cudaEventRecord(e, s_));
    N4J_CUDA(cudaStreamWaitEvent(s2_, e, 0));
}
inline void OuterNet::join_side() {
    N4J_CUDA(cudaEventRecord(ev_join_, s2_));
    N4J_CUDA(cudaStreamWaitEvent(s_, ev_join_, 0));
}

// The dynamic shared-memory limits are per-function attributes shared with every Engine of the process (whose own maps are smaller):
// raised before each step's launches.
inline void OuterNet::set_tc_attributes() {
    N4J_CUDA(cudaFuncSetAttribute(k_conv3x3_tc<3, 0, 0, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)conv_smem_max_));
    N4J_CUDA(cudaFuncSetAttribute(k_conv3x3_tc<1, 0, 0, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)conv_smem_max_));
    N4J_CUDA(cudaFuncSetAttribute(k_conv3x3_tc<3, 0, 0, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)conv_smem_max_));
    N4J_CUDA(cudaFuncSetAttribute(k_conv3x3_tc<1, 0, 0, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)conv_smem_max_));
    N4J_CUDA(cudaFuncSetAttribute(k_conv_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wgrad_smem_max_));
}

inline void OuterNet::conv_tc(const float* img, const float* wimg, float* out, const ConvGeom& g, int act, bool sub, int sub_off, int Ho) {
    ConvTcArgs a{};
    a.img = img; a.wimg = wimg; a.out = fixed_ref(out); a.g = g; a.act = act;
    a.sub_s = 2; a.sub_off = sub_off; a.sub_H = Ho; a.sub_W = Ho;
    const size_t smem = conv_tc_smem_bytes(g);
    const dim3 grid(g.tiles), block(CONV_TC_THREADS);
    if (sub) {
        if (passes_ == 1) k_conv3x3_tc<1, 0, 0, 0, 1><<<grid, block, smem, s_>>>(a);
        else k_conv3x3_tc<3, 0, 0, 0, 1><<<grid, block, smem, s_>>>(a);
    } else {
        if (passes_ == 1) k_conv3x3_tc<1, 0, 0, 0, 0><<<grid, block, smem, s_>>>(a);
        else k_conv3x3_tc<3, 0, 0, 0, 0><<<grid, block, smem, s_>>>(a);
    }
}

// dw_flat: DL4J's [co][ci][3][3] slot of the gradient view
inline void OuterNet::wgrad_tc(cudaStream_t st, const float* x, const float* g, const WgradTcGeom& wg, float* dw_flat) {
    k_conv_wgrad_tc<<<dim3(wg.tiles), WGRAD_TC_THREADS, conv_wgrad_tc_smem_bytes(wg), st>>>(fixed_ref(const_cast<float*>(x)), fixed_ref(const_cast<float*>(g)), wg,
                                                                                             wpart_tc_, (long long*)nullptr);
    const int per = (wg.tiles + WMID - 1) / WMID, nq = (wg.tiles + per - 1) / per;
    k_on_wgrad_sum1<<<dim3(36, nq), 256, 0, st>>>(wpart_tc_, wmid_, wg.tiles, per);
    k_on_wgrad_sum2<<<144, 256, 0, st>>>(wmid_, dw_flat, nq);
}

inline void OuterNet::forward_tc(const float* p, const float* images, const float* labels, bool train, n4j_stats* ode_fwd) {
    const int B = B_, C = 64;
    set_tc_attributes();
    N4J_CUDA(cudaMemsetAsync(acc_pool_, 0, acc_pool_bytes_, s_));
    const long long r0 = (long long)B * 26 * 26;
    k_on_in_conv_fwd<<<grid_rows(r0), 256, 0, s_>>>(images, p + off_in_w_, p + off_in_b_, a0n_, B, 28, 28, train ? tb_[0].acc1 : (double*)nullptr);
    {
        OnWeightJobs jobs{};
        for (int r = 0; r < 2; ++r) {
            const float* w1 = p + blk_[r].o_c1;
            const float* w2 = p + blk_[r].o_c2;
            jobs.w[4 * r] = w1; jobs.img[4 * r] = tb_[r].w1f; jobs.mode[4 * r] = 0;
            jobs.w[4 * r + 1] = w1; jobs.img[4 * r + 1] = tb_[r].w1b; jobs.mode[4 * r + 1] = 1;
            jobs.w[4 * r + 2] = w2; jobs.img[4 * r + 2] = tb_[r].w2f; jobs.mode[4 * r + 2] = 0;
            jobs.w[4 * r + 3] = w2; jobs.img[4 * r + 3] = tb_[r].w2b; jobs.mode[4 * r + 3] = 1;
        }
        k_on_weights_to_tc<<<dim3(36, 8), 256, 0, s_>>>(jobs);
    }
    const float* a = a0n_;
    for (int r = 0; r < 2; ++r) {
        TcBlock& t = tb_[r];
        const Block& b = blk_[r];
        const long long ri = (long long)B * b.Hin * b.Hin, ro = (long long)B * b.Hout * b.Hout;
        if (r > 0 && train) k_on_bn_stats<<<grid_rows(ri), 256, 0, s_>>>(a, ri, t.acc1);        // (block 0: the input convolution accumulated them)
        k_on_bn_apply<<<grid(ri * 16), 256, 0, s_>>>(a, t.acc1, ri, 1e-5f, p + b.o_bn1, p + b.o_bn1 + C, train ? nullptr : p + b.o_bn1 + 2 * C, t.n1, t.n1img, t.gi);
        fork_side();                                                                                     // downSample beside firstConvStem
        k_on_ds_fwd<<<(unsigned)((ro + ON_DS_PIX - 1) / ON_DS_PIX), 256, 0, s2_>>>(t.n1, p + b.o_ds, t.dsy, B, b.Hin, b.Hout);
        N4J_CUDA(cudaEventRecord(ev_ds_[r], s2_));
        conv_tc(t.n1img, t.w1f, t.f1, t.gi, N4J_ACT_RELU, true, t.sub_off, b.Hout);                      // firstConvStem: stride 2, ReLU
        if (train) k_on_bn_stats<<<grid_rows(ro), 256, 0, s_>>>(t.f1, ro, t.acc2);
        k_on_bn_apply<<<grid(ro * 16), 256, 0, s_>>>(t.f1, t.acc2, ro, 1e-5f, p + b.o_bn2, p + b.o_bn2 + C, train ? nullptr : p + b.o_bn2 + 2 * C, t.n2, t.n2img, t.go);
        conv_tc(t.n2img, t.w2f, t.f2, t.go, N4J_ACT_IDENTITY, false, 0, 0);                              // secondConvStem
        N4J_CUDA(cudaStreamWaitEvent(s_, ev_ds_[r], 0));
        k_outer_add<<<grid(ro * 64), 256, 0, s_>>>(t.dsy, t.f2, t.a_out, ro * 64);                       // stemAdd_r
        a = t.a_out;
    }
    N4J_CUDA(cudaGetLastError());
    ode_->set_nhwc_io(true);
    ode_->bind_params(p + off_ode_, n_ode_);
    ode_->forward(a, time_, 2, 0, z1n_, ode_fwd);
    const long long rz = (long long)B * 49;
    if (train) k_on_bn_stats<<<grid_rows(rz), 256, 0, s_>>>(z1n_, rz, acco_);
    k_on_bn_apply<<<grid(rz * 16), 256, 0, s_>>>(z1n_, acco_, rz, 1e-5f, p + off_bno_, p + off_bno_ + C, train ? nullptr : p + off_bno_ + 2 * C, non_, (float*)nullptr, ConvGeom{});
    k_on_pool_fwd<<<(B * 64 + 255) / 256, 256, 0, s_>>>(non_, pool_, B, 49);
    k_on_output_fwd<<<(B * 32 + 255) / 256, 256, 0, s_>>>(pool_, p + off_ow_, p + off_ob_, labels, probs_, labels ? score_parts_ : nullptr, B);
    N4J_CUDA(cudaGetLastError());
}

inline void OuterNet::backward_tc(const float* images, n4j_stats* ode_bwd) {
    const int B = B_, C = 64;
    const long long rz = (long long)B * 49;
    set_tc_attributes();      // (an Engine constructed since forward_tc would have lowered the shared limits again)
    k_on_output_bwd<<<(704 + B * 64 + 255) / 256, 256, 0, s_>>>(pool_, p_ + off_ow_, probs_, lab_, score_parts_, g_ + off_ow_, g_ + off_ob_, dpool_, score_, B);
    k_on_pool_bwd<<<grid(rz * 64), 256, 0, s_>>>(dpool_, dA_, rz * 64, 49);
    k_on_bn_bwd_stats<<<grid_rows(rz), 256, 0, s_>>>(z1n_, non_, dA_, acco_, rz, 1e-5f);
    OnBnBwdOut oo{dB_, nullptr, ConvGeom{}, 7, 7, 1, 0, 0};
    k_on_bn_bwd_dx<<<grid(rz * 16), 256, 0, s_>>>(z1n_, non_, dA_, acco_, rz, 1e-5f, p_ + off_bno_, g_ + off_bno_, 0.1f, oo);
    N4J_CUDA(cudaGetLastError());
    ode_->backward(nullptr, dB_, time_, 2, N4J_TIMEGRAD_NONE, g_ + off_ode_, dC_, nullptr, ode_bwd);
    const float* da = dC_;            // epsilon at the output of block r
    k_compact_to_image<<<grid(rz * 16), 256, 0, s_>>>(fixed_ref(dC_), tb_[1].go, tb_[1].daimg);
    for (int r = 1; r >= 0; --r) {
        TcBlock& t = tb_[r];
        const Block& b = blk_[r];
        const float* a_in = r == 0 ? a0n_ : tb_[0].a_out;
        const long long ri = (long long)B * b.Hin * b.Hin, ro = (long long)B * b.Hout * b.Hout;
        // secondConvStem: weight gradient (side branch, with downSample's: both only need the block's output epsilon), epsilon of secondNorm's output
        fork_side();
        wgrad_tc(s2_, t.n2, da, t.wo, g_ + b.o_c2);
        const int dsb = (int)std::min<long long>((ro + ON_DS_WG_PIX - 1) / ON_DS_WG_PIX, 148);
        k_on_ds_wgrad<<<dsb, 256, 0, s2_>>>(t.n1, da, dspart_, B, b.Hin, b.Hout);
        k_on_sum_parts<<<128, 256, 0, s2_>>>(dspart_, g_ + b.o_ds, 4096, dsb, 4096);
        conv_tc(t.daimg, t.w2b, t.t1, t.go, N4J_ACT_IDENTITY, false, 0, 0);
        // secondNorm backward; its dx, masked by firstConvStem's ReLU, scattered to the sampled pixels of the full-resolution map
        k_on_bn_bwd_stats<<<grid_rows(ro), 256, 0, s_>>>(t.f1, t.n2, t.t1, t.acc2, ro, 1e-5f);
        OnBnBwdOut o2{t.gup, t.gupimg, t.gi, b.Hout, b.Hin, 2, t.sub_off, 1};
        k_on_bn_bwd_dx<<<grid(ro * 16), 256, 0, s_>>>(t.f1, t.n2, t.t1, t.acc2, ro, 1e-5f, p_ + b.o_bn2, g_ + b.o_bn2, 0.1f, o2);
        // firstConvStem: weight gradient on the side branch, epsilon of n1 on the chain (downSample's epsilon is added onto it)
        fork_side();
        wgrad_tc(s2_, t.n1, t.gup, t.wi, g_ + b.o_c1);
        conv_tc(t.gupimg, t.w1b, t.dn1, t.gi, N4J_ACT_IDENTITY, false, 0, 0);
        k_on_ds_dgrad_add<<<(unsigned)((ro + ON_DS_PIX - 1) / ON_DS_PIX), 256, 0, s_>>>(da, p_ + b.o_ds, t.dn1, B, b.Hin, b.Hout);
        // firstNorm backward -> epsilon at the block's input (and, for block 1, its operand image for block 0's secondConvStem dgrad)
        k_on_bn_bwd_stats<<<grid_rows(ri), 256, 0, s_>>>(a_in, t.n1, t.dn1, t.acc1, ri, 1e-5f);
        OnBnBwdOut o1{t.da_in, r == 1 ? tb_[0].daimg : nullptr, r == 1 ? tb_[0].go : ConvGeom{}, b.Hin, b.Hin, 1, 0, 0};
        k_on_bn_bwd_dx<<<grid(ri * 16), 256, 0, s_>>>(a_in, t.n1, t.dn1, t.acc1, ri, 1e-5f, p_ + b.o_bn1, g_ + b.o_bn1, 0.1f, o1);
        da = t.da_in;
    }
    const long long r0 = (long long)B * 26 * 26;
    const int inb = (int)std::min<long long>((r0 + 15) / 16, 148 * 2);
    k_on_in_conv_wgrad<<<inb, 256, 0, s_>>>(images, da, inpart_, B, 28, 28);
    // inpart: [co][tap] then [co] per block; the gradient view stores the bias first (bias-first convolution parameters)
    k_on_sum_parts<<<18, 256, 0, s_>>>(inpart_, g_ + off_in_w_, 576, inb, 640);
    k_on_sum_parts<<<2, 256, 0, s_>>>(inpart_ + 576, g_ + off_in_b_, 64, inb, 640);
    join_side();
    N4J_CUDA(cudaGetLastError());
}

}  // namespace n4j
