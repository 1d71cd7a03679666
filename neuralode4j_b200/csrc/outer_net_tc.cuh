// outer_net_tc.cuh — kernels of the FAST path of the network around the ODE block, 64 kernels per layer.
//
// Same layers as outer_net.cuh (examples/mnist/ResBlockStem.java:25-48, Output.java:27-35), other data path: activations are NHWC
// ([B*H*W][64], the ODE engine's internal layout, so the block takes and returns them without a permutation), and every 64 -> 64
// 3x3 convolution of the stem — forward, input gradient and weight gradient — runs on the tcgen05 kernels of the hot path
// (conv_tc.cuh, conv_wgrad_tc.cuh; 3xTF32, fp32-faithful):
//   * stride 1 (secondConvStem): as in the ODE block, from a prepared operand image;
//   * stride 2 (firstConvStem, ConvolutionMode.Same): the launch computes the stride-1 "same" convolution of the full-resolution
//     image and its epilogue stores only the sampled pixels (k_conv3x3_tc<.., SUB = 1>); backward, the gradient is scattered into a
//     zero full-resolution tensor (compact + operand image), which turns dgrad and wgrad into the stride-1 kernels as well.
// What stays on the CUDA cores is small: the 1 -> 64 input convolution (9 products per output), the 1x1 stride-2 downSample
// (64 products), BatchNorm, pooling, the output layer.  Their short contractions (9 / 64 products, a slice of pixels) use fp32 FMA like the
// tensor-core accumulators beside them — B200's FP64 pipe is an order of magnitude slower — while every long reduction (BatchNorm
// statistics, sums over blocks of pixels, the output layer) accumulates in double; elementwise expressions keep the oracle's order
// (oracle/outer_oracle.py).  BatchNorm statistics are raw double sums that every
// consumer finalises itself (the engine's deferred scheme, layer_kernels.cuh), so no kernel waits for a last block.
#pragma once

#include "conv_tc.cuh"
#include "conv_wgrad_tc.cuh"

namespace n4j {

// accumulators of one BatchNorm layer: [0,64) sum x, [64,128) sum x^2, [128,192) sum dy, [192,256) sum dy*xhat — ACC_STRIDE doubles apart
constexpr int ON_BN_ACCS = 256;

__device__ __forceinline__ void on_bn_final(const double* __restrict__ acc, long long M, float eps, int c, float& mean, float& inv, float& var_f) {
    const double m = acc[c * ACC_STRIDE] / (double)M;
    double var = acc[(64 + c) * ACC_STRIDE] / (double)M - m * m;
    if (var < 0) var = 0;
    mean = (float)m;
    inv = (float)(1.0 / sqrt(var + (double)eps));
    var_f = (float)var;
}

// inputConv: y[(b,oh,ow)][co] = (float)sum_{kh,kw} x[b,oh+kh,ow+kw] * w[co,kh,kw] + bias[co]      (1 -> 64, 3x3, Truncate)
// thread = (output pixel, 4-channel chunk); also accumulates the statistics of the BatchNorm behind it (acc may be null)
__global__ void __launch_bounds__(256) k_on_in_conv_fwd(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                         float* __restrict__ y, int B, int H, int W, double* __restrict__ acc) {
    __shared__ float s_w[9][64];
    __shared__ float s_b[64];
    __shared__ double s_red[2][16][64];
    for (int i = threadIdx.x; i < 576; i += 256) s_w[i % 9][i / 9] = w[i];
    if (threadIdx.x < 64) s_b[threadIdx.x] = bias[threadIdx.x];
    __syncthreads();
    const int Ho = H - 2, Wo = W - 2;
    const long long n = (long long)B * Ho * Wo;
    const int chunk = threadIdx.x & 15, c0 = chunk * 4;
    double s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
    for (long long p = (long long)blockIdx.x * 16 + (threadIdx.x >> 4); p < n; p += (long long)gridDim.x * 16) {
        const int ow = (int)(p % Wo), oh = (int)((p / Wo) % Ho), b = (int)(p / ((long long)Wo * Ho));
        const float* xb = x + ((long long)b * H + oh) * W + ow;
        float a[4] = {0.f, 0.f, 0.f, 0.f};       // nine products: fp32 FMA (the FP64 pipe of this part is an order of magnitude slower)
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const float xv = xb[(t / 3) * W + (t % 3)];
#pragma unroll
            for (int k = 0; k < 4; ++k) a[k] = fmaf(xv, s_w[t][c0 + k], a[k]);
        }
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            o[k] = __fadd_rn(a[k], s_b[c0 + k]);
            s1[k] += (double)o[k]; s2[k] += (double)o[k] * (double)o[k];
        }
        reinterpret_cast<float4*>(y)[p * 16 + chunk] = make_float4(o[0], o[1], o[2], o[3]);
    }
    if (acc) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { s_red[0][threadIdx.x >> 4][c0 + k] = s1[k]; s_red[1][threadIdx.x >> 4][c0 + k] = s2[k]; }
        __syncthreads();
        if (threadIdx.x < 128) {
            const int q = threadIdx.x >> 6, c = threadIdx.x & 63;
            double t = 0;
            for (int r = 0; r < 16; ++r) t += s_red[q][r][c];
            atomicAdd(acc + (q * 64 + c) * ACC_STRIDE, t);
        }
    }
}

// column sums of an NHWC tensor [M][64]: acc[c] += sum x, acc[64 + c] += sum x^2
__global__ void __launch_bounds__(256) k_on_bn_stats(const float* __restrict__ x, long long M, double* __restrict__ acc) {
    __shared__ double s_red[2][16][64];
    const int chunk = threadIdx.x & 15, c0 = chunk * 4;
    double s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
    for (long long r = (long long)blockIdx.x * 16 + (threadIdx.x >> 4); r < M; r += (long long)gridDim.x * 16) {
        const float4 v = reinterpret_cast<const float4*>(x)[r * 16 + chunk];
        const float a[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) { s1[k] += (double)a[k]; s2[k] += (double)a[k] * (double)a[k]; }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) { s_red[0][threadIdx.x >> 4][c0 + k] = s1[k]; s_red[1][threadIdx.x >> 4][c0 + k] = s2[k]; }
    __syncthreads();
    if (threadIdx.x < 128) {
        const int q = threadIdx.x >> 6, c = threadIdx.x & 63;
        double t = 0;
        for (int r = 0; r < 16; ++r) t += s_red[q][r][c];
        atomicAdd(acc + (q * 64 + c) * ACC_STRIDE, t);
    }
}

// y = relu(gamma * ((x - mean) * invstd) + beta): compact NHWC and (img != null) the zero-padded hi/lo operand image of the convolution behind.
// run == null: batch statistics from the raw sums in acc (training); else the layer's running mean / var (inference, ComputationGraph.output)
__global__ void __launch_bounds__(256) k_on_bn_apply(const float* __restrict__ x, const double* __restrict__ acc, long long M, float eps,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ run,
                                                      float* __restrict__ y, float* __restrict__ img, ConvGeom g) {
    __shared__ float s_m[64], s_i[64], s_g[64], s_b[64];
    if (threadIdx.x < 64) {
        if (run) {
            s_m[threadIdx.x] = run[threadIdx.x];
            s_i[threadIdx.x] = (float)(1.0 / sqrt((double)run[64 + threadIdx.x] + (double)eps));
        } else {
            float var;
            on_bn_final(acc, M, eps, threadIdx.x, s_m[threadIdx.x], s_i[threadIdx.x], var);
        }
        s_g[threadIdx.x] = gamma[threadIdx.x]; s_b[threadIdx.x] = beta[threadIdx.x];
    }
    __syncthreads();
    const long long n4 = M * 16;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const int chunk = (int)(i & 15), c = chunk * 4;
        const float4 v = reinterpret_cast<const float4*>(x)[i];
        const float a[4] = {v.x, v.y, v.z, v.w};
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float t = __fadd_rn(__fmul_rn(s_g[c + k], __fmul_rn(__fsub_rn(a[k], s_m[c + k]), s_i[c + k])), s_b[c + k]);
            o[k] = t > 0.f ? t : 0.f;
        }
        const float4 ov = make_float4(o[0], o[1], o[2], o[3]);
        reinterpret_cast<float4*>(y)[i] = ov;
        if (img) store_split4(img, g, chunk, padded_row(g, i >> 4), ov);
    }
}

// downSample: 1x1 convolution, stride 2, no bias: y[(b,oh,ow)][co] = (float)sum_ci x[(b,2oh,2ow)][ci] * w[co][ci]
// block = 64 output pixels; thread = 4 pixels x 4 output channels (one 128-bit weight load and four broadcast activation loads per 16 products)
constexpr int ON_DS_PIX = 64;
__global__ void __launch_bounds__(256) k_on_ds_fwd(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y, int B, int Hi, int Ho) {
    __shared__ __align__(16) float s_w[64][68];      // [ci][co]
    __shared__ __align__(16) float s_x[ON_DS_PIX][64];
    for (int i = threadIdx.x; i < 4096; i += 256) s_w[i & 63][i >> 6] = w[i];      // w[co][ci]
    const long long P = (long long)B * Ho * Ho, p0 = (long long)blockIdx.x * ON_DS_PIX;
    for (int i = threadIdx.x; i < ON_DS_PIX * 16; i += 256) {
        const long long p = p0 + (i >> 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p < P) {
            const int ow = (int)(p % Ho), oh = (int)((p / Ho) % Ho), b = (int)(p / ((long long)Ho * Ho));
            v = reinterpret_cast<const float4*>(x)[(((long long)b * Hi + 2 * oh) * Hi + 2 * ow) * 16 + (i & 15)];
        }
        reinterpret_cast<float4*>(&s_x[i >> 4][0])[i & 15] = v;
    }
    __syncthreads();
    const int c0 = (threadIdx.x & 15) * 4, r0 = (threadIdx.x >> 4) * 4;
    float a[4][4];      // 64 products per output: fp32 FMA, like the TMEM accumulators of the 3x3 convolutions beside it
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) a[i][k] = 0.f;
    for (int ci = 0; ci < 64; ++ci) {
        const float4 wv = *reinterpret_cast<const float4*>(&s_w[ci][c0]);
        const float ws[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float xv = s_x[r0 + i][ci];
#pragma unroll
            for (int k = 0; k < 4; ++k) a[i][k] = fmaf(xv, ws[k], a[i][k]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (p0 + r0 + i < P)
            reinterpret_cast<float4*>(y)[(p0 + r0 + i) * 16 + (threadIdx.x & 15)] = make_float4(a[i][0], a[i][1], a[i][2], a[i][3]);
}

// dn1[(b,2oh,2ow)][ci] += (float)sum_co g[(b,oh,ow)][co] * w[co][ci]       (downSample's epsilon added onto firstConv's, in place)
__global__ void __launch_bounds__(256) k_on_ds_dgrad_add(const float* __restrict__ g, const float* __restrict__ w, float* __restrict__ dx, int B, int Hi, int Ho) {
    __shared__ __align__(16) float s_w[64][64];      // [co][ci]
    __shared__ __align__(16) float s_g[ON_DS_PIX][64];
    for (int i = threadIdx.x; i < 1024; i += 256) reinterpret_cast<float4*>(&s_w[0][0])[i] = reinterpret_cast<const float4*>(w)[i];
    const long long P = (long long)B * Ho * Ho, p0 = (long long)blockIdx.x * ON_DS_PIX;
    for (int i = threadIdx.x; i < ON_DS_PIX * 16; i += 256) {
        const long long p = p0 + (i >> 4);
        reinterpret_cast<float4*>(&s_g[i >> 4][0])[i & 15] = p < P ? reinterpret_cast<const float4*>(g)[p * 16 + (i & 15)] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    const int c0 = (threadIdx.x & 15) * 4, r0 = (threadIdx.x >> 4) * 4;
    float a[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) a[i][k] = 0.f;
    for (int co = 0; co < 64; ++co) {
        const float4 wv = *reinterpret_cast<const float4*>(&s_w[co][c0]);
        const float ws[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float gv = s_g[r0 + i][co];
#pragma unroll
            for (int k = 0; k < 4; ++k) a[i][k] = fmaf(gv, ws[k], a[i][k]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long p = p0 + r0 + i;
        if (p >= P) break;
        const int ow = (int)(p % Ho), oh = (int)((p / Ho) % Ho), b = (int)(p / ((long long)Ho * Ho));
        float4* d = reinterpret_cast<float4*>(dx) + (((long long)b * Hi + 2 * oh) * Hi + 2 * ow) * 16 + (threadIdx.x & 15);
        const float4 u = *d;
        *d = make_float4(__fadd_rn(a[i][0], u.x), __fadd_rn(a[i][1], u.y), __fadd_rn(a[i][2], u.z), __fadd_rn(a[i][3], u.w));
    }
}

// downSample weight gradient: dw[co][ci] = sum_p x[(b,2oh,2ow)][ci] * g[p][co]; a block walks 64-pixel slices (grid stride), thread = 4 ci x 4 co,
// one double partial [co][ci] per block
constexpr int ON_DS_WG_PIX = 64;
__global__ void __launch_bounds__(256) k_on_ds_wgrad(const float* __restrict__ x, const float* __restrict__ g, double* __restrict__ part, int B, int Hi, int Ho) {
    __shared__ __align__(16) float s_x[ON_DS_WG_PIX][64], s_g[ON_DS_WG_PIX][64];
    const long long P = (long long)B * Ho * Ho;
    const int ci0 = (threadIdx.x >> 4) * 4, co0 = (threadIdx.x & 15) * 4;
    double a[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) a[i][j] = 0.0;
    for (long long p0 = (long long)blockIdx.x * ON_DS_WG_PIX; p0 < P; p0 += (long long)gridDim.x * ON_DS_WG_PIX) {
        __syncthreads();
        for (int i = threadIdx.x; i < ON_DS_WG_PIX * 16; i += 256) {
            const long long p = p0 + (i >> 4);
            float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), gv = xv;
            if (p < P) {
                const int ow = (int)(p % Ho), oh = (int)((p / Ho) % Ho), b = (int)(p / ((long long)Ho * Ho));
                xv = reinterpret_cast<const float4*>(x)[(((long long)b * Hi + 2 * oh) * Hi + 2 * ow) * 16 + (i & 15)];
                gv = reinterpret_cast<const float4*>(g)[p * 16 + (i & 15)];
            }
            reinterpret_cast<float4*>(&s_x[i >> 4][0])[i & 15] = xv;
            reinterpret_cast<float4*>(&s_g[i >> 4][0])[i & 15] = gv;
        }
        __syncthreads();
        float f[4][4];      // the 64 pixels of a slice in fp32 FMA, slices and blocks in double
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) f[i][j] = 0.f;
        for (int r = 0; r < ON_DS_WG_PIX; ++r) {
            const float4 xv = *reinterpret_cast<const float4*>(&s_x[r][ci0]);
            const float4 gv = *reinterpret_cast<const float4*>(&s_g[r][co0]);
            const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, gs[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) f[i][j] = fmaf(xs[i], gs[j], f[i][j]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) a[i][j] += (double)f[i][j];
    }
    double* o = part + (long long)blockIdx.x * 4096;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) o[(co0 + j) * 64 + ci0 + i] = a[i][j];      // [co][ci], DL4J's order
}
// dst[i] = (float) sum_q part[q * stride + i]: eight lanes per output stride over the partials, combined in lane order (deterministic)
__global__ void __launch_bounds__(256) k_on_sum_parts(const double* __restrict__ part, float* __restrict__ dst, int n, int nparts, int stride) {
    __shared__ double s_p[8][32];
    const int i = blockIdx.x * 32 + (threadIdx.x & 31), ql = threadIdx.x >> 5;
    double s = 0.0;
    if (i < n)
        for (int q = ql; q < nparts; q += 8) s += part[(long long)q * stride + i];
    s_p[ql][threadIdx.x & 31] = s;
    __syncthreads();
    if (ql == 0 && i < n) {
        double t = s_p[0][threadIdx.x];
#pragma unroll
        for (int k = 1; k < 8; ++k) t += s_p[k][threadIdx.x];
        dst[i] = (float)t;
    }
}

// inputConv weight / bias gradient: dw[co][t] = sum x[b,oh+kh,ow+kw] * g[(b,oh,ow)][co], db[co] = sum g.  A block walks pixels with a grid
// stride (thread = 4-channel chunk x 16 pixel lanes), one partial of 640 doubles per block: dw [co][tap], then db [co]
__global__ void __launch_bounds__(256) k_on_in_conv_wgrad(const float* __restrict__ x, const float* __restrict__ g, double* __restrict__ part, int B, int H, int W) {
    __shared__ double s_red[16][64];
    const int Ho = H - 2, Wo = W - 2;
    const long long n = (long long)B * Ho * Wo;
    const int chunk = threadIdx.x & 15, c0 = chunk * 4, sub = threadIdx.x >> 4;
    float a[10][4];      // a thread's own pixels (a few dozen) in fp32 FMA; the 16 pixel lanes and the blocks are summed in double
#pragma unroll
    for (int t = 0; t < 10; ++t)
#pragma unroll
        for (int k = 0; k < 4; ++k) a[t][k] = 0.f;
    for (long long p = (long long)blockIdx.x * 16 + sub; p < n; p += (long long)gridDim.x * 16) {
        const int ow = (int)(p % Wo), oh = (int)((p / Wo) % Ho), b = (int)(p / ((long long)Wo * Ho));
        const float4 gv = reinterpret_cast<const float4*>(g)[p * 16 + chunk];
        const float gs[4] = {gv.x, gv.y, gv.z, gv.w};
        const float* xb = x + ((long long)b * H + oh) * W + ow;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const float xv = xb[(t / 3) * W + (t % 3)];
#pragma unroll
            for (int k = 0; k < 4; ++k) a[t][k] = fmaf(xv, gs[k], a[t][k]);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) a[9][k] += gs[k];
    }
    double* o = part + (long long)blockIdx.x * 640;
#pragma unroll
    for (int t = 0; t < 10; ++t) {
#pragma unroll
        for (int k = 0; k < 4; ++k) s_red[sub][c0 + k] = (double)a[t][k];
        __syncthreads();
        if (threadIdx.x < 64) {
            double sm = 0.0;
            for (int r = 0; r < 16; ++r) sm += s_red[r][threadIdx.x];
            o[t < 9 ? threadIdx.x * 9 + t : 576 + threadIdx.x] = sm;       // dw [co][tap] then db [co]
        }
        __syncthreads();
    }
}

// global average pooling, NHWC: pool[b][c] = (float)(sum_p x[(b,p)][c] / HW)
__global__ void __launch_bounds__(256) k_on_pool_fwd(const float* __restrict__ x, float* __restrict__ p, int B, int HW) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * 64) return;
    const int b = i >> 6, c = i & 63;
    double s = 0.0;
    for (int q = 0; q < HW; ++q) s += (double)x[((long long)b * HW + q) * 64 + c];
    p[i] = (float)(s / (double)HW);
}
__global__ void __launch_bounds__(256) k_on_pool_bwd(const float* __restrict__ dp, float* __restrict__ dx, long long n, int HW) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dx[i] = __fdiv_rn(dp[(i / (64LL * HW)) * 64 + (i & 63)], (float)HW);
}

// OutputLayer forward: one warp per example.  logits = pool W + b (W [C,10] f-order), softmax, LossMCXENT term of the example
__global__ void __launch_bounds__(256) k_on_output_fwd(const float* __restrict__ pool, const float* __restrict__ W, const float* __restrict__ bias,
                                                        const float* __restrict__ labels, float* __restrict__ probs, double* __restrict__ score_parts, int B) {
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (b >= B) return;
    const double p0 = (double)pool[b * 64 + lane], p1 = (double)pool[b * 64 + 32 + lane];
    float lg[10];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 10; ++j) {
        const double acc = warp_sum(p0 * (double)W[lane + 64 * j] + p1 * (double)W[32 + lane + 64 * j]);
        lg[j] = __fadd_rn((float)acc, bias[j]);
        mx = fmaxf(mx, lg[j]);
    }
    if (lane != 0) return;
    double se = 0.0;
#pragma unroll
    for (int j = 0; j < 10; ++j) { lg[j] = expf(__fsub_rn(lg[j], mx)); se += (double)lg[j]; }
    const float den = (float)se;
    double sc = 0.0;
#pragma unroll
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
// OutputLayer backward, one thread per result: dW[c + 64 j], db[j], dpool[b][c], and the score
__global__ void __launch_bounds__(256) k_on_output_bwd(const float* __restrict__ pool, const float* __restrict__ W, const float* __restrict__ probs,
                                                        const float* __restrict__ labels, const double* __restrict__ score_parts, float* __restrict__ dW,
                                                        float* __restrict__ db, float* __restrict__ dpool, float* __restrict__ score, int B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 640) {
        const int c = i & 63, j = i >> 6;
        double acc = 0.0;
        for (int b = 0; b < B; ++b) acc += (double)pool[b * 64 + c] * (double)__fsub_rn(probs[b * 10 + j], labels[b * 10 + j]);
        dW[i] = (float)acc;
    } else if (i < 650) {
        const int j = i - 640;
        double acc = 0.0;
        for (int b = 0; b < B; ++b) acc += (double)__fsub_rn(probs[b * 10 + j], labels[b * 10 + j]);
        db[j] = (float)acc;
    } else if (i == 650) {
        double sm = 0.0;
        for (int b = 0; b < B; ++b) sm += score_parts[b];
        *score = (float)(sm / (double)B);
    } else if (i >= 704 && i < 704 + B * 64) {
        const int k = i - 704, b = k >> 6, c = k & 63;
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < 10; ++j) acc += (double)__fsub_rn(probs[b * 10 + j], labels[b * 10 + j]) * (double)W[c + 64 * j];
        dpool[k] = (float)acc;
    }
}

// the hi/lo weight images of several convolutions in one launch: blockIdx.y selects (weights, image, mode)
struct OnWeightJobs { const float* w[8]; float* img[8]; int mode[8]; };
__global__ void __launch_bounds__(256) k_on_weights_to_tc(OnWeightJobs jobs) {
    const float* __restrict__ w = jobs.w[blockIdx.y];
    float* __restrict__ wimg = jobs.img[blockIdx.y];
    const int mode = jobs.mode[blockIdx.y];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 64 * 64 * 9; i += gridDim.x * blockDim.x) {
        const int tap = i % 9; const int r = i / 9; const int ci = r % 64; const int co = r / 64;
        float hi, lo;
        tc::split_tf32(w[i], hi, lo);
        const int t2 = mode == 0 ? tap : 8 - tap;
        const int n = mode == 0 ? co : ci;
        const int k = mode == 0 ? ci : co;
        const long long base = (long long)t2 * 8192 + ((k >> 2) * 128 + n) * 4 + (k & 3);
        wimg[base] = hi;
        wimg[base + 64 * 4] = lo;
    }
}

// BatchNorm backward statistics: dy = y > 0 ? g : 0;  acc[128 + c] += sum dy, acc[192 + c] += sum dy * xhat
__global__ void __launch_bounds__(256) k_on_bn_bwd_stats(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ g,
                                                          double* __restrict__ acc, long long M, float eps) {
    __shared__ float s_m[64], s_i[64];
    __shared__ double s_red[2][16][64];
    if (threadIdx.x < 64) { float var; on_bn_final(acc, M, eps, threadIdx.x, s_m[threadIdx.x], s_i[threadIdx.x], var); }
    __syncthreads();
    const int chunk = threadIdx.x & 15, c0 = chunk * 4;
    double s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
    for (long long r = (long long)blockIdx.x * 16 + (threadIdx.x >> 4); r < M; r += (long long)gridDim.x * 16) {
        const float4 xv = reinterpret_cast<const float4*>(x)[r * 16 + chunk], yv = reinterpret_cast<const float4*>(y)[r * 16 + chunk],
                     gv = reinterpret_cast<const float4*>(g)[r * 16 + chunk];
        const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ys[4] = {yv.x, yv.y, yv.z, yv.w}, gs[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float dy = ys[k] > 0.f ? gs[k] : 0.f;
            const float xh = __fmul_rn(__fsub_rn(xs[k], s_m[c0 + k]), s_i[c0 + k]);
            s1[k] += (double)dy; s2[k] += (double)dy * (double)xh;
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) { s_red[0][threadIdx.x >> 4][c0 + k] = s1[k]; s_red[1][threadIdx.x >> 4][c0 + k] = s2[k]; }
    __syncthreads();
    if (threadIdx.x < 128) {
        const int q = threadIdx.x >> 6, c = threadIdx.x & 63;
        double t = 0;
        for (int r = 0; r < 16; ++r) t += s_red[q][r][c];
        atomicAdd(acc + (128 + q * 64 + c) * ACC_STRIDE, t);
    }
}

// BatchNorm backward dx = gamma*invstd * ((dy - mean(dy)) - xhat * mean(dy*xhat)); block 0 also writes the layer's slice of the gradient view
// (dgamma, dbeta, and the running-statistics "gradients" (running - batch) * (1 - decay), OdeGraphHelper.java:105-119).
// Destination: compact NHWC `dx` (and, img != null, the operand image of geometry gi) of a [B, Hd, Hd] tensor; the source pixel (b, h, w) of the
// [B, Hs, Hs] tensor goes to (b, h * up_s + up_off, w * up_s + up_off) — up_s = 1: same place; up_s = 2: the zero-upsampled gradient of a
// stride-2 convolution (the other pixels of dx / img stay zero: they were zeroed once and are never written).  xmask: additionally multiply by
// (x > 0): x is the ReLU output of the convolution in front (firstConvStem), whose epsilon this then is.
struct OnBnBwdOut {
    float* dx;
    float* img;
    ConvGeom gi;
    int Hs, Hd, up_s, up_off, xmask;
};
__global__ void __launch_bounds__(256) k_on_bn_bwd_dx(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ g,
                                                       const double* __restrict__ acc, long long M, float eps, const float* __restrict__ bn /*gamma,beta,mean,var*/,
                                                       float* __restrict__ gbn, float one_minus_decay, OnBnBwdOut o) {
    __shared__ float s_m[64], s_i[64], s_k[64], s_m1[64], s_m2[64];
    if (threadIdx.x < 64) {
        const int c = threadIdx.x;
        float var;
        on_bn_final(acc, M, eps, c, s_m[c], s_i[c], var);
        const double sdy = acc[(128 + c) * ACC_STRIDE], sdyxh = acc[(192 + c) * ACC_STRIDE];
        s_m1[c] = (float)(sdy / (double)M); s_m2[c] = (float)(sdyxh / (double)M);
        s_k[c] = __fmul_rn(bn[c], s_i[c]);
        if (blockIdx.x == 0) {
            gbn[c] = (float)sdyxh; gbn[64 + c] = (float)sdy;
            gbn[128 + c] = __fmul_rn(__fsub_rn(bn[128 + c], s_m[c]), one_minus_decay);
            gbn[192 + c] = __fmul_rn(__fsub_rn(bn[192 + c], var), one_minus_decay);
        }
    }
    __syncthreads();
    const long long n4 = M * 16;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const int chunk = (int)(i & 15), c = chunk * 4;
        const float4 xv = reinterpret_cast<const float4*>(x)[i], yv = reinterpret_cast<const float4*>(y)[i], gv = reinterpret_cast<const float4*>(g)[i];
        const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ys[4] = {yv.x, yv.y, yv.z, yv.w}, gs[4] = {gv.x, gv.y, gv.z, gv.w};
        float r[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float dy = ys[k] > 0.f ? gs[k] : 0.f;
            const float xh = __fmul_rn(__fsub_rn(xs[k], s_m[c + k]), s_i[c + k]);
            r[k] = __fmul_rn(s_k[c + k], __fsub_rn(__fsub_rn(dy, s_m1[c + k]), __fmul_rn(xh, s_m2[c + k])));
            if (o.xmask && !(xs[k] > 0.f)) r[k] = 0.f;
        }
        long long row = i >> 4;
        if (o.up_s != 1) {
            const int w = (int)(row % o.Hs), h = (int)((row / o.Hs) % o.Hs);
            const long long b = row / ((long long)o.Hs * o.Hs);
            row = (b * o.Hd + h * o.up_s + o.up_off) * o.Hd + w * o.up_s + o.up_off;
        }
        const float4 ov = make_float4(r[0], r[1], r[2], r[3]);
        reinterpret_cast<float4*>(o.dx)[row * 16 + chunk] = ov;
        if (o.img) store_split4(o.img, o.gi, chunk, padded_row(o.gi, row), ov);
    }
}

// sum of the per-tile partials of k_conv_wgrad_tc in tile order (deterministic) straight into DL4J's [co][ci][3][3] gradient slots:
// two levels, so that hundreds of tiles (26 x 26 maps) do not become one long dependent chain per thread
// level 1: mid[q][i] = sum of partials [q*per, (q+1)*per);  level 2: dst[co][ci][tap] = sum_q mid[q][tap][ci][co]
__global__ void __launch_bounds__(256) k_on_wgrad_sum1(const float* __restrict__ part, float* __restrict__ mid, int nparts, int per) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;      // float4 index, 9216 per partial
    const int q = blockIdx.y;
    if (i >= 9216) return;
    const int b0 = q * per, b1 = min(nparts, b0 + per);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int b = b0; b < b1; ++b) {
        const float4 v = reinterpret_cast<const float4*>(part)[(long long)b * 9216 + i];
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    reinterpret_cast<float4*>(mid)[(long long)q * 9216 + i] = s;
}
__global__ void __launch_bounds__(256) k_on_wgrad_sum2(const float* __restrict__ mid, float* __restrict__ dst, int nq) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;      // element of [tap][ci][co]
    if (i >= 36864) return;
    float s = 0.f;
    for (int q = 0; q < nq; ++q) s += mid[(long long)q * 36864 + i];
    const int co = i & 63, ci = (i >> 6) & 63, tap = i >> 12;
    dst[(co * 64 + ci) * 9 + tap] = s;
}

}  // namespace n4j
