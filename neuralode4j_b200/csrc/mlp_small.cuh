// mlp_small.cuh — one launch per RHS evaluation for small dense inner graphs (the reference's spiral latent ODE block,
// examples/spiral/LatentOdeBlock.java:31-45: 4-20-20-4; the ANODE block, examples/anode/MlpBlock.java:42-60 behind
// OdeBlockWithTime.java:26-40: (d+1)-32-32-d).
//
// Such a block is pure launch latency when every layer runs as its own kernels (an augmented evaluation of the spiral block was
// 13 launches of ~3 us each): here ONE kernel does the whole ForwardPass.calculateDerivative (ForwardPass.java:41-102) and, for
// the adjoint, BackpropagateAdjoint.backPropagate (BackpropagateAdjoint.java:87-143) for all layers.
//
// A block owns 16 batch rows; the (row, feature) items of every layer are spread over its 512 threads.  Activations and deltas
// of every layer live in shared memory, column-major [feature][row] (conflict-free).  The arithmetic is that of the exact dense kernels (layer_kernels.cuh: k_dense_*_exact): every contraction
// accumulates exact fp32 products in double and rounds once, so the result is bit-comparable with the oracle.  Parameter
// gradients: the block's threads switch roles (one parameter each, rows summed in order), per-block partials go to global
// memory and are summed in two levels of fixed order (groups of blocks, then the groups) — deterministic, no floating-point atomics.
#pragma once

#include "common.cuh"
#include "layer_kernels.cuh"

namespace n4j {

constexpr int MLP_MAX_LAYERS = 6;
constexpr int MLP_ROWS = 16;         // batch rows per block (compile-time: the code is latency-bound and wants its loops unrolled).
                                     // Measured on B200, spiral block, 1000 rows: 16 rows x 128 threads -> 29 us per augmented
                                     // evaluation; 128 rows x 512 threads (8 blocks) -> 72 us
constexpr int MLP_U = 1;             // independent accumulation chains per thread
constexpr int MLP_THREADS = 512;     // the (row, feature) items of every layer are spread over all threads of the block.  Measured on B200 (spiral
                                     // block, ms per forward + adjoint): 128 threads x 4 chains 49.0, 256 x 2 41.9, 512 x 1 36.8, 1024 x 1 36.0 —
                                     // one warp per scheduler cannot hide its own latencies, more warps with one chain each can
constexpr int MLP_LD = MLP_ROWS + 1; // doubles between consecutive features of the activation / delta records: the parameter-gradient pass reads
                                     // one ROW of many FEATURES per warp — with a stride of 16 doubles (128 B) every lane hit the same bank
                                     // (measured: 69 % of the kernel's shared-memory wavefronts were conflict replays, the pass 39 % of its samples)

struct MlpSmallArgs {
    int nl;                           // dense layers
    int tcat_k;                       // time features appended in front of the first dense layer (0: autonomous)
    int D;                            // state width
    int nin[MLP_MAX_LAYERS], nout[MLP_MAX_LAYERS], act[MLP_MAX_LAYERS], has_bias[MLP_MAX_LAYERS];
    int p_off[MLP_MAX_LAYERS];        // offset of the layer's [W | b] inside the flat parameter vector (= inside the theta block)
    int a_off[MLP_MAX_LAYERS + 1];    // offset (in features) of the layer INPUT inside the per-row activation record; [nl] = output
    int a_tot;                        // features per row record
    int n_params;
    const float* params;
    Ref in, out;
    long long B;
    long long off_a, off_theta, off_t;
    int write_t;                      // the adjoint state carries a time slot
    TimeSpec ts;
    const Ctrl* c;
    double* part;                     // [gridDim.x + groups][n_params + 1] per-block partial parameter gradients (+ time epsilon), then per-group sums
    unsigned int* ticket;             // [1 + groups]: [0] groups finished, [1 + g] blocks of group g finished
    int group;                        // blocks per group of the two-level partial sum (~sqrt(gridDim.x))
    float* dt_out;
};

// everything is held as double in shared memory (values are fp32-representable: rounded to float where the reference rounds),
// so the inner loops are LDS.64 + DFMA without conversions
inline size_t mlp_small_smem_bytes(const MlpSmallArgs& a) {
    return sizeof(double) * ((size_t)a.n_params + 2 * (size_t)a.a_tot * MLP_LD + MLP_ROWS);
}

template <bool AUG>
__global__ void __launch_bounds__(MLP_THREADS) k_mlp_small(MlpSmallArgs a) {
    constexpr int rs = 4, rmask = MLP_ROWS - 1;
    static_assert((1 << rs) == MLP_ROWS, "MLP_ROWS must be 16");
    pdl_begin();
    extern __shared__ __align__(16) unsigned char mlp_raw[];
    double* s_dt = reinterpret_cast<double*>(mlp_raw);        // [ROWS] per-row time epsilon
    double* s_w = s_dt + MLP_ROWS;                            // [n_params]
    double* s_act = s_w + a.n_params;                         // [a_tot][ROWS]: inputs of every layer, then the output
    double* s_del = s_act + (size_t)a.a_tot * MLP_LD;         // [a_tot][ROWS]: epsilon / delta at the same positions   (rows MLP_LD apart)
    const int tid = threadIdx.x;
    const long long row0 = (long long)blockIdx.x * MLP_ROWS;
    for (int i = tid; i < a.n_params; i += MLP_THREADS) s_w[i] = (double)a.params[i];
    const float* __restrict__ zin = resolve(a.in);
    float* __restrict__ out = resolve(a.out);
    // ---- graph input: [z row | t x tcat_k]  (NoTimeInput.java:32-42 / TimeInput.java:27-35 + ShapeMatchVertex merge)
    const float t = a.tcat_k ? eval_time(a.c, a.ts) : 0.f;
    for (int idx = tid; idx < (a.D + a.tcat_k) * MLP_ROWS; idx += MLP_THREADS) {
        const int r = idx & rmask, i = idx >> rs;
        const long long row = row0 + r;
        s_act[(a.a_off[0] + i) * MLP_LD + r] = i < a.D ? (row < a.B ? (double)zin[row * a.D + i] : 0.0) : (double)t;
    }
    __syncthreads();
    // ---- forward: y = act(W^T x + b), contraction in double (k_dense_fwd_exact)
    for (int l = 0; l < a.nl; ++l) {
        const int nin = a.nin[l], nout = a.nout[l];
        const double* W = s_w + a.p_off[l];
        const double* bias = a.has_bias[l] ? W + nin * nout : nullptr;
        // MLP_U items per thread at a time: each keeps its own sequential sum (bit-identical to one chain), the chains interleave
        for (int base = tid; base < nout * MLP_ROWS; base += MLP_THREADS * MLP_U) {
            double acc[MLP_U];
            const double* x[MLP_U];
            const double* wc[MLP_U];
#pragma unroll
            for (int u = 0; u < MLP_U; ++u) {
                const int idx = min(base + u * MLP_THREADS, nout * MLP_ROWS - 1);
                const int r = idx & rmask, o = idx >> rs;
                x[u] = s_act + a.a_off[l] * MLP_LD + r;
                wc[u] = W + o * nin;
                acc[u] = bias ? bias[o] : 0.0;
            }
            for (int i = 0; i < nin; ++i) {
#pragma unroll
                for (int u = 0; u < MLP_U; ++u) acc[u] += x[u][i * MLP_LD] * wc[u][i];
            }
#pragma unroll
            for (int u = 0; u < MLP_U; ++u) {
                const int idx = base + u * MLP_THREADS;
                if (idx < nout * MLP_ROWS) s_act[(a.a_off[l + 1] + (idx >> rs)) * MLP_LD + (idx & rmask)] = (double)act_fwd(a.act[l], (float)acc[u]);
            }
        }
        __syncthreads();
    }
    for (int idx = tid; idx < a.D * MLP_ROWS; idx += MLP_THREADS) {
        const int i = idx % a.D, r = idx / a.D;                     // feature fastest: contiguous global rows
        const long long row = row0 + r;
        if (row < a.B) out[row * a.D + i] = (float)s_act[(a.a_off[a.nl] + i) * MLP_LD + r];
    }
    if constexpr (!AUG) return;
    // ---- backward: epsilon at the output = -a (BackpropagateAdjoint.java:75), delta = act'(y) * epsilon, dgrad in double
    for (int idx = tid; idx < a.D * MLP_ROWS; idx += MLP_THREADS) {
        const int i = idx % a.D, r = idx / a.D;
        const long long row = row0 + r;
        s_del[(a.a_off[a.nl] + i) * MLP_LD + r] = row < a.B ? (double)__fmul_rn(-1.f, zin[a.off_a + row * a.D + i]) : 0.0;
    }
    __syncthreads();
    for (int l = a.nl - 1; l >= 0; --l) {
        const int nin = a.nin[l], nout = a.nout[l];
        const double* W = s_w + a.p_off[l];
        for (int idx = tid; idx < nout * MLP_ROWS; idx += MLP_THREADS) {      // delta in place of the epsilon w.r.t. the layer output
            const int p = (a.a_off[l + 1] + (idx >> rs)) * MLP_LD + (idx & rmask);
            s_del[p] = (double)act_bwd(a.act[l], (float)s_act[p], (float)s_del[p]);
        }
        __syncthreads();
        for (int base = tid; base < nin * MLP_ROWS; base += MLP_THREADS * MLP_U) {   // epsilon w.r.t. the layer input (k_dense_dgrad_exact)
            double acc[MLP_U];
            const double* d[MLP_U];
            const double* wr[MLP_U];
#pragma unroll
            for (int u = 0; u < MLP_U; ++u) {
                const int idx = min(base + u * MLP_THREADS, nin * MLP_ROWS - 1);
                const int r = idx & rmask, i = idx >> rs;
                d[u] = s_del + a.a_off[l + 1] * MLP_LD + r;
                wr[u] = W + i;
                acc[u] = 0.0;
            }
            for (int o = 0; o < nout; ++o) {
#pragma unroll
                for (int u = 0; u < MLP_U; ++u) acc[u] += d[u][o * MLP_LD] * wr[u][o * nin];
            }
#pragma unroll
            for (int u = 0; u < MLP_U; ++u) {
                const int idx = base + u * MLP_THREADS;
                if (idx < nin * MLP_ROWS) s_del[(a.a_off[l] + (idx >> rs)) * MLP_LD + (idx & rmask)] = (double)(float)acc[u];
            }
        }
        __syncthreads();
    }
    for (int idx = tid; idx < a.D * MLP_ROWS; idx += MLP_THREADS) {
        const int i = idx % a.D, r = idx / a.D;
        const long long row = row0 + r;
        if (row < a.B) out[a.off_a + row * a.D + i] = (float)s_del[(a.a_off[0] + i) * MLP_LD + r];     // updateZAdjoint
    }
    if (tid < MLP_ROWS) {
        double dtr = 0;
        for (int j = 0; j < a.tcat_k; ++j) dtr += s_del[(a.a_off[0] + a.D + j) * MLP_LD + tid];         // DuplicateScalarToShape.backprop
        s_dt[tid] = row0 + tid < a.B ? dtr : 0.0;
    }
    __syncthreads();
    // ---- parameter gradients of this block's rows: one parameter per thread at a time, rows in order (k_dense_wgrad_exact)
    const int np1 = a.n_params + 1;
    double* mypart = a.part + (size_t)blockIdx.x * np1;
    for (int base = tid; base < np1; base += MLP_THREADS * MLP_U) {
        double acc[MLP_U];
        const double* xa[MLP_U];      // nullptr: plain column sum (bias gradient / time epsilon)
        const double* da[MLP_U];
#pragma unroll
        for (int u = 0; u < MLP_U; ++u) {
            const int p = min(base + u * MLP_THREADS, np1 - 1);
            acc[u] = 0.0;
            if (p == a.n_params) {
                xa[u] = nullptr; da[u] = s_dt;
            } else {
                int l = 0;
                while (l + 1 < a.nl && p >= a.p_off[l + 1]) ++l;
                const int q0 = p - a.p_off[l], nin = a.nin[l], nout = a.nout[l];
                const double* d = s_del + a.a_off[l + 1] * MLP_LD;
                if (q0 < nin * nout) { xa[u] = s_act + (a.a_off[l] + q0 % nin) * MLP_LD; da[u] = d + (q0 / nin) * MLP_LD; }
                else { xa[u] = nullptr; da[u] = d + (q0 - nin * nout) * MLP_LD; }
            }
        }
        // rows in four quarters summed side by side and then in quarter order: a fixed tree (deterministic), four times shorter chains
        double accq[MLP_U][4];
#pragma unroll
        for (int u = 0; u < MLP_U; ++u)
#pragma unroll
            for (int j = 0; j < 4; ++j) accq[u][j] = 0.0;
        constexpr int quarter = MLP_ROWS >> 2;
        for (int q = 0; q < quarter; ++q) {
#pragma unroll
            for (int u = 0; u < MLP_U; ++u)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int rr = j * quarter + q;
                    accq[u][j] += xa[u] ? xa[u][rr] * da[u][rr] : da[u][rr];
                }
        }
#pragma unroll
        for (int u = 0; u < MLP_U; ++u) {
            acc[u] = ((accq[u][0] + accq[u][1]) + accq[u][2]) + accq[u][3];
            const int p = base + u * MLP_THREADS;
            if (p < np1) mypart[p] = acc[u];
        }
    }
    // ---- sum of the per-block partials, two levels, both in a fixed order (deterministic): the last block of every group of `a.group`
    // consecutive blocks sums that group's partials, the last group to finish sums the group sums.  Each level is one round of
    // independent loads per thread instead of one block walking all gridDim.x partials (measured: that tail was 40 % of the kernel).
    __threadfence();
    __syncthreads();
    __shared__ int s_last;
    const int G = a.group, grp = (int)blockIdx.x / G, ngroups = ((int)gridDim.x + G - 1) / G;
    const int gsize = min(G, (int)gridDim.x - grp * G);
    if (tid == 0) {
        const unsigned int tk = atomicAdd(a.ticket + 1 + grp, 1u);
        s_last = (tk == (unsigned int)gsize - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double* gpart = a.part + (size_t)gridDim.x * np1;          // [ngroups][np1]
    constexpr int BU = 8;
    for (int p = tid; p < np1; p += MLP_THREADS) {
        double tot = 0.0;
        for (int q0 = 0; q0 < gsize; q0 += BU) {
            double v[BU];
#pragma unroll
            for (int q = 0; q < BU; ++q) v[q] = q0 + q < gsize ? __ldcg(a.part + (size_t)(grp * G + q0 + q) * np1 + p) : 0.0;
#pragma unroll
            for (int q = 0; q < BU; ++q) tot += v[q];
        }
        gpart[(size_t)grp * np1 + p] = tot;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        a.ticket[1 + grp] = 0u;
        const unsigned int tk = atomicAdd(a.ticket, 1u);
        s_last = (tk == (unsigned int)ngroups - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int p = tid; p < np1; p += MLP_THREADS) {
        double tot = 0.0;
        for (int q0 = 0; q0 < ngroups; q0 += BU) {
            double v[BU];
#pragma unroll
            for (int q = 0; q < BU; ++q) v[q] = q0 + q < ngroups ? __ldcg(gpart + (size_t)(q0 + q) * np1 + p) : 0.0;
#pragma unroll
            for (int q = 0; q < BU; ++q) tot += v[q];
        }
        if (p == a.n_params) {
            *a.dt_out = (float)tot;
            if (a.write_t) out[a.off_t] = (float)tot;                                    // TimeInput.update: tAdjoint.assign
        } else {
            out[a.off_theta + p] = (float)tot;
        }
    }
    if (tid == 0) *a.ticket = 0u;
}

}  // namespace n4j
