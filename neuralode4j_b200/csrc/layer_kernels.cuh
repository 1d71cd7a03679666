// layer_kernels.cuh — the RHS f(z,t) of the ODE block and its vector-Jacobian products.
//
// The reference evaluates the inner ComputationGraph layer by layer through DL4J
// (ForwardPass.evaluate, src/main/java/ode/vertex/impl/helper/forward/ForwardPass.java:62-102;
// BackpropagateAdjoint.backPropagate, .../backward/BackpropagateAdjoint.java:87-143).  Here every layer is
// a CUDA kernel working on the engine's INTERNAL layout: activations are row-major [rows, channels] with
// rows = B*H*W (NHWC) or B (dense), so BatchNorm statistics are column sums and a 3x3 convolution is nine
// row-shifted GEMMs.  First/last tensors of the chain are `Ref`s resolved from the control block (the solver's
// rotating state / K rows), everything in between lives in static scratch buffers.
//
// Arithmetic contract shared with oracle/node_oracle.cpp: elementwise math in fp32 without contraction,
// reductions (BatchNorm statistics, exact dense contractions) in double.  The tiled SGEMM / implicit-GEMM
// kernels accumulate in fp32 and agree with the oracle to ~1e-6 relative.
#pragma once

#include "common.cuh"

namespace n4j {

// ---------------------------------------------------------------------------------------------------
// Layout transforms at the boundary
// ---------------------------------------------------------------------------------------------------
// NCHW [B,C,S] -> NHWC [B,S,C]   (S = H*W)
__global__ void __launch_bounds__(256) k_nchw_to_nhwc(const float* __restrict__ src, float* __restrict__ dst, int B, int C, int S) {
    pdl_begin();
    const long long n = (long long)B * C * S;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        const int c = (int)(i % C);
        const long long r = i / C;
        const int s = (int)(r % S);
        const long long b = r / S;
        dst[i] = src[(b * C + c) * S + s];
    }
}
__global__ void __launch_bounds__(256) k_nhwc_to_nchw(const float* __restrict__ src, float* __restrict__ dst, int B, int C, int S) {
    pdl_begin();
    const long long n = (long long)B * C * S;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        const int s = (int)(i % S);
        const long long r = i / S;
        const int c = (int)(r % C);
        const long long b = r / C;
        dst[i] = src[(b * S + s) * C + c];
    }
}
// time-first internal [T][n] <-> reference layouts.  rank 2: [B,F,T]; rank 4: [B,T,C,H,W] (MultiStep.java:49-60)
__global__ void __launch_bounds__(256) k_timefirst_to_ref(const float* __restrict__ src, float* __restrict__ dst, int T, int B, int C, int S, int rank4) {
    pdl_begin();
    const long long per = (long long)B * C * S;
    const long long n = per * T;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        if (rank4) {   // dst index: (((b*T + k)*C + c)*S + s)   src: [k][b][s][c]
            const int s = (int)(i % S); long long r = i / S;
            const int c = (int)(r % C); r /= C;
            const int k = (int)(r % T); const long long b = r / T;
            dst[i] = src[(long long)k * per + (b * S + s) * C + c];
        } else {       // dst index: ((b*C + f)*T + k)            src: [k][b][f]
            const int k = (int)(i % T); long long r = i / T;
            const int f = (int)(r % C); const long long b = r / C;
            dst[i] = src[(long long)k * per + b * C + f];
        }
    }
}
__global__ void __launch_bounds__(256) k_ref_to_timefirst(const float* __restrict__ src, float* __restrict__ dst, int T, int B, int C, int S, int rank4) {
    pdl_begin();
    const long long per = (long long)B * C * S;
    const long long n = per * T;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        if (rank4) {
            const int s = (int)(i % S); long long r = i / S;
            const int c = (int)(r % C); r /= C;
            const int k = (int)(r % T); const long long b = r / T;
            dst[(long long)k * per + (b * S + s) * C + c] = src[i];
        } else {
            const int k = (int)(i % T); long long r = i / T;
            const int f = (int)(r % C); const long long b = r / C;
            dst[(long long)k * per + b * C + f] = src[i];
        }
    }
}
// conv weights: DL4J [Co,Ci,3,3] c-order -> Wf [tap][ci][co] and Wb [tap][co][ci]
__global__ void __launch_bounds__(256) k_conv_w_to_internal(const float* __restrict__ w, float* __restrict__ wf, float* __restrict__ wb, int Co, int Ci) {
    pdl_begin();
    const int n = Co * Ci * 9;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int tap = i % 9; const int r = i / 9; const int ci = r % Ci; const int co = r / Ci;
        const float v = w[i];
        wf[(tap * Ci + ci) * Co + co] = v;
        wb[(tap * Co + co) * Ci + ci] = v;
    }
}
// theta-slot [tap][ci][co] <-> flat gradient view [Co,Ci,3,3]
__global__ void __launch_bounds__(256) k_conv_g_internal_to_flat(const float* __restrict__ gi, float* __restrict__ gflat, int Co, int Ci) {
    pdl_begin();
    const int n = Co * Ci * 9;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int tap = i % 9; const int r = i / 9; const int ci = r % Ci; const int co = r / Ci;
        gflat[i] = gi[(tap * Ci + ci) * Co + co];
    }
}
__global__ void __launch_bounds__(256) k_conv_g_flat_to_internal(const float* __restrict__ gflat, float* __restrict__ gi, int Co, int Ci) {
    pdl_begin();
    const int n = Co * Ci * 9;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int tap = i % 9; const int r = i / 9; const int ci = r % Ci; const int co = r / Ci;
        gi[(tap * Ci + ci) * Co + co] = gflat[i];
    }
}

// ---------------------------------------------------------------------------------------------------
// Elementwise
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_act_fwd(Ref xin, Ref yout, int act, long long n) {
    pdl_begin();
    const float* __restrict__ x = resolve(xin);
    float* __restrict__ y = resolve(yout);
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) y[i] = act_fwd(act, x[i]);
}
// g' = act'(y) * (gscale * g)
__global__ void __launch_bounds__(256) k_act_bwd(Ref gin, float gscale, Ref yref, int act, Ref gout, long long n) {
    pdl_begin();
    const float* __restrict__ g = resolve(gin);
    const float* __restrict__ y = resolve(yref);
    float* __restrict__ o = resolve(gout);
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step)
        o[i] = act_bwd(act, y[i], __fmul_rn(gscale, g[i]));
}

// ---------------------------------------------------------------------------------------------------
// Time input: the reference's "timeConc" first vertex, ShapeMatchVertex(MergeVertex) fed by TimeInput
// (ode/vertex/impl/ShapeMatchVertex.java:50-79, util/preproc/DuplicateScalarToShape.java:41-60,
// ode/vertex/impl/helper/TimeInput.java:27-50): k channels holding the current time are appended to every row.
// The time of an evaluation is derived on the device from the control block, the way the reference derives it
// from `time.add(state.timeOffset)` (FirstOrderEquationWithState.java:78-85).
// ---------------------------------------------------------------------------------------------------
struct TimeSpec {
    int mode;      // 0: t   1: t + h (every stage, quirk A.3)   2: t + coef*h (N4J_FLAG_EXACT_STAGE_TIMES)   3: value
    float coef;
    float value;
};
__device__ __forceinline__ float eval_time(const Ctrl* c, const TimeSpec& ts) {
    switch (ts.mode) {
        case 0: return c->t;
        case 1: return __fadd_rn(c->t, c->h);
        case 2: return __fadd_rn(c->t, __fmul_rn(ts.coef, c->h));
        default: return ts.value;
    }
}
// out[r][0:C] = in[r][0:C];  out[r][C:C+k] = t
__global__ void __launch_bounds__(256) k_time_concat_fwd(const Ctrl* c, TimeSpec ts, Ref xin, Ref yout, long long rows, int C, int k) {
    pdl_begin();
    const float* __restrict__ x = resolve(xin);
    float* __restrict__ y = resolve(yout);
    const float t = eval_time(c, ts);
    const int Co = C + k;
    const long long n = rows * Co;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        const long long r = i / Co;
        const int ch = (int)(i - r * Co);
        y[i] = ch < C ? x[r * C + ch] : t;
    }
}
// dx[r][0:C] = gscale*g[r][0:C];  dt = sum_r sum_{j<k} gscale*g[r][C+j]  (DuplicateScalarToShape.backprop: output.sum()).
// The sum goes to *dt_out and, when the adjoint state carries a time slot, to its K row (TimeInput.update: tAdjoint.assign).
__global__ void __launch_bounds__(256) k_time_concat_bwd(Ref gin, float gscale, Ref dxout, long long rows, int C, int k, double* acc,
                                                          unsigned int* ticket, CommDev* cm, float* dt_out, Ref slot, int write_slot) {
    pdl_begin();
    const float* __restrict__ g = resolve(gin);
    float* __restrict__ dx = resolve(dxout);
    const int Co = C + k;
    const long long n = rows * Co;
    const long long step = (long long)gridDim.x * blockDim.x;
    double v[1] = {0.0};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        const long long r = i / Co;
        const int ch = (int)(i - r * Co);
        const float gv = __fmul_rn(gscale, g[i]);
        if (ch < C) dx[r * C + ch] = gv; else v[0] += (double)gv;
    }
    if (publish_and_elect<1>(v, acc, ticket, cm) && threadIdx.x == 0) {
        *dt_out = (float)v[0];
        if (write_slot) *resolve(slot) = (float)v[0];
    }
}

// ---------------------------------------------------------------------------------------------------
// BatchNorm, training mode always (SingleStep.java:37): batch mean / biased variance over rows.
// Statistics: column sums in double, last block finalises mean and 1/sqrt(var+eps) and re-arms the
// accumulators.  Requires C % 4 == 0 and (C/4) dividing 256 (vector path) else the generic path.
// ---------------------------------------------------------------------------------------------------
struct BnScratch {
    double* acc;            // [2*C] accumulators, ACC_STRIDE doubles apart
    unsigned int* ticket;   // [1]
    float* mean;            // [C]
    float* invstd;          // [C]
    float* m_dy;            // [C]  backward: mean(dy)
    float* m_dyxh;          // [C]  backward: mean(dy*xhat)
    CommDev* cm;            // multi-GPU: statistics are summed over all ranks' rows (nullptr on a single GPU)
    long long rows_total;   // rows over all ranks
    // Deferred finalisation (single GPU, 64 channels): producers only ACCUMULATE (no fence, no ticket, no last-block read-back:
    // measured 2.0-2.2 us per launch on B200), every consumer derives mean / invstd / mean(dy) / mean(dy*xhat) from the raw double
    // sums itself — the same double expressions the last block used, so the floats are identical.  Two accumulator sets
    // alternate between consecutive evaluations; the producer of one evaluation re-arms the set of the next one.
    int deferred;
    double* accf;           // forward sums (x, x^2) of THIS evaluation, [2*64] strided
    double* accb;           // backward sums (dy, dy*xhat)
    double* accf_other;     // the sets the next evaluation will accumulate into
    double* accb_other;
    float eps;
    long long rows;         // local rows (B*H*W)
    // Sharded deferred scheme (cm != nullptr && deferred): every rank's producers only accumulate their LOCAL raw sums, exactly as on one
    // GPU — no ticket, no last-block read-back, no exchange inside the producer.  The exchange rides on the CONSUMER kernels: block 0
    // of every consumer pushes its rank's local sums of the set (two 16-byte LL cells per channel: payload halves + the set's sequence
    // number as tag) straight into every peer's mailbox, and every block adds the peers' cells to its local sums in rank order
    // (bitwise identical totals on all ranks).  The push is issued in the first instructions of the consumer, the totals are needed
    // only after its operand loads are in flight, so the NVLink flight (~1 us) hides behind work the kernel does anyway.
    // The tag is the number of times the set has been produced: its producer kernel bumps the counter (block 0), all ranks run the
    // same kernel sequence, consumers start after the producer finished (stream order).  Pushing is idempotent (later consumers of
    // the same set repeat it with the same values).
    // Default mode only (fast_final != 0; the parity mode keeps the divisions, which are the oracle's): the consumers' finalisation multiplies by
    // the precomputed 1 / rows and takes rsqrt() instead of 1 / sqrt() — double-precision division and square root are long dependent
    // instruction sequences on this part's slow FP64 pipe, and the finalisation sits between a consumer's operand loads and their first use.
    // The results differ from the exact forms in the last bit of a DOUBLE, i.e. the floats are the same up to rare rounding ties; every
    // consumer of a set goes through the same two functions (bn_fwd_final / bn_bwd_final, conv_finalize_stats), so they agree with each other.
    int fast_final;
    double inv_rows;        // 1 / rows, or 1 / rows_total on the sharded path
    unsigned int* seqf;     // sequence counters of THIS evaluation's forward / backward set (device memory of this rank)
    unsigned int* seqb;
    long long mbf, mbb;     // byte offsets inside a comm region of the two sets' mailboxes: cell[source rank][128], 16 bytes each
};

// thread c (< 64) of one block: local sums s0 (slot c) and s1 (slot 64 + c) of a set -> every peer's mailbox
__device__ __forceinline__ void bn_push2(CommDev* cm, long long mb, unsigned int tag, int c, double s0, double s1) {
    const int me = cm->rank, W = cm->world;
    const unsigned long long hi = (unsigned long long)tag << 32;
    const unsigned long long b0 = (unsigned long long)__double_as_longlong(s0), b1 = (unsigned long long)__double_as_longlong(s1);
#pragma unroll
    for (int q = 0; q < COMM_MAX_WORLD; ++q) {
        if (q < W && q != me) {
            unsigned long long* cell = reinterpret_cast<unsigned long long*>(cm->peer[q] + cm->stat_off + mb) + ((long long)me * 128 + c) * 2;
            asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(cell), "l"(hi | (b0 & 0xFFFFFFFFull)), "l"(hi | (b0 >> 32)) : "memory");
            asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(cell + 128), "l"(hi | (b1 & 0xFFFFFFFFull)), "l"(hi | (b1 >> 32)) : "memory");
        }
    }
}
// one peer's cells of channel c: spins until both sums carry the tag (bounded), returns them as doubles
__device__ __forceinline__ void bn_poll_peer(CommDev* cm, const unsigned long long* mbox, int q, unsigned int tag, int c, double& p0, double& p1) {
    const unsigned long long* cell = mbox + ((long long)q * 128 + c) * 2;
    unsigned long long a0 = 0, a1 = 0, b0 = 0, b1 = 0;
    bool ok = false;
    for (unsigned int spin = 0; spin < (1u << 26); ++spin) {
        asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a0), "=l"(a1) : "l"(cell) : "memory");
        asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(b0), "=l"(b1) : "l"(cell + 128) : "memory");
        if ((unsigned int)(a0 >> 32) == tag && (unsigned int)(a1 >> 32) == tag && (unsigned int)(b0 >> 32) == tag && (unsigned int)(b1 >> 32) == tag) { ok = true; break; }
        if ((spin & 255) == 255) __nanosleep(32);
    }
    if (!ok) cm->status = 1;      // the controller turns this into N4J_ERR_COMM at the next error norm
    p0 = __longlong_as_double((long long)((a0 & 0xFFFFFFFFull) | (a1 << 32)));
    p1 = __longlong_as_double((long long)((b0 & 0xFFFFFFFFull) | (b1 << 32)));
}
// any thread c (< 64): s0 / s1 (local sums in, global totals out), peers' contributions added in rank order.
// BATCH: the cells of ALL peers are requested before the first tag is looked at (one L2 round trip for the whole group instead of one
// per peer: at eight ranks 0.3 us instead of 2 us) — 32 more registers, for kernels that have them; the conv kernels gather with all
// sixteen warps instead (conv_tc.cuh).
template <bool BATCH>
__device__ __forceinline__ void bn_gather2(CommDev* cm, long long mb, unsigned int tag, int c, double& s0, double& s1) {
    const int me = cm->rank, W = cm->world;
    const unsigned long long* mbox = reinterpret_cast<const unsigned long long*>(cm->peer[me] + cm->stat_off + mb);
    double t0 = 0.0, t1 = 0.0;
    if (BATCH) {
        unsigned long long v[COMM_MAX_WORLD][4];
        bool ok = false;
        for (unsigned int spin = 0; spin < (1u << 26) && !ok; ++spin) {
            ok = true;
#pragma unroll
            for (int q = 0; q < COMM_MAX_WORLD; ++q) {
                if (q < W && q != me) {
                    const unsigned long long* cell = mbox + ((long long)q * 128 + c) * 2;
                    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(v[q][0]), "=l"(v[q][1]) : "l"(cell) : "memory");
                    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(v[q][2]), "=l"(v[q][3]) : "l"(cell + 128) : "memory");
                }
            }
#pragma unroll
            for (int q = 0; q < COMM_MAX_WORLD; ++q)
                if (q < W && q != me &&
                    ((unsigned int)(v[q][0] >> 32) != tag || (unsigned int)(v[q][1] >> 32) != tag || (unsigned int)(v[q][2] >> 32) != tag || (unsigned int)(v[q][3] >> 32) != tag))
                    ok = false;
            if (!ok && (spin & 255) == 255) __nanosleep(32);
        }
        if (!ok) cm->status = 1;
#pragma unroll
        for (int q = 0; q < COMM_MAX_WORLD; ++q) {
            if (q >= W) continue;
            if (q == me) { t0 += s0; t1 += s1; continue; }
            t0 += __longlong_as_double((long long)((v[q][0] & 0xFFFFFFFFull) | (v[q][1] << 32)));
            t1 += __longlong_as_double((long long)((v[q][2] & 0xFFFFFFFFull) | (v[q][3] << 32)));
        }
    } else {
        for (int q = 0; q < W; ++q) {
            if (q == me) { t0 += s0; t1 += s1; continue; }
            double p0, p1;
            bn_poll_peer(cm, mbox, q, tag, c, p0, p1);
            t0 += p0; t1 += p1;
        }
    }
    s0 = t0; s1 = t1;
}
__device__ __forceinline__ bool bn_first_block() { return blockIdx.x == 0 && blockIdx.y == 0; }

template <bool SH = true, bool BATCH = true>
__device__ __forceinline__ void bn_fwd_final(const BnScratch& sc, int c, float& mean, float& inv) {
    double s0 = sc.accf[c * ACC_STRIDE], s1 = sc.accf[(64 + c) * ACC_STRIDE];
    double M = (double)sc.rows;
    if (SH && sc.cm) {
        const unsigned int tag = *((volatile unsigned int*)sc.seqf);
        if (bn_first_block()) bn_push2(sc.cm, sc.mbf, tag, c, s0, s1);
        bn_gather2<BATCH>(sc.cm, sc.mbf, tag, c, s0, s1);
        M = (double)sc.rows_total;
    }
    if (sc.fast_final) {
        const double m = s0 * sc.inv_rows;
        double var = s1 * sc.inv_rows - m * m;
        if (var < 0) var = 0;
        mean = (float)m;
        inv = (float)rsqrt(var + (double)sc.eps);
        return;
    }
    const double m = s0 / M;
    double var = s1 / M - m * m;
    if (var < 0) var = 0;
    mean = (float)m;
    inv = (float)(1.0 / sqrt(var + (double)sc.eps));
}
__device__ __forceinline__ void bn_bwd_final(const BnScratch& sc, int c, float& m_dy, float& m_dyxh, float& dgamma, float& dbeta) {
    double sdy = sc.accb[c * ACC_STRIDE], sdyxh = sc.accb[(64 + c) * ACC_STRIDE];
    double M = (double)sc.rows;
    if (sc.cm) {
        const unsigned int tag = *((volatile unsigned int*)sc.seqb);
        if (bn_first_block()) bn_push2(sc.cm, sc.mbb, tag, c, sdy, sdyxh);
        bn_gather2<true>(sc.cm, sc.mbb, tag, c, sdy, sdyxh);
        M = (double)sc.rows_total;
    }
    dgamma = (float)sdyxh;
    dbeta = (float)sdy;
    if (sc.fast_final) { m_dy = (float)(sdy * sc.inv_rows); m_dyxh = (float)(sdyxh * sc.inv_rows); return; }
    m_dy = (float)(sdy / M);
    m_dyxh = (float)(sdyxh / M);
}
// producers: block 0 re-arms the accumulator set of the NEXT evaluation (all its readers finished before this kernel started)
__device__ __forceinline__ void bn_rearm(double* other, int tid) { if (blockIdx.x == 0 && tid < 128) other[tid * ACC_STRIDE] = 0.0; }
// ... and, on the sharded path, announces that the set it accumulates into has been produced once more (the consumers' tag)
__device__ __forceinline__ void bn_bump(unsigned int* seq, int tid) { if (seq && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) *seq += 1u; }

// Last block, all threads: gather the 2*C strided accumulators into tot[] (shared), re-arm them, and (multi-GPU) sum over the ranks.
__device__ __forceinline__ void bn_collect(const BnScratch& sc, int C, double* tot) {
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
        tot[i] = *((volatile double*)(sc.acc + i * ACC_STRIDE));
        sc.acc[i * ACC_STRIDE] = 0.0;
    }
    __syncthreads();
    if (sc.cm) comm_allreduce_block(sc.cm, tot, 2 * C);
}

__global__ void __launch_bounds__(256) k_bn_stats(Ref xin, long long M, int C, float eps, BnScratch sc) {
    pdl_begin();
    const float* __restrict__ x = resolve(xin);
    const int cg = C >> 2;                       // float4 groups per row
    const int rows_per_iter = 256 / cg;
    const int my_g = threadIdx.x % cg, my_r = threadIdx.x / cg;
    double s[4] = {0, 0, 0, 0}, ss[4] = {0, 0, 0, 0};
    for (long long r = (long long)blockIdx.x * rows_per_iter + my_r; r < M; r += (long long)gridDim.x * rows_per_iter) {
        const float4 v = reinterpret_cast<const float4*>(x + r * C)[my_g];
        s[0] += (double)v.x; ss[0] += (double)v.x * (double)v.x;
        s[1] += (double)v.y; ss[1] += (double)v.y * (double)v.y;
        s[2] += (double)v.z; ss[2] += (double)v.z * (double)v.z;
        s[3] += (double)v.w; ss[3] += (double)v.w * (double)v.w;
    }
    // reduce over the rows_per_iter threads that share a channel group
    __shared__ double sm[2][4][256];
    __shared__ int s_last;
#pragma unroll
    for (int k = 0; k < 4; ++k) { sm[0][k][threadIdx.x] = s[k]; sm[1][k][threadIdx.x] = ss[k]; }
    __syncthreads();
    if (threadIdx.x < cg) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double a = 0, b = 0;
            for (int r = 0; r < rows_per_iter; ++r) { a += sm[0][k][r * cg + threadIdx.x]; b += sm[1][k][r * cg + threadIdx.x]; }
            const int c = threadIdx.x * 4 + k;
            atomicAdd((sc.deferred ? sc.accf : sc.acc) + c * ACC_STRIDE, a);
            atomicAdd((sc.deferred ? sc.accf : sc.acc) + (C + c) * ACC_STRIDE, b);
        }
        __threadfence();
    }
    if (sc.deferred) { bn_rearm(sc.accf_other, threadIdx.x); bn_bump(sc.cm ? sc.seqf : nullptr, threadIdx.x); return; }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int tk = atomicAdd(sc.ticket, 1u);
        s_last = (tk == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        double* tot = &sm[0][0][0];                 // 2048 doubles, free again
        bn_collect(sc, C, tot);
        const double Mt = sc.cm ? (double)sc.rows_total : (double)M;
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            const double mean = tot[c] / Mt;
            double var = tot[C + c] / Mt - mean * mean;
            if (var < 0) var = 0;
            sc.mean[c] = (float)mean;
            sc.invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
        }
        if (threadIdx.x == 0) *sc.ticket = 0u;
    }
}
// generic channel count: one block per channel
__global__ void __launch_bounds__(256) k_bn_stats_generic(Ref xin, long long M, int C, float eps, BnScratch sc) {
    pdl_begin();
    const float* __restrict__ x = resolve(xin);
    const int c = blockIdx.x;
    double v[2] = {0, 0};
    for (long long r = threadIdx.x; r < M; r += blockDim.x) { const double q = (double)x[r * C + c]; v[0] += q; v[1] += q * q; }
    block_sum<2>(v);
    if (threadIdx.x == 0) {
        const double mean = v[0] / (double)M;
        double var = v[1] / (double)M - mean * mean;
        if (var < 0) var = 0;
        sc.mean[c] = (float)mean;
        sc.invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    }
}
// y = act(gamma * ((x - mean) * invstd) + beta)
__global__ void __launch_bounds__(256) k_bn_apply(Ref xin, Ref yout, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                   BnScratch sc, int act, long long M, int C) {
    pdl_begin();
    const float* __restrict__ x = resolve(xin);
    float* __restrict__ y = resolve(yout);
    const long long n = M * C;
    const long long step = (long long)gridDim.x * blockDim.x;
    __shared__ float s_mean[64], s_inv[64];
    if (sc.deferred) {
        if (threadIdx.x < 64) bn_fwd_final(sc, threadIdx.x, s_mean[threadIdx.x], s_inv[threadIdx.x]);
        __syncthreads();
    }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        const int c = (int)(i % C);
        const float d = __fsub_rn(x[i], sc.deferred ? s_mean[c] : sc.mean[c]);
        const float h = __fmul_rn(d, sc.deferred ? s_inv[c] : sc.invstd[c]);
        y[i] = act_fwd(act, __fadd_rn(__fmul_rn(gamma[c], h), beta[c]));
    }
}
// 64-channel float4 forms of the two elementwise BatchNorm kernels left on the MNIST chain.  A thread owns one 4-channel chunk
// (threadIdx & 15) and rows (global thread / 16) + k * (threads / 16).  With deferred statistics the loads of the FIRST item are
// issued before the finalisation (64 threads, double division / sqrt, ~0.7-1.5 us) so that the two latencies overlap.
__global__ void __launch_bounds__(256) k_bn_apply64(Ref xin, Ref yout, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                     BnScratch sc, int act, long long M) {
    pdl_begin();
    const float4* __restrict__ x = reinterpret_cast<const float4*>(resolve(xin));
    float4* __restrict__ y = reinterpret_cast<float4*>(resolve(yout));
    const long long n4 = M * 16, step = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c0 = (threadIdx.x & 15) * 4;
    float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n4) xv = x[i];
    __shared__ float s_mean[64], s_inv[64];
    if (threadIdx.x < 64) {
        if (sc.deferred) bn_fwd_final(sc, threadIdx.x, s_mean[threadIdx.x], s_inv[threadIdx.x]);
        else { s_mean[threadIdx.x] = sc.mean[threadIdx.x]; s_inv[threadIdx.x] = sc.invstd[threadIdx.x]; }
    }
    __syncthreads();
    float mean[4], inv[4], ga[4], be[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { mean[k] = s_mean[c0 + k]; inv[k] = s_inv[c0 + k]; ga[k] = gamma[c0 + k]; be[k] = beta[c0 + k]; }
    for (; i < n4; i += step) {
        const float xx[4] = {xv.x, xv.y, xv.z, xv.w};
        if (i + step < n4) xv = x[i + step];
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] = act_fwd(act, __fadd_rn(__fmul_rn(ga[k], __fmul_rn(__fsub_rn(xx[k], mean[k]), inv[k])), be[k]));
        y[i] = make_float4(o[0], o[1], o[2], o[3]);
    }
}
// dx = gamma*invstd * ((dy - mean(dy)) - xhat*mean(dy*xhat)),  dy = act'(y) * (gscale*g)
__global__ void __launch_bounds__(256) k_bn_bwd_dx64(Ref gin, float gscale, Ref yref, Ref xin, const float* __restrict__ gamma, BnScratch sc,
                                                      int act, long long M, Ref dxout, Ref dgamma_ref, Ref dbeta_ref) {
    pdl_begin();
    const float4* __restrict__ g = reinterpret_cast<const float4*>(resolve(gin));
    const float4* __restrict__ y = reinterpret_cast<const float4*>(resolve(yref));
    const float4* __restrict__ x = reinterpret_cast<const float4*>(resolve(xin));
    float4* __restrict__ dx = reinterpret_cast<float4*>(resolve(dxout));
    const long long n4 = M * 16, step = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c0 = (threadIdx.x & 15) * 4;
    float4 gv = make_float4(0.f, 0.f, 0.f, 0.f), yv = gv, xv = gv;
    if (i < n4) { gv = g[i]; yv = y[i]; xv = x[i]; }
    __shared__ float s_st[4][64];      // mean, invstd, mean(dy), mean(dy*xhat)
    if (threadIdx.x < 64) {
        const int c = threadIdx.x;
        if (sc.deferred) {
            float dg, db;
            bn_fwd_final(sc, c, s_st[0][c], s_st[1][c]);
            bn_bwd_final(sc, c, s_st[2][c], s_st[3][c], dg, db);
            if (blockIdx.x == 0) { resolve(dgamma_ref)[c] = dg; resolve(dbeta_ref)[c] = db; }
        } else {
            s_st[0][c] = sc.mean[c]; s_st[1][c] = sc.invstd[c]; s_st[2][c] = sc.m_dy[c]; s_st[3][c] = sc.m_dyxh[c];
        }
    }
    __syncthreads();
    float mean[4], inv[4], kk[4], mdy[4], mdyxh[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        mean[k] = s_st[0][c0 + k]; inv[k] = s_st[1][c0 + k]; mdy[k] = s_st[2][c0 + k]; mdyxh[k] = s_st[3][c0 + k];
        kk[k] = __fmul_rn(gamma[c0 + k], inv[k]);
    }
    for (; i < n4; i += step) {
        const float gg[4] = {gv.x, gv.y, gv.z, gv.w}, yy[4] = {yv.x, yv.y, yv.z, yv.w}, xx[4] = {xv.x, xv.y, xv.z, xv.w};
        if (i + step < n4) { gv = g[i + step]; yv = y[i + step]; xv = x[i + step]; }
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float dy = act_bwd(act, yy[k], __fmul_rn(gscale, gg[k]));
            const float xh = __fmul_rn(__fsub_rn(xx[k], mean[k]), inv[k]);
            o[k] = __fmul_rn(kk[k], __fsub_rn(__fsub_rn(dy, mdy[k]), __fmul_rn(xh, mdyxh[k])));
        }
        dx[i] = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// backward statistics: dy = act'(y)*(gscale*g);  sums of dy and dy*xhat (double) -> dgamma, dbeta, means
__global__ void __launch_bounds__(256) k_bn_bwd_stats(Ref gin, float gscale, Ref yref, Ref xin, BnScratch sc, int act, long long M, int C,
                                                       Ref dgamma_ref, Ref dbeta_ref) {
    pdl_begin();
    const float* __restrict__ g = resolve(gin);
    const float* __restrict__ y = resolve(yref);
    const float* __restrict__ x = resolve(xin);
    // generic mapping: thread owns channel (tid % C) when C <= 256 and divides 256, else strided
    const int c_lane = threadIdx.x % C;
    const bool fast = (C <= 256) && (256 % C == 0);
    __shared__ double sm[2][256];
    __shared__ int s_last;
    if (fast) {
        const int rows_per_iter = 256 / C, my_r = threadIdx.x / C;
        const float mean = sc.mean[c_lane], inv = sc.invstd[c_lane];
        double a = 0, b = 0;
        for (long long r = (long long)blockIdx.x * rows_per_iter + my_r; r < M; r += (long long)gridDim.x * rows_per_iter) {
            const long long i = r * C + c_lane;
            const float dy = act_bwd(act, y[i], __fmul_rn(gscale, g[i]));
            const float xh = __fmul_rn(__fsub_rn(x[i], mean), inv);
            a += (double)dy; b += (double)dy * (double)xh;
        }
        sm[0][threadIdx.x] = a; sm[1][threadIdx.x] = b;
        __syncthreads();
        if (threadIdx.x < C) {
            double sa = 0, sb = 0;
            for (int r = 0; r < rows_per_iter; ++r) { sa += sm[0][r * C + threadIdx.x]; sb += sm[1][r * C + threadIdx.x]; }
            atomicAdd(sc.acc + threadIdx.x * ACC_STRIDE, sa);
            atomicAdd(sc.acc + (C + threadIdx.x) * ACC_STRIDE, sb);
            __threadfence();
        }
    } else {
        // slow path: block loops over channels; grid-stride over rows inside
        for (int c = 0; c < C; ++c) {
            double v[2] = {0, 0};
            const float mean = sc.mean[c], inv = sc.invstd[c];
            for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < M; r += (long long)gridDim.x * blockDim.x) {
                const long long i = r * C + c;
                const float dy = act_bwd(act, y[i], __fmul_rn(gscale, g[i]));
                const float xh = __fmul_rn(__fsub_rn(x[i], mean), inv);
                v[0] += (double)dy; v[1] += (double)dy * (double)xh;
            }
            block_sum<2>(v);
            if (threadIdx.x == 0) { atomicAdd(sc.acc + c * ACC_STRIDE, v[0]); atomicAdd(sc.acc + (C + c) * ACC_STRIDE, v[1]); __threadfence(); }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int tk = atomicAdd(sc.ticket, 1u);
        s_last = (tk == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        float* __restrict__ dgamma = resolve(dgamma_ref);
        float* __restrict__ dbeta = resolve(dbeta_ref);
        __shared__ double tot[512];
        bn_collect(sc, C, tot);
        const double Mt = sc.cm ? (double)sc.rows_total : (double)M;
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            const double sdy = tot[c], sdyxh = tot[C + c];
            dgamma[c] = (float)sdyxh;
            dbeta[c] = (float)sdy;
            sc.m_dy[c] = (float)(sdy / Mt);
            sc.m_dyxh[c] = (float)(sdyxh / Mt);
        }
        if (threadIdx.x == 0) *sc.ticket = 0u;
    }
}
// 64-channel fast path of the backward statistics: float4 loads, 16 rows per block iteration, one block per SM.
// Optionally also applies the FORWARD BatchNorm (fuse_fwd): y = act(gamma*xhat+beta) is written to `yout` in the same pass,
// which merges the last forward kernel of an augmented evaluation with the first backward one.
__global__ void __launch_bounds__(256) k_bn_bwd_stats64(Ref gin, float gscale, Ref yref, Ref xin, BnScratch sc, int act, long long M,
                                                         Ref dgamma_ref, Ref dbeta_ref, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, int fuse_fwd) {
    pdl_begin();
    const float* __restrict__ g = resolve(gin);
    float* __restrict__ y = resolve(yref);
    const float* __restrict__ x = resolve(xin);
    const int chunk = threadIdx.x & 15, rg = threadIdx.x >> 4;
    const long long rstep = (long long)gridDim.x * 16;
    long long r = (long long)blockIdx.x * 16 + rg;
    // the loads of the first row are issued before the statistics are finalised (the two latencies overlap)
    float4 gv = make_float4(0.f, 0.f, 0.f, 0.f), xv = gv, yv = gv;
    if (r < M) {
        gv = reinterpret_cast<const float4*>(g)[r * 16 + chunk];
        xv = reinterpret_cast<const float4*>(x)[r * 16 + chunk];
        if (!fuse_fwd) yv = reinterpret_cast<const float4*>(y)[r * 16 + chunk];
    }
    float mean[4], inv[4], ga[4], be[4];
    __shared__ float s_mi[2][64];
    if (sc.deferred) {       // one channel per thread, once per block (double division / square root are long sequences)
        if (threadIdx.x < 64) bn_fwd_final(sc, threadIdx.x, s_mi[0][threadIdx.x], s_mi[1][threadIdx.x]);
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (sc.deferred) { mean[k] = s_mi[0][chunk * 4 + k]; inv[k] = s_mi[1][chunk * 4 + k]; }
        else { mean[k] = sc.mean[chunk * 4 + k]; inv[k] = sc.invstd[chunk * 4 + k]; }
        ga[k] = fuse_fwd ? gamma[chunk * 4 + k] : 0.f; be[k] = fuse_fwd ? beta[chunk * 4 + k] : 0.f;
    }
    double a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
    for (; r < M; r += rstep) {
        const long long i = r * 16 + chunk;
        const float gg[4] = {gv.x, gv.y, gv.z, gv.w}, xx[4] = {xv.x, xv.y, xv.z, xv.w};
        float yy[4] = {yv.x, yv.y, yv.z, yv.w};
        if (r + rstep < M) {
            const long long in = (r + rstep) * 16 + chunk;
            gv = reinterpret_cast<const float4*>(g)[in];
            xv = reinterpret_cast<const float4*>(x)[in];
            if (!fuse_fwd) yv = reinterpret_cast<const float4*>(y)[in];
        }
        if (fuse_fwd) {
#pragma unroll
            for (int k = 0; k < 4; ++k) yy[k] = act_fwd(act, __fadd_rn(__fmul_rn(ga[k], __fmul_rn(__fsub_rn(xx[k], mean[k]), inv[k])), be[k]));
            reinterpret_cast<float4*>(y)[i] = make_float4(yy[0], yy[1], yy[2], yy[3]);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float dy = act_bwd(act, yy[k], __fmul_rn(gscale, gg[k]));
            const float xh = __fmul_rn(__fsub_rn(xx[k], mean[k]), inv[k]);
            a[k] += (double)dy; b[k] += (double)dy * (double)xh;
        }
    }
    __shared__ double sm[2][4][256];
    __shared__ int s_last;
#pragma unroll
    for (int k = 0; k < 4; ++k) { sm[0][k][threadIdx.x] = a[k]; sm[1][k][threadIdx.x] = b[k]; }
    __syncthreads();
    if (threadIdx.x < 16) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double sa = 0, sb = 0;
            for (int r = 0; r < 16; ++r) { sa += sm[0][k][r * 16 + threadIdx.x]; sb += sm[1][k][r * 16 + threadIdx.x]; }
            atomicAdd((sc.deferred ? sc.accb : sc.acc) + (threadIdx.x * 4 + k) * ACC_STRIDE, sa);
            atomicAdd((sc.deferred ? sc.accb : sc.acc) + (64 + threadIdx.x * 4 + k) * ACC_STRIDE, sb);
        }
        __threadfence();
    }
    if (sc.deferred) { bn_rearm(sc.accb_other, threadIdx.x); bn_bump(sc.cm ? sc.seqb : nullptr, threadIdx.x); return; }   // dgamma / dbeta / means: written by the dx kernel behind this one
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int tk = atomicAdd(sc.ticket, 1u);
        s_last = (tk == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        float* __restrict__ dgamma = resolve(dgamma_ref);
        float* __restrict__ dbeta = resolve(dbeta_ref);
        double* tot = &sm[0][0][0];
        bn_collect(sc, 64, tot);
        const double Mt = sc.cm ? (double)sc.rows_total : (double)M;
        if (threadIdx.x < 64) {
            const int c = threadIdx.x;
            const double sdy = tot[c], sdyxh = tot[64 + c];
            dgamma[c] = (float)sdyxh;
            dbeta[c] = (float)sdy;
            sc.m_dy[c] = (float)(sdy / Mt);
            sc.m_dyxh[c] = (float)(sdyxh / Mt);
        }
        if (threadIdx.x == 0) *sc.ticket = 0u;
    }
}

// dx = gamma*invstd * ((dy - mean(dy)) - xhat*mean(dy*xhat))
__global__ void __launch_bounds__(256) k_bn_bwd_dx(Ref gin, float gscale, Ref yref, Ref xin, const float* __restrict__ gamma, BnScratch sc,
                                                    int act, long long M, int C, Ref dxout, Ref dgamma_ref, Ref dbeta_ref,
                                                    const float* __restrict__ beta) {
    pdl_begin();
    const float* __restrict__ g = resolve(gin);
    const float* __restrict__ y = resolve(yref);
    const float* __restrict__ x = resolve(xin);
    float* __restrict__ dx = resolve(dxout);
    const long long n = M * C;
    const long long step = (long long)gridDim.x * blockDim.x;
    __shared__ float s_st[4][64];      // mean, invstd, mean(dy), mean(dy*xhat)
    if (sc.deferred) {
        if (threadIdx.x < 64) {
            const int c = threadIdx.x;
            float dg, db;
            bn_fwd_final(sc, c, s_st[0][c], s_st[1][c]);
            bn_bwd_final(sc, c, s_st[2][c], s_st[3][c], dg, db);
            if (blockIdx.x == 0) { resolve(dgamma_ref)[c] = dg; resolve(dbeta_ref)[c] = db; }
        }
        __syncthreads();
    }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        const int c = (int)(i % C);
        const float inv = sc.deferred ? s_st[1][c] : sc.invstd[c];
        const float xh = __fmul_rn(__fsub_rn(x[i], sc.deferred ? s_st[0][c] : sc.mean[c]), inv);
        const float dy = act_bwd(act, y[i], __fmul_rn(gscale, g[i]));      // (recomputing the ReLU mask from x measured slower here: 7.6 vs 6.4 us)
        const float k = __fmul_rn(gamma[c], inv);
        const float t1 = __fsub_rn(dy, sc.deferred ? s_st[2][c] : sc.m_dy[c]);
        const float t2 = __fmul_rn(xh, sc.deferred ? s_st[3][c] : sc.m_dyxh[c]);
        dx[i] = __fmul_rn(k, __fsub_rn(t1, t2));
    }
}

// ---------------------------------------------------------------------------------------------------
// Dense, exact path (small layers: spiral 4-20-20-4, the reference's golden 2x2 case): contractions in double,
// bit-compatible with the oracle.  W is DL4J's f-order [nIn,nOut]: W(i,o) = Wt[i + o*nIn], bias after it.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_dense_fwd_exact(Ref xin, Ref yout, const float* __restrict__ Wt, const float* __restrict__ bias,
                                                          int act, long long B, int nin, int nout) {
    pdl_begin();
    const float* __restrict__ x = resolve(xin);
    float* __restrict__ y = resolve(yout);
    const long long n = B * nout;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += step) {
        const int o = (int)(idx % nout);
        const long long b = idx / nout;
        double acc = bias ? (double)bias[o] : 0.0;
        const float* xr = x + b * nin;
        const float* wc = Wt + (long long)o * nin;
        for (int i = 0; i < nin; ++i) acc += (double)xr[i] * (double)wc[i];
        y[idx] = act_fwd(act, (float)acc);
    }
}
// dW(i,o) = sum_b x(b,i) d(b,o)  -> out[i + o*nIn];  db(o) = sum_b d(b,o) -> out[nIn*nOut + o]
// one block per output element: the batch is split over the 128 threads (double partial sums, block reduce)
__global__ void __launch_bounds__(128) k_dense_wgrad_exact(Ref xin, Ref din, Ref dwout, long long B, int nin, int nout, int has_bias) {
    pdl_begin();
    const float* __restrict__ x = resolve(xin);
    const float* __restrict__ d = resolve(din);
    float* __restrict__ dw = resolve(dwout);
    const int n = nin * nout + (has_bias ? nout : 0);
    for (int idx = blockIdx.x; idx < n; idx += gridDim.x) {
        double acc[1] = {0.0};
        if (idx < nin * nout) {
            const int i = idx % nin, o = idx / nin;
            for (long long b = threadIdx.x; b < B; b += blockDim.x) acc[0] += (double)x[b * nin + i] * (double)d[b * nout + o];
        } else {
            const int o = idx - nin * nout;
            for (long long b = threadIdx.x; b < B; b += blockDim.x) acc[0] += (double)d[b * nout + o];
        }
        block_sum<1>(acc);
        if (threadIdx.x == 0) dw[idx] = (float)acc[0];
    }
}
__global__ void __launch_bounds__(256) k_dense_dgrad_exact(Ref din, const float* __restrict__ Wt, Ref dxout, long long B, int nin, int nout) {
    pdl_begin();
    const float* __restrict__ d = resolve(din);
    float* __restrict__ dx = resolve(dxout);
    const long long n = B * nin;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += step) {
        const int i = (int)(idx % nin);
        const long long b = idx / nin;
        double acc = 0;
        for (int o = 0; o < nout; ++o) acc += (double)d[b * nout + o] * (double)Wt[i + (long long)o * nin];
        dx[idx] = (float)acc;
    }
}
constexpr int COLSUM_RG = 32;      // row groups per block (block = 32 columns x COLSUM_RG row groups)
__global__ void __launch_bounds__(32 * COLSUM_RG) k_colsum(Ref din, Ref out, long long B, int n, int negate) {   // bias gradient: column sums in double
    pdl_begin();
    const float* __restrict__ d = resolve(din);
    float* __restrict__ o = resolve(out);
    __shared__ double sm[COLSUM_RG][33];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), rg = threadIdx.x >> 5;
    double acc = 0;
    if (c < n) {
        // rows b = rg, rg + RG, ... summed in order; eight independent loads in flight per trip (the loop is latency-bound: with 8 row
        // groups the 1024-row bias gradient of the time-series block took 12.5 us, sixteen dependent trips)
        long long b = rg;
        for (; b + 7 * COLSUM_RG < B; b += 8 * COLSUM_RG) {
            float v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = d[(b + COLSUM_RG * q) * n + c];
#pragma unroll
            for (int q = 0; q < 8; ++q) acc += (double)v[q];
        }
        for (; b < B; b += COLSUM_RG) acc += (double)d[b * n + c];
    }
    sm[rg][threadIdx.x & 31] = acc;
    __syncthreads();
    if (rg == 0 && c < n) {
        double t = 0;
#pragma unroll
        for (int r = 0; r < COLSUM_RG; ++r) t += sm[r][threadIdx.x];
        o[c] = negate ? -(float)t : (float)t;
    }
}

// ---------------------------------------------------------------------------------------------------
// Tiled fp32 GEMM with pluggable operand addressing: C[M,N] (+bias, act) = A[M,K] * B[K,N].
// 64x64 block tile, 16-deep K tile, 256 threads, 4x4 register tile.  Serves dense fwd/dgrad/wgrad and the
// 3x3 convolution as implicit GEMM (fwd, dgrad, wgrad with split-K).  This is the SIMT baseline the tcgen05
// kernels are checked against; it stays as the fallback for shapes the tensor-core path does not take.
// ---------------------------------------------------------------------------------------------------
enum { G_DENSE_FWD = 0, G_DENSE_DGRAD = 1, G_DENSE_WGRAD = 2, G_CONV_FWD = 3, G_CONV_DGRAD = 4, G_CONV_WGRAD = 5 };

struct GemmArgs {
    Ref a, b, c;
    const float* w;        // weights when an operand is a static parameter buffer (else null)
    const float* bias;     // dense fwd
    int M, N, K;
    int act;
    int H, W, Ci, Co;      // conv geometry
    int splits;            // split-K factor (grid.z); partials go to c + z*M*N when splits > 1
    int negate;            // store -C: the epsilon of the block's last layer is -a (BackpropagateAdjoint.java:75); folding the sign into the
                           // weight-gradient / dgrad GEMMs saves the kernel that would materialise -a first (exact: a sign flip does not round)
};

template <int MODE> struct Addr {
    // value of A(m,k) / B(k,n); pa/pb are the resolved operand base pointers
    __device__ static __forceinline__ float A(const GemmArgs& g, const float* pa, int m, int k, int hh, int ww) {
        if (MODE == G_DENSE_FWD || MODE == G_DENSE_DGRAD) return pa[(long long)m * g.K + k];
        if (MODE == G_DENSE_WGRAD) return pa[(long long)k * g.M + m];
        if (MODE == G_CONV_FWD || MODE == G_CONV_DGRAD) {
            const int cin = MODE == G_CONV_FWD ? g.Ci : g.Co;
            const int tap = k / cin, ch = k - tap * cin;
            int dh = tap / 3 - 1, dw = tap % 3 - 1;
            if (MODE == G_CONV_DGRAD) { dh = -dh; dw = -dw; }
            const int h2 = hh + dh, w2 = ww + dw;
            if ((unsigned)h2 >= (unsigned)g.H || (unsigned)w2 >= (unsigned)g.W) return 0.f;
            return pa[((long long)m + dh * g.W + dw) * cin + ch];
        }
        // G_CONV_WGRAD: m = tap*Ci + ci, k = pixel row; hh/ww are unused, decode pixel here
        {
            const int tap = m / g.Ci, ci = m - tap * g.Ci;
            const int dh = tap / 3 - 1, dw = tap % 3 - 1;
            const int s = k % (g.H * g.W);
            const int h2 = s / g.W + dh, w2 = s % g.W + dw;
            if ((unsigned)h2 >= (unsigned)g.H || (unsigned)w2 >= (unsigned)g.W) return 0.f;
            return pa[((long long)k + dh * g.W + dw) * g.Ci + ci];
        }
    }
    __device__ static __forceinline__ float B(const GemmArgs& g, const float* pb, int k, int n) {
        if (MODE == G_DENSE_FWD) return pb[(long long)n * g.K + k];          // W(i=k,o=n) = Wt[k + n*nIn]
        if (MODE == G_DENSE_DGRAD) return pb[(long long)k * g.N + n];        // W(i=n,o=k) = Wt[n + k*nIn]
        if (MODE == G_DENSE_WGRAD) return pb[(long long)k * g.N + n];        // x[b=k, i=n]
        return pb[(long long)k * g.N + n];                                     // conv: [k][n] row-major
    }
};

// AccT = float: fp32 FMA accumulation (default).  AccT = double: exact products, double accumulation, one rounding —
// order-independent and therefore bit-compatible with the oracle (N4J_FLAG_EXACT_CONTRACTIONS, parity mode).
// TM: rows of the block tile / 16 (4: 64 x 64 tile, 4 x 4 outputs per thread; 3: 48 x 64 — parity mode is bound by the FP64 pipe, and the 98
// tiles of 64 rows of the MNIST convolutions leave a third of the SMs idle where 131 tiles of 48 rows still are one wave; 1: 16 x 64,
// measured slower: its shared-memory loads become the bound).
template <int MODE, typename AccT, int TM = 4, int BK = 16>
__global__ void __launch_bounds__(256) k_gemm_tiled(GemmArgs g) {
    pdl_begin();
    constexpr int BM = 16 * TM, BN = 64;
    // the tiles are held in the accumulation type: in parity mode (double) every element is converted ONCE when it is stored, not once
    // per product (F2F.F64.F32 is a slow instruction: 128 conversions per thread and K tile next to 256 DFMA)
    __shared__ __align__(16) AccT As[BK][BM + 4];
    __shared__ __align__(16) AccT Bs[BK][BN + 4];
    const float* __restrict__ pa = resolve(g.a);
    const float* __restrict__ pb = g.w ? g.w : resolve(g.b);
    float* __restrict__ pc = resolve(g.c);
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    // K range of this split
    const int kchunk = (g.K + g.splits - 1) / g.splits;
    const int kbeg = blockIdx.z * kchunk, kend = min(g.K, kbeg + kchunk);
    // loader mapping: the A tile is BM x BK (EA elements per thread), the B tile BK x 64 (EB per thread).  The fastest-varying thread index
    // follows the operand's contiguous dimension: k for the dense forward / dgrad A operand (rows of x / delta), the conv forward / dgrad A
    // operand (channels of one tap) and the dense forward B operand (columns of the f-order weights), m / n elsewhere.  (The m-fastest
    // mapping on a k-contiguous operand made every lane of a warp touch its own cache line: measured 32 us for the 1024 x 256 x 256
    // forward GEMM of the time-series block against 13.5 us for its weight-gradient GEMM.)  BK = 32 (dense modes with a long K): these
    // GEMMs are a chain of K/BK dependent trips to L2 — one tile ahead in registers — so fewer, larger trips.
    constexpr bool A_KFAST = MODE == G_DENSE_FWD || MODE == G_DENSE_DGRAD || MODE == G_CONV_FWD || MODE == G_CONV_DGRAD;
    constexpr bool B_KFAST = MODE == G_DENSE_FWD;
    constexpr int EA = BM * BK / 256, EB = BN * BK / 256;
    constexpr int ARS = 256 / BK;        // rows (m or n) covered per pass of a k-fastest loader
    constexpr int AKS = 256 / BM, BKS = 256 / BN;   // k rows covered per pass of an m- / n-fastest loader
    static_assert(A_KFAST || 256 % BM == 0, "the m-fastest A loader needs a row tile that divides 256 (TM = 3 only with k-fastest A operands)");
    const int la_m = A_KFAST ? threadIdx.x / BK : threadIdx.x % BM, la_k = A_KFAST ? threadIdx.x % BK : threadIdx.x / BM;
    const int lb_n = B_KFAST ? threadIdx.x / BK : threadIdx.x % BN, lb_k = B_KFAST ? threadIdx.x % BK : threadIdx.x / BN;
    int hh[EA], ww[EA];      // conv: pixel coordinates of the rows this thread loads
#pragma unroll
    for (int j = 0; j < EA; ++j) { hh[j] = 0; ww[j] = 0; }
    if (MODE == G_CONV_FWD || MODE == G_CONV_DGRAD) {
#pragma unroll
        for (int j = 0; j < EA; ++j) {
            const int m = m0 + la_m + ARS * j;
            const int s = m % (g.H * g.W);
            hh[j] = s / g.W; ww[j] = s % g.W;
        }
    }
    AccT acc[TM][4];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = (AccT)0;

    // register prefetch: the loads of K tile i+1 are in flight while tile i is multiplied (the summation order is unchanged)
    float ra[EA], rb[EB];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int j = 0; j < EA; ++j) {
            const int k = A_KFAST ? k0 + la_k : k0 + la_k + AKS * j, m = A_KFAST ? m0 + la_m + ARS * j : m0 + la_m;
            ra[j] = (m < g.M && k < kend) ? Addr<MODE>::A(g, pa, m, k, hh[j], ww[j]) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < EB; ++j) {
            const int kb = B_KFAST ? k0 + lb_k : k0 + lb_k + BKS * j, n = B_KFAST ? n0 + lb_n + ARS * j : n0 + lb_n;
            rb[j] = (n < g.N && kb < kend) ? Addr<MODE>::B(g, pb, kb, n) : 0.f;
        }
    };
    fetch(kbeg);
    for (int k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
        for (int j = 0; j < EA; ++j) {
            if (A_KFAST) As[la_k][la_m + ARS * j] = (AccT)ra[j]; else As[la_k + AKS * j][la_m] = (AccT)ra[j];
        }
#pragma unroll
        for (int j = 0; j < EB; ++j) {
            if (B_KFAST) Bs[lb_k][lb_n + ARS * j] = (AccT)rb[j]; else Bs[lb_k + BKS * j][lb_n] = (AccT)rb[j];
        }
        __syncthreads();
        if (k0 + BK < kend) fetch(k0 + BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            AccT a[TM], b[4];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[k][ty * TM + i];
#pragma unroll
            for (int i = 0; i < 4; ++i) b[i] = Bs[k][tx * 4 + i];      // 128-bit shared loads (rows are 16-byte aligned)
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (sizeof(AccT) == 8) acc[i][j] += a[i] * b[j];
                    else acc[i][j] = (AccT)fmaf((float)a[i], (float)b[j], (float)acc[i][j]);
                }
        }
        __syncthreads();
    }
    float* __restrict__ out = pc + (g.splits > 1 ? (long long)blockIdx.z * g.M * g.N : 0);
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + ty * TM + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= g.N) continue;
            float v;
            if (g.splits == 1) {
                if (sizeof(AccT) == 8) v = (float)(acc[i][j] + (AccT)(g.bias ? g.bias[n] : 0.f));
                else v = g.bias ? (float)acc[i][j] + g.bias[n] : (float)acc[i][j];
                v = act_fwd(g.act, v);
            } else {
                v = (float)acc[i][j];
            }
            out[(long long)m * g.N + n] = g.negate ? -v : v;
        }
    }
}

// Parity-mode weight gradient of the 3x3 convolution: dW[tap][ci][co] = sum_p X[p + shift(tap)][ci] * G[p][co], exact products accumulated in
// double in ASCENDING PIXEL ORDER per output (the order of k_gemm_tiled<G_CONV_WGRAD, double> and of the oracle: the result is bitwise the
// same), but on 16 x 16 output tiles with 2 x 2 outputs per thread: 9*Ci*Co/256 blocks (144 for 64 channels) instead of the nine 64 x 64
// tiles of the generic kernel, whose 6272-long chains on nine SMs were 70 % of the parity-mode step.  K cannot be split: a double sum is
// only reproducible inside one sequential chain.
constexpr int WGX_BK = 32;
__global__ void __launch_bounds__(256) k_conv_wgrad_exact(GemmArgs g) {
    pdl_begin();
    __shared__ double As[WGX_BK][16];
    __shared__ double Bs[WGX_BK][16];
    __shared__ unsigned short s_mask[1024];         // per pixel of one image: which of the nine taps stay inside the image
    const float* __restrict__ pa = resolve(g.a);
    const float* __restrict__ pb = resolve(g.b);
    float* __restrict__ pc = resolve(g.c);
    const int HW = g.H * g.W;
    for (int s = threadIdx.x; s < HW; s += 256) {
        const int h = s / g.W, w = s - h * g.W;
        unsigned int mk = 0;
        for (int t = 0; t < 9; ++t) {
            const int h2 = h + t / 3 - 1, w2 = w + t % 3 - 1;
            if ((unsigned)h2 < (unsigned)g.H && (unsigned)w2 < (unsigned)g.W) mk |= 1u << t;
        }
        s_mask[s] = (unsigned short)mk;
    }
    const int m0 = blockIdx.y * 16, n0 = blockIdx.x * 16;
    // ONE output per thread (tx: output channel, ty: tap * Ci + ci): the machine has only 9*Ci*Co independent chains to work on, and a
    // double-precision FMA re-issues slowly from the same warp (measured: four chains per thread on two warps took 126 cycles per pixel) —
    // sixteen warps with one chain each use all four schedulers
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    // loader: column (m or n) = tid % 16, pixel = tid / 16 + 16 j
    const int lc = tx, lk = ty;
    const int lm = m0 + lc, ln = n0 + lc;
    const bool m_ok = lm < g.M, n_ok = ln < g.N;
    const int tap = m_ok ? lm / g.Ci : 0, ci = m_ok ? lm - tap * g.Ci : 0;
    const long long shift = (long long)((tap / 3 - 1) * g.W + (tap % 3 - 1)) * g.Ci + ci;
    __syncthreads();
    double acc = 0.0;
    float ra[2], rb[2];
    auto fetch = [&](int k0) {
        int sp = (k0 + lk) % HW;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int k = k0 + lk + 16 * j;
            const bool in = k < g.K;
            ra[j] = (in && m_ok && ((s_mask[sp] >> tap) & 1)) ? pa[(long long)k * g.Ci + shift] : 0.f;
            rb[j] = (in && n_ok) ? pb[(long long)k * g.N + ln] : 0.f;
            sp += 16;
            while (sp >= HW) sp -= HW;
        }
    };
    fetch(0);
    for (int k0 = 0; k0 < g.K; k0 += WGX_BK) {
#pragma unroll
        for (int j = 0; j < 2; ++j) { As[lk + 16 * j][lc] = (double)ra[j]; Bs[lk + 16 * j][lc] = (double)rb[j]; }
        __syncthreads();
        if (k0 + WGX_BK < g.K) fetch(k0 + WGX_BK);
#pragma unroll
        for (int k = 0; k < WGX_BK; ++k) acc += As[k][ty] * Bs[k][tx];
        __syncthreads();
    }
    const int m = m0 + ty, n = n0 + tx;
    if (m < g.M && n < g.N) pc[(long long)m * g.N + n] = (float)acc;
}

// deterministic split-K reduction: dst[i] = sum_z part[z][i]
__global__ void __launch_bounds__(256) k_splitk_reduce(const float* __restrict__ part, Ref dstref, long long n, int splits) {
    pdl_begin();
    float* __restrict__ dst = resolve(dstref);
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        float s = part[i];
        for (int z = 1; z < splits; ++z) s += part[(long long)z * n + i];
        dst[i] = s;
    }
}

}  // namespace n4j
