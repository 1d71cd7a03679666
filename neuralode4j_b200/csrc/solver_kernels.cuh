// solver_kernels.cuh — the bandwidth-bound half of the hot path: DP5(4) stage combination, error norm,
// device-side step controller, initial-step heuristic and dense output.  Each is ONE vectorised (128-bit)
// pass over the concatenated minibatch state; reductions are warp-shuffle + block reduce in double with a
// last-block finalise that also runs the controller.
//
// Replaces (reference file:line, all under /root/reference/src/main/java/ode/solve):
//   impl/util/FirstOrderEquationWithState.java:42-51   State.step          -> k_stage_combine
//   impl/DormandPrince54Solver.java:78-93              DormandPrince54Mse  -> k_error_ctrl
//   impl/AdaptiveRungeKuttaSolver.java:76-103,147-181  accept/reject loop  -> k_error_ctrl (last block)
//   impl/util/AdaptiveRungeKuttaStepPolicy.java:90-153 initializeStep/step -> k_init_sums, k_init_final
//   impl/util/InterpolatingStepListener.java:59-150 + Interpolation.java:33-84 -> k_dense_output
#pragma once

#include "common.cuh"
#include "layer_kernels.cuh"

namespace n4j {

// DormandPrince54Solver.java:27-48 (tableau), rounded to fp32 like ND4J's default data type does.
__constant__ float c_A[6][6] = {
    {(float)(1.0 / 5.0), 0, 0, 0, 0, 0},
    {(float)(3.0 / 40.0), (float)(9.0 / 40.0), 0, 0, 0, 0},
    {(float)(44.0 / 45.0), (float)(-56.0 / 15.0), (float)(32.0 / 9.0), 0, 0, 0},
    {(float)(19372.0 / 6561.0), (float)(-25360.0 / 2187.0), (float)(64448.0 / 6561.0), (float)(-212.0 / 729.0), 0, 0},
    {(float)(9017.0 / 3168.0), (float)(-355.0 / 33.0), (float)(46732.0 / 5247.0), (float)(49.0 / 176.0), (float)(-5103.0 / 18656.0), 0},
    {(float)(35.0 / 384.0), 0.f, (float)(500.0 / 1113.0), (float)(125.0 / 192.0), (float)(-2187.0 / 6784.0), (float)(11.0 / 84.0)}};
__constant__ float c_E[7] = {(float)(71.0 / 57600.0), 0.f, (float)(-71.0 / 16695.0), (float)(71.0 / 1920.0),
                             (float)(-17253.0 / 339200.0), (float)(22.0 / 525.0), (float)(-1.0 / 40.0)};
__constant__ float c_C[6] = {(float)(1.0 / 5.0), (float)(3.0 / 10.0), (float)(4.0 / 5.0), (float)(8.0 / 9.0), 1.f, 1.f};
__constant__ float c_CMID[7] = {(float)(6025192743.0 / 30085553152.0 / 2.0), 0.f, (float)(51252292925.0 / 65400821598.0 / 2.0),
                                (float)(-2691868925.0 / 45128329728.0 / 2.0), (float)(187940372067.0 / 1594534317056.0 / 2.0),
                                (float)(-1776094331.0 / 19743644256.0 / 2.0), (float)(11237099.0 / 235043384.0 / 2.0)};

__device__ __forceinline__ float4 ld4(const float* p, long long i) { return reinterpret_cast<const float4*>(p)[i]; }
__device__ __forceinline__ void st4(float* p, long long i, float4 v) { reinterpret_cast<float4*>(p)[i] = v; }

// ---------------------------------------------------------------------------------------------------
// K1: yWorking = y + (a_S * h) . K[0:S]      S = 1..6;  S = 0 is the init-step probe y + h*K0.
// Zero coefficients (a_62) are skipped at compile time.  Algorithmic traffic: (nnz + 2) vectors.
// ---------------------------------------------------------------------------------------------------
// Optional fusion: when the inner graph starts with a 64-channel BatchNorm, the pass that writes yWorking also produces
// that layer's batch statistics (column sums of the NHWC state in double, last block finalises mean / invstd), so the
// RHS evaluation that follows starts directly with the BatchNorm apply.
struct BnFuse {
    int enabled;
    BnScratch sc;
    float eps;
    long long rows;   // B*H*W
    long long nz4;    // float4 count of the z block of the state (the adjoint state has more blocks behind it)
};

template <int S>
__global__ void __launch_bounds__(256) k_stage_combine(const Ctrl* __restrict__ c, float* __restrict__ Ybase,
                                                        const float* __restrict__ Kbase, long long stride, long long i_begin, long long n4,
                                                        BnFuse bf) {
    pdl_begin();
    const float h = c->h;
    const int yc = c->y_cur;
    const float* __restrict__ y = Ybase + (long long)yc * stride;
    float* __restrict__ yw = Ybase + (long long)(yc ^ 1) * stride;
    constexpr int NS = S == 0 ? 1 : S;
    float coef[NS];
    const float* __restrict__ kr[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) {
        coef[j] = S == 0 ? h : __fmul_rn(c_A[S - 1][j], h);   // stepCoeffPerStage.mul(step)
        kr[j] = Kbase + (long long)c->krow[j] * stride;
    }
    double s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = i_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += step) {   // float4 range [i_begin, n4)
        const float4 yv = ld4(y, i);
        float4 k0 = ld4(kr[0], i);
        float4 acc = make_float4(__fmul_rn(coef[0], k0.x), __fmul_rn(coef[0], k0.y), __fmul_rn(coef[0], k0.z), __fmul_rn(coef[0], k0.w));
#pragma unroll
        for (int j = 1; j < NS; ++j) {
            if (S == 6 && j == 1) continue;   // a_62 == 0
            const float4 kj = ld4(kr[j], i);
            acc.x = __fmaf_rn(coef[j], kj.x, acc.x);
            acc.y = __fmaf_rn(coef[j], kj.y, acc.y);
            acc.z = __fmaf_rn(coef[j], kj.z, acc.z);
            acc.w = __fmaf_rn(coef[j], kj.w, acc.w);
        }
        const float4 o = make_float4(__fadd_rn(yv.x, acc.x), __fadd_rn(yv.y, acc.y), __fadd_rn(yv.z, acc.z), __fadd_rn(yv.w, acc.w));
        st4(yw, i, o);
        if (bf.enabled && i < bf.nz4) {      // 64 channels = 16 float4 per row; this thread always sees chunk (threadIdx.x & 15)
            s1[0] += (double)o.x; s2[0] += (double)o.x * (double)o.x;
            s1[1] += (double)o.y; s2[1] += (double)o.y * (double)o.y;
            s1[2] += (double)o.z; s2[2] += (double)o.z * (double)o.z;
            s1[3] += (double)o.w; s2[3] += (double)o.w * (double)o.w;
        }
    }
    if (bf.enabled) {
        __shared__ double sm[2][4][256];
        __shared__ int s_last;
#pragma unroll
        for (int k = 0; k < 4; ++k) { sm[0][k][threadIdx.x] = s1[k]; sm[1][k][threadIdx.x] = s2[k]; }
        __syncthreads();
        if (threadIdx.x < 16) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                double a = 0, b = 0;
                for (int r = 0; r < 16; ++r) { a += sm[0][k][r * 16 + threadIdx.x]; b += sm[1][k][r * 16 + threadIdx.x]; }
                const int ch = threadIdx.x * 4 + k;
                atomicAdd((bf.sc.deferred ? bf.sc.accf : bf.sc.acc) + ch * ACC_STRIDE, a);
                atomicAdd((bf.sc.deferred ? bf.sc.accf : bf.sc.acc) + (64 + ch) * ACC_STRIDE, b);
            }
            __threadfence();
        }
        if (bf.sc.deferred) { bn_rearm(bf.sc.accf_other, threadIdx.x); bn_bump(bf.sc.cm ? bf.sc.seqf : nullptr, threadIdx.x); return; }   // the consumers finalise (layer_kernels.cuh: BnScratch)
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned int tk = atomicAdd(bf.sc.ticket, 1u);
            s_last = (tk == gridDim.x - 1);
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            double* tot = &sm[0][0][0];
            bn_collect(bf.sc, 64, tot);
            const double Mt = bf.sc.cm ? (double)bf.sc.rows_total : (double)bf.rows;
            if (threadIdx.x < 64) {
                const int ch = threadIdx.x;
                const double mean = tot[ch] / Mt;
                double var = tot[64 + ch] / Mt - mean * mean;
                if (var < 0) var = 0;
                bf.sc.mean[ch] = (float)mean;
                bf.sc.invstd[ch] = (float)(1.0 / sqrt(var + (double)bf.eps));
            }
            if (threadIdx.x == 0) *bf.sc.ticket = 0u;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Stage combination that also FINISHES the evaluation in front of it (64-channel NHWC blocks, deferred statistics): the last
// elementwise kernel of that evaluation only feeds this pass, so it runs inside it and the K row is written from here.
//   MODE 1 (forward solves): K[S-1].z = act(BatchNorm_last(x))                      (was k_bn_apply64)
//   MODE 2 (adjoint solves): K[S-1].a = BatchNorm_first backward dx                 (was k_bn_bwd_dx64; K[S-1].z is already there)
// The arithmetic is the unfused kernels' expression for expression (the K value is used from the register instead of being
// re-loaded), so results are bit-identical.  MODE 2 reads the BatchNorm input x from the z block of yWorking, which this very
// pass overwrites: one thread owns float4 i of the z block AND of the a block, and reads x before it writes.
// ---------------------------------------------------------------------------------------------------
struct FuseLast {
    int mode;                 // 0: nothing left over
    int act;
    Ref x;                    // MODE 1: input of the last BatchNorm
    Ref g, y;                 // MODE 2: gradient w.r.t. the first BatchNorm's output, that output (compact)
    float gscale;
    const float* gamma;
    const float* beta;
    BnScratch bn;             // statistics sets of the evaluation being finished
    long long a4;             // MODE 2: float4 offset of the a block inside a state vector
    long long dgamma_off;     // MODE 2: element offset of dgamma inside the K row (dbeta follows 64 floats later)
};

template <int S, int MODE>
__global__ void __launch_bounds__(256) k_stage_combine_last(const Ctrl* __restrict__ c, float* __restrict__ Ybase, float* __restrict__ Kbase,
                                                             long long stride, BnFuse bf, FuseLast fl) {
    static_assert(S >= 2, "the first stage reads the FSAL row");
    pdl_begin();
    const float h = c->h;
    const int yc = c->y_cur;
    const float* __restrict__ y = Ybase + (long long)yc * stride;
    float* __restrict__ yw = Ybase + (long long)(yc ^ 1) * stride;
    float coef[S];
    const float* __restrict__ kr[S];
#pragma unroll
    for (int j = 0; j < S; ++j) {
        coef[j] = __fmul_rn(c_A[S - 1][j], h);
        kr[j] = Kbase + (long long)c->krow[j] * stride;
    }
    float* __restrict__ klast = Kbase + (long long)c->krow[S - 1] * stride;
    const int c0 = (threadIdx.x & 15) * 4;      // grid stride is a multiple of 16 float4: a thread stays on one 4-channel chunk
    __shared__ float s_st[4][64];               // mean, invstd, mean(dy), mean(dy*xhat) of the BatchNorm being finished
    const float* __restrict__ xin = MODE == 1 ? resolve(fl.x) : yw;
    const float* __restrict__ gin = MODE == 2 ? resolve(fl.g) : nullptr;
    const float* __restrict__ yin = MODE == 2 ? resolve(fl.y) : nullptr;
    double s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
    const long long step = (long long)gridDim.x * blockDim.x;
    const long long i_first = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    // z block of the stage state (and the next evaluation's first-BatchNorm sums); kl: the fresh K value when this very kernel produced it (MODE 1)
    auto z_part = [&](long long i, const float* kl) {
        const float4 yv0 = ld4(y, i);
        const float4 k0 = ld4(kr[0], i);
        float4 acc = make_float4(__fmul_rn(coef[0], k0.x), __fmul_rn(coef[0], k0.y), __fmul_rn(coef[0], k0.z), __fmul_rn(coef[0], k0.w));
#pragma unroll
        for (int j = 1; j < S; ++j) {
            if (S == 6 && j == 1) continue;   // a_62 == 0
            const float4 kj = (MODE == 1 && j == S - 1) ? make_float4(kl[0], kl[1], kl[2], kl[3]) : ld4(kr[j], i);
            acc.x = __fmaf_rn(coef[j], kj.x, acc.x);
            acc.y = __fmaf_rn(coef[j], kj.y, acc.y);
            acc.z = __fmaf_rn(coef[j], kj.z, acc.z);
            acc.w = __fmaf_rn(coef[j], kj.w, acc.w);
        }
        const float4 o = make_float4(__fadd_rn(yv0.x, acc.x), __fadd_rn(yv0.y, acc.y), __fadd_rn(yv0.z, acc.z), __fadd_rn(yv0.w, acc.w));
        st4(yw, i, o);
        s1[0] += (double)o.x; s2[0] += (double)o.x * (double)o.x;
        s1[1] += (double)o.y; s2[1] += (double)o.y * (double)o.y;
        s1[2] += (double)o.z; s2[2] += (double)o.z * (double)o.z;
        s1[3] += (double)o.w; s2[3] += (double)o.w * (double)o.w;
    };
    // MODE 2: the z block does not depend on the statistics this kernel has to wait for (the backward sums the dgrad launch in front of it
    // produced — on the sharded path also the peers' share of them, one NVLink flight away): a thread's first ZIT items are combined BEFORE
    // the finalisation, with the BatchNorm input x — which lives in the z block of yWorking and is overwritten by that very store — kept in
    // registers for the a block afterwards.  The sums are pushed to the peers first of all.  Same expressions, same per-thread order of the
    // statistics accumulation: bit-identical results.
    constexpr int ZIT = 2;
    float4 xs[ZIT];
    if (MODE == 2) {
        if (fl.bn.deferred && fl.bn.cm && bn_first_block() && threadIdx.x < 64) {
            const unsigned int tagb = *((volatile unsigned int*)fl.bn.seqb);
            bn_push2(fl.bn.cm, fl.bn.mbb, tagb, threadIdx.x, fl.bn.accb[threadIdx.x * ACC_STRIDE], fl.bn.accb[(64 + threadIdx.x) * ACC_STRIDE]);
        }
#pragma unroll
        for (int it = 0; it < ZIT; ++it) {
            const long long i = i_first + it * step;
            if (i < bf.nz4) { xs[it] = ld4(xin, i); z_part(i, nullptr); }
        }
    }
    if (threadIdx.x < 64) {
        const int ch = threadIdx.x;
        bn_fwd_final(fl.bn, ch, s_st[0][ch], s_st[1][ch]);
        if (MODE == 2) {
            float dg, db;
            bn_bwd_final(fl.bn, ch, s_st[2][ch], s_st[3][ch], dg, db);
            if (blockIdx.x == 0) { klast[fl.dgamma_off + ch] = dg; klast[fl.dgamma_off + 64 + ch] = db; }
        }
    }
    __syncthreads();
    float mean[4], inv[4], ga[4], be[4], mdy[4], mdyxh[4], kk[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        mean[k] = s_st[0][c0 + k]; inv[k] = s_st[1][c0 + k];
        ga[k] = fl.gamma[c0 + k]; be[k] = MODE == 1 ? fl.beta[c0 + k] : 0.f;
        mdy[k] = MODE == 2 ? s_st[2][c0 + k] : 0.f; mdyxh[k] = MODE == 2 ? s_st[3][c0 + k] : 0.f;
        kk[k] = __fmul_rn(ga[k], inv[k]);
    }
    // MODE 2: K value of the a block (the first BatchNorm's backward dx) and the a block of the stage state
    auto a_part = [&](long long i, const float4 xv) {
        const float xx[4] = {xv.x, xv.y, xv.z, xv.w};
        const float4 gv = ld4(gin, i), yv = ld4(yin, i);
        const float gg[4] = {gv.x, gv.y, gv.z, gv.w}, yy[4] = {yv.x, yv.y, yv.z, yv.w};
        float kl[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float dy = act_bwd(fl.act, yy[k], __fmul_rn(fl.gscale, gg[k]));
            const float xh = __fmul_rn(__fsub_rn(xx[k], mean[k]), inv[k]);
            kl[k] = __fmul_rn(kk[k], __fsub_rn(__fsub_rn(dy, mdy[k]), __fmul_rn(xh, mdyxh[k])));
        }
        const long long ia = fl.a4 + i;
        st4(klast, ia, make_float4(kl[0], kl[1], kl[2], kl[3]));
        const float4 ya = ld4(y, ia);
        const float4 k0 = ld4(kr[0], ia);
        float4 acc = make_float4(__fmul_rn(coef[0], k0.x), __fmul_rn(coef[0], k0.y), __fmul_rn(coef[0], k0.z), __fmul_rn(coef[0], k0.w));
#pragma unroll
        for (int j = 1; j < S; ++j) {
            if (S == 6 && j == 1) continue;
            const float4 kj = j == S - 1 ? make_float4(kl[0], kl[1], kl[2], kl[3]) : ld4(kr[j], ia);
            acc.x = __fmaf_rn(coef[j], kj.x, acc.x);
            acc.y = __fmaf_rn(coef[j], kj.y, acc.y);
            acc.z = __fmaf_rn(coef[j], kj.z, acc.z);
            acc.w = __fmaf_rn(coef[j], kj.w, acc.w);
        }
        st4(yw, ia, make_float4(__fadd_rn(ya.x, acc.x), __fadd_rn(ya.y, acc.y), __fadd_rn(ya.z, acc.z), __fadd_rn(ya.w, acc.w)));
    };
    if (MODE == 1) {
        for (long long i = i_first; i < bf.nz4; i += step) {
            const float4 xv = ld4(xin, i);
            const float xx[4] = {xv.x, xv.y, xv.z, xv.w};
            float kl[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) kl[k] = act_fwd(fl.act, __fadd_rn(__fmul_rn(ga[k], __fmul_rn(__fsub_rn(xx[k], mean[k]), inv[k])), be[k]));
            st4(klast, i, make_float4(kl[0], kl[1], kl[2], kl[3]));
            z_part(i, kl);
        }
    } else {
#pragma unroll
        for (int it = 0; it < ZIT; ++it) {
            const long long i = i_first + it * step;
            if (i < bf.nz4) a_part(i, xs[it]);
        }
        for (long long i = i_first + ZIT * step; i < bf.nz4; i += step) {      // states longer than ZIT items per thread: x, then z, then a
            const float4 xv = ld4(xin, i);
            z_part(i, nullptr);
            a_part(i, xv);
        }
    }
    // statistics of the first BatchNorm for the evaluation that follows: accumulate only (deferred statistics, BnScratch)
    __shared__ double sm[2][4][256];
#pragma unroll
    for (int k = 0; k < 4; ++k) { sm[0][k][threadIdx.x] = s1[k]; sm[1][k][threadIdx.x] = s2[k]; }
    __syncthreads();
    if (threadIdx.x < 16) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double a = 0, b = 0;
            for (int r = 0; r < 16; ++r) { a += sm[0][k][r * 16 + threadIdx.x]; b += sm[1][k][r * 16 + threadIdx.x]; }
            const int ch = threadIdx.x * 4 + k;
            atomicAdd(bf.sc.accf + ch * ACC_STRIDE, a);
            atomicAdd(bf.sc.accf + (64 + ch) * ACC_STRIDE, b);
        }
    }
    // MODE 2 still reads the set behind accf_other (the first BatchNorm's sums of the evaluation being finished) in every block:
    // the operand prologue of the next convolution re-arms it instead (ConvTcArgs::prearm)
    if (MODE == 1) bn_rearm(bf.sc.accf_other, threadIdx.x);
    bn_bump(bf.sc.cm ? bf.sc.seqf : nullptr, threadIdx.x);
}

// ---------------------------------------------------------------------------------------------------
// Scalar controller pieces (run by one thread)
// ---------------------------------------------------------------------------------------------------
// AdaptiveRungeKuttaStepPolicy.step :142-153
__device__ __forceinline__ float next_step(float h, float err, const SolverParams& p) {
    float fac = __fmul_rn((float)pow((double)err, -1.0 / (double)p.order), p.safety);
    fac = fminf(p.max_growth, fmaxf(p.min_reduction, fac));
    const float sign = h > 0.f ? 1.f : (h < 0.f ? -1.f : 0.f);
    float v = __fmul_rn(__fmul_rn(fac, h), sign);
    v = fminf(p.hmax, fmaxf(p.hmin, v));
    return __fmul_rn(v, sign);
}

// TimeLimitForwards/Backwards.isLastStep, AdaptiveRungeKuttaSolver.java:76-103: clamps c->h for the NEXT trial.
__device__ __forceinline__ void clamp_to_end(Ctrl* c) {
    const float tph = __fadd_rn(c->t, c->h);
    const double tl = (double)c->t_end;
    const bool hit = c->backward ? ((double)tph <= tl) : ((double)tph >= tl);
    if (hit) {
        c->h = (float)(tl - (double)c->t);
        c->is_last = 1;
    } else {
        c->is_last = 0;
    }
}

// ---------------------------------------------------------------------------------------------------
// K2+K3: error norm + accept/reject + new step, one pass.  Reads the 6 K rows with e_j != 0, y and yWorking
// (8 vectors).  Multi-GPU: `peer_sum` hook is applied by the caller before the controller (see engine).
// ---------------------------------------------------------------------------------------------------
struct DenseOutArgs {
    const float* wanted;   // requested times (device), may be null
    int n_wanted;
};

__device__ __forceinline__ void controller_after_error(Ctrl* c, double sum, const SolverParams& p, n4j_step_rec* log,
                                                       DenseOutArgs dout) {
    const float err = (float)sqrt(sum / (double)p.n_norm);
    const float h = c->h;
    const float t_old = c->t;
    c->err = err;
    c->nfe += 6;
    c->trials += 1;
    c->trials_this_solve += 1;
    const bool ok = (double)err < 1.0;                                  // AdaptiveRungeKuttaSolver.java:170
    c->accepted_now = ok ? 1 : 0;
    if (ok) {
        c->t_prev = t_old;
        c->h_used = h;
        c->t = __fadd_rn(t_old, h);                                     // update(): time.addi(timeOffset)
        c->y_cur ^= 1;                                                  // y.assign(yWorking): pointer flip
#pragma unroll
        for (int j = 0; j < 7; ++j) c->krow_prev[j] = c->krow[j];
        const int r0 = c->krow[0];                                      // shiftDerivative(): K0 <- K6
        c->krow[0] = c->krow[6];
        c->krow[6] = r0;
        c->accepted += 1;
        c->interp_alias = (c->steps_this_solve == 0) ? 1 : 0;
        c->steps_this_solve += 1;
        // InterpolatingStepListener.step :72-96 — which requested times fall strictly inside the step
        int first = -1, cnt = 0;
        if (dout.wanted) {
            const double lo = h > 0.f ? (double)t_old : (double)c->t;
            const double hi = h > 0.f ? (double)c->t : (double)t_old;
            for (int i = 0; i < dout.n_wanted; ++i) {
                const double w = (double)dout.wanted[i];
                if (w > lo && w < hi) { if (first < 0) first = i; ++cnt; }
            }
        }
        c->interp_first = first;
        c->interp_cnt = cnt;
    } else {
        c->rejected += 1;
        c->interp_cnt = 0;
    }
    if (log && c->log_n < p.log_cap) {
        n4j_step_rec r;
        r.t = c->t; r.h = h; r.err = err; r.accepted = ok ? 1 : 0; r.solve = (int)c->solves - 1;
        log[c->log_n] = r;
    }
    c->log_n += 1;
    int done = (c->is_last && ok) ? 1 : 0;                               // :161
    c->h = next_step(h, err, p);                                         // :164
    if (!done) clamp_to_end(c);                                          // :149 of the next iteration
    if (!(err == err) || !(c->h == c->h)) { c->status = N4J_ERR_ILLEGAL_STATE; done = 1; }   // NaN guard
    if (!done && c->trials_this_solve >= p.max_trials) { c->status = N4J_ERR_MAX_STEPS; done = 1; }
    // multi-GPU: a peer that did not show up within the spin limit leaves stale sums behind — stop the solve instead of stepping on them
    if (c->comm && *((volatile int*)&c->comm->status) != 0) { c->status = N4J_ERR_COMM; done = 1; }
    c->done = done;
}

__global__ void __launch_bounds__(256) k_error_ctrl(Ctrl* c, const float* __restrict__ Ybase, const float* __restrict__ Kbase,
                                                     long long stride, long long n4, SolverParams p, n4j_step_rec* log,
                                                     DenseOutArgs dout, cudaGraphConditionalHandle cond, int use_cond) {
    pdl_begin();
    const float h = c->h;
    const int yc = c->y_cur;
    const float* __restrict__ y0 = Ybase + (long long)yc * stride;
    const float* __restrict__ y1 = Ybase + (long long)(yc ^ 1) * stride;
    const float* __restrict__ k0 = Kbase + (long long)c->krow[0] * stride;
    const float* __restrict__ k2 = Kbase + (long long)c->krow[2] * stride;
    const float* __restrict__ k3 = Kbase + (long long)c->krow[3] * stride;
    const float* __restrict__ k4 = Kbase + (long long)c->krow[4] * stride;
    const float* __restrict__ k5 = Kbase + (long long)c->krow[5] * stride;
    const float* __restrict__ k6 = Kbase + (long long)c->krow[6] * stride;
    const float e0 = c_E[0], e2 = c_E[2], e3 = c_E[3], e4 = c_E[4], e5 = c_E[5], e6 = c_E[6];
    double s[1] = {0.0};
    const long long step = (long long)gridDim.x * blockDim.x;
    const long long n4_eff = p.count_rep ? n4 : min(n4, p.rep_begin4);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4_eff; i += step) {
        const float4 a0 = ld4(k0, i), a2 = ld4(k2, i), a3 = ld4(k3, i), a4 = ld4(k4, i), a5 = ld4(k5, i), a6 = ld4(k6, i);
        const float4 u = ld4(y0, i), v = ld4(y1, i);
#define N4J_ERR_LANE(f)                                                                     \
        {                                                                                   \
            float es = __fmul_rn(e0, a0.f);                                                 \
            es = __fmaf_rn(e2, a2.f, es); es = __fmaf_rn(e3, a3.f, es);                     \
            es = __fmaf_rn(e4, a4.f, es); es = __fmaf_rn(e5, a5.f, es);                     \
            es = __fmaf_rn(e6, a6.f, es);                                                   \
            const float ys = fmaxf(fabsf(u.f), fabsf(v.f));                                 \
            const float tol = __fadd_rn(__fmul_rn(ys, p.rtol), p.atol);                     \
            const float ratio = __fmul_rn(__fdiv_rn(es, tol), h);                           \
            s[0] += (double)ratio * (double)ratio;                                          \
        }
        N4J_ERR_LANE(x) N4J_ERR_LANE(y) N4J_ERR_LANE(z) N4J_ERR_LANE(w)
#undef N4J_ERR_LANE
    }
    if (publish_and_elect<1>(s, c->red_acc, c->ticket, c->comm)) {
        if (threadIdx.x == 0) {
            const double sum = s[0];
            controller_after_error(c, sum, p, log, dout);
            if (use_cond) cudaGraphSetConditional(cond, c->done ? 0u : 1u);
        }
    }
}

// Variant without the controller: DormandPrince54Mse on caller data (node4j_error_norm).
__global__ void __launch_bounds__(256) k_error_only(const float* __restrict__ K, const float* __restrict__ y0, const float* __restrict__ y1,
                                                     long long n, float h, float atol, float rtol, double* acc) {
    pdl_begin();
    double s[1] = {0.0};
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        float es = __fmul_rn(c_E[0], K[i]);
        es = __fmaf_rn(c_E[2], K[2 * n + i], es); es = __fmaf_rn(c_E[3], K[3 * n + i], es);
        es = __fmaf_rn(c_E[4], K[4 * n + i], es); es = __fmaf_rn(c_E[5], K[5 * n + i], es);
        es = __fmaf_rn(c_E[6], K[6 * n + i], es);
        const float ys = fmaxf(fabsf(y0[i]), fabsf(y1[i]));
        const float tol = __fadd_rn(__fmul_rn(ys, rtol), atol);
        const float ratio = __fmul_rn(__fdiv_rn(es, tol), h);
        s[0] += (double)ratio * (double)ratio;
    }
    block_sum<1>(s);
    if (threadIdx.x == 0) atomicAdd(acc, s[0]);
}

// ---------------------------------------------------------------------------------------------------
// K4: initial step (AdaptiveRungeKuttaStepPolicy.initializeStep :90-139), two reduction passes.
// ---------------------------------------------------------------------------------------------------
__global__ void k_solve_begin(Ctrl* c, const float* tpair) {
    pdl_begin();
    // integrate(): new FirstOrderEquationWithState(equation, t[0], yOut.assign(y0)) — state reset for one solve
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const int pair_index = c->seg_index;
        c->seg_index = pair_index + 1;
        const float t0 = tpair[pair_index * 2], t1 = tpair[pair_index * 2 + 1];
        c->t = t0; c->t_start = t0; c->t_end = t1;
        c->backward = !(t1 > t0);                   // t.argMax()==0
        c->done = 0; c->is_last = 0;   // status stays sticky across the segments of one call
        c->trials_this_solve = 0; c->steps_this_solve = 0;
        c->accepted_now = 0; c->interp_cnt = 0; c->interp_alias = 0;
        for (int j = 0; j < 7; ++j) { c->krow[j] = j; c->krow_prev[j] = j; }
    }
}

__global__ void __launch_bounds__(256) k_init_sums(Ctrl* c, const float* __restrict__ Ybase, const float* __restrict__ Kbase,
                                                    long long stride, long long n4, SolverParams p) {
    pdl_begin();
    const float* __restrict__ y = Ybase + (long long)c->y_cur * stride;
    const float* __restrict__ k0 = Kbase + (long long)c->krow[0] * stride;
    double s[2] = {0.0, 0.0};
    const long long step = (long long)gridDim.x * blockDim.x;
    const long long n4_eff = p.count_rep ? n4 : min(n4, p.rep_begin4);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4_eff; i += step) {
        const float4 u = ld4(y, i), d = ld4(k0, i);
#define N4J_LANE(f)                                                               \
        {                                                                         \
            const float scal = __fadd_rn(__fmul_rn(fabsf(u.f), p.rtol), p.atol);  \
            const float r1 = __fdiv_rn(u.f, scal), r2 = __fdiv_rn(d.f, scal);     \
            s[0] += (double)r1 * (double)r1;                                      \
            s[1] += (double)r2 * (double)r2;                                      \
        }
        N4J_LANE(x) N4J_LANE(y) N4J_LANE(z) N4J_LANE(w)
#undef N4J_LANE
    }
    if (publish_and_elect<2>(s, c->red_acc, c->ticket, c->comm)) {
        if (threadIdx.x == 0) {
            const float Yr = (float)s[0];
            const float Dr = (float)s[1];
            float h = ((double)Yr < 1.0e-10 || (double)Dr < 1.0e-10) ? 1e-6f : __fmul_rn(sqrtf(__fdiv_rn(Yr, Dr)), 0.01f);
            if (c->backward) h = -h;
            c->h = h;          // probe step for equation.step(ones(1), step)
            c->init_D = Dr;
        }
    }
}

__global__ void __launch_bounds__(256) k_init_final(Ctrl* c, const float* __restrict__ Ybase, const float* __restrict__ Kbase,
                                                     long long stride, long long n4, SolverParams p,
                                                     cudaGraphConditionalHandle cond, int use_cond) {
    pdl_begin();
    const float* __restrict__ y = Ybase + (long long)c->y_cur * stride;
    const float* __restrict__ k0 = Kbase + (long long)c->krow[0] * stride;
    const float* __restrict__ k1 = Kbase + (long long)c->krow[1] * stride;
    double s[1] = {0.0};
    const long long step = (long long)gridDim.x * blockDim.x;
    const long long n4_eff = p.count_rep ? n4 : min(n4, p.rep_begin4);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4_eff; i += step) {
        const float4 u = ld4(y, i), a = ld4(k0, i), b = ld4(k1, i);
#define N4J_LANE(f)                                                               \
        {                                                                         \
            const float scal = __fadd_rn(__fmul_rn(fabsf(u.f), p.rtol), p.atol);  \
            const float r = __fdiv_rn(__fsub_rn(b.f, a.f), scal);                 \
            s[0] += (double)r * (double)r;                                        \
        }
        N4J_LANE(x) N4J_LANE(y) N4J_LANE(z) N4J_LANE(w)
#undef N4J_LANE
    }
    if (publish_and_elect<1>(s, c->red_acc, c->ticket, c->comm)) {
        if (threadIdx.x == 0) {
            const float DDs = (float)s[0];
            float h = c->h;
            const float DD = __fdiv_rn(sqrtf(DDs), h);                       // signed (:118,122)
            const float m = fmaxf(sqrtf(c->init_D), DD);
            float h1;
            if ((double)m < 1e-15) h1 = fmaxf(1e-6f, __fmul_rn(fabsf(h), 0.001f));
            else h1 = (float)pow((double)__fdiv_rn(0.01f, m), 1.0 / (double)p.order);
            h = fminf(__fmul_rn(fabsf(h), 100.f), h1);
            h = fmaxf(h, __fmul_rn(fabsf(c->t_start), 1e-12f));
            h = fmaxf(p.hmin, h);
            h = fminf(p.hmax, h);
            if (c->backward) h = -h;
            c->h = h;
            c->nfe += 3;                // :131, policy :92 (redundant, not recomputed here), policy :112
            c->solves += 1;
            clamp_to_end(c);            // first trial's isLastStep
            int done = 0;
            if (!(h == h)) { c->status = N4J_ERR_ILLEGAL_STATE; done = 1; }
            c->done = done;
            if (use_cond) cudaGraphSetConditional(cond, done ? 0u : 1u);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// K5: dense output for the requested times bracketed by the step that was just accepted.
// out is time-first [n_wanted][n]; launched after k_error_ctrl in every trial of an interpolating solve and
// returns immediately when there is nothing to do.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_dense_output(const Ctrl* __restrict__ c, const float* __restrict__ Ybase,
                                                       const float* __restrict__ Kbase, long long stride, long long n,
                                                       const float* __restrict__ wanted, float* __restrict__ out, long long out_stride) {
    pdl_begin();
    if (!c->accepted_now || c->interp_cnt <= 0) return;
    const float h = c->h_used;
    const double t0 = (double)c->t_prev, t1 = (double)c->t;
    const int yc = c->y_cur;
    const float* __restrict__ y1 = Ybase + (long long)yc * stride;
    const float* __restrict__ y0 = c->interp_alias ? y1 : Ybase + (long long)(yc ^ 1) * stride;
    const float* __restrict__ kr[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) kr[j] = Kbase + (long long)c->krow_prev[j] * stride;
    const double dtd = (double)h;
    const float F0[5] = {(float)(-2 * dtd), (float)(2 * dtd), -8.f, -8.f, 16.f};
    const float F1[5] = {(float)(5 * dtd), (float)(-3 * dtd), 18.f, 14.f, -32.f};
    const float F2[5] = {(float)(-4 * dtd), (float)dtd, -11.f, -5.f, 16.f};
    const int first = c->interp_first, cnt = c->interp_cnt;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        float kv[7];
#pragma unroll
        for (int j = 0; j < 7; ++j) kv[j] = kr[j][i];
        float o = __fmul_rn(__fmul_rn(kv[0], c_CMID[0]), h);                  // scaledDotProduct :122-128
#pragma unroll
        for (int j = 1; j < 7; ++j) o = __fadd_rn(o, __fmul_rn(__fmul_rn(kv[j], c_CMID[j]), h));
        const float a = y0[i], b = y1[i];
        const float ymid = __fadd_rn(a, o);
        const float in[5] = {kv[0], kv[6], a, b, ymid};
        float cf[5];
        {
            float q = __fmul_rn(in[0], F0[0]);
#pragma unroll
            for (int j = 1; j < 5; ++j) q = __fadd_rn(q, __fmul_rn(in[j], F0[j]));
            cf[0] = q;
            q = __fmul_rn(in[0], F1[0]);
#pragma unroll
            for (int j = 1; j < 5; ++j) q = __fadd_rn(q, __fmul_rn(in[j], F1[j]));
            cf[1] = q;
            q = __fmul_rn(in[0], F2[0]);
#pragma unroll
            for (int j = 1; j < 5; ++j) q = __fadd_rn(q, __fmul_rn(in[j], F2[j]));
            cf[2] = q;
        }
        cf[3] = __fmul_rn(kv[0], h);
        cf[4] = a;
        for (int w = first; w < first + cnt; ++w) {
            const double x = ((double)wanted[w] - t0) / (t1 - t0);                // Interpolation.interpolate :73-84
            double xs[5]; xs[4] = 1.0;
#pragma unroll
            for (int q = 3; q >= 0; --q) xs[q] = xs[q + 1] * x;
            float r = __fmul_rn(cf[0], (float)xs[0]);
#pragma unroll
            for (int q = 1; q < 5; ++q) r = __fadd_rn(r, __fmul_rn(cf[q], (float)xs[q]));
            out[(long long)w * out_stride + i] = r;
        }
    }
}

// InterpolatingStepListener.begin :59-69 / done :142-150 edge cases: copy the state if a requested time coincides
// (1e-10) with the start / end time.
__global__ void __launch_bounds__(256) k_dense_edge(const Ctrl* __restrict__ c, const float* __restrict__ Ybase, long long stride,
                                                     long long n, const float* __restrict__ wanted, int which_wanted, int at_end,
                                                     float* __restrict__ out, long long out_stride) {
    pdl_begin();
    const double tref = at_end ? (double)c->t : (double)c->t_start;
    if (!(fabs(tref - (double)wanted[which_wanted]) < 1e-10)) return;
    const float* __restrict__ y = Ybase + (long long)c->y_cur * stride;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step)
        out[(long long)which_wanted * out_stride + i] = y[i];
}

// ---------------------------------------------------------------------------------------------------
// Small helpers around a solve
// ---------------------------------------------------------------------------------------------------
// dst <- Y[y_cur]  (the caller's yOut after integrate, AdaptiveRungeKuttaSolver.java:126)
__global__ void __launch_bounds__(256) k_copy_from_state(const Ctrl* __restrict__ c, const float* __restrict__ Ybase, long long stride,
                                                          long long off, float* __restrict__ dst, long long n) {
    pdl_begin();
    const float* __restrict__ y = Ybase + (long long)c->y_cur * stride + off;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) dst[i] = y[i];
}
// Y[y_cur][off:off+n] <- src (+ optional add)
__global__ void __launch_bounds__(256) k_copy_to_state(const Ctrl* __restrict__ c, float* __restrict__ Ybase, long long stride, long long off,
                                                        const float* __restrict__ src, const float* __restrict__ add, long long n) {
    pdl_begin();
    float* __restrict__ y = Ybase + (long long)c->y_cur * stride + off;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step)
        y[i] = add ? __fadd_rn(src[i], add[i]) : src[i];
}
__global__ void __launch_bounds__(256) k_fill(float* __restrict__ p, float v, long long n) {
    pdl_begin();
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) p[i] = v;
}

// <a, b> in double -> (float) out[idx]; used for CalcTimeGrad.calcTimeAdjointT1 :52-58
__global__ void __launch_bounds__(256) k_dot(const float* __restrict__ a, const float* __restrict__ b, long long n, double* acc) {
    pdl_begin();
    double s[1] = {0.0};
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) s[0] += (double)a[i] * (double)b[i];
    block_sum<1>(s);
    if (threadIdx.x == 0) atomicAdd(acc, s[0]);
}

// ---------------------------------------------------------------------------------------------------
// Multi-GPU: all-reduce (sum) of a parameter-gradient slot over the ranks — the "parameter-gradient allreduce" of the
// adjoint solve, done per RHS evaluation so that the theta block of the state (and its error-norm term) stays global.
// One-shot over NVLink peer memory: copy to the local exchange buffer, signal every peer, wait for all, sum the peers'
// buffers in rank order.  Grid <= 64 blocks so every block is resident while it waits.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_comm_allreduce_vec(CommDev* cm, Ref bufref, long long n) {
    pdl_begin();
    float* __restrict__ buf = resolve(bufref);
    __shared__ int s_last;
    const unsigned long long seq = cm->seq_x + 1;
    const int par = (int)(seq & 1), me = cm->rank, W = cm->world;
    float* xb = reinterpret_cast<float*>(cm->peer[me] + COMM_OFF_XBUF) + (long long)par * cm->xcap;
    const long long step = (long long)gridDim.x * blockDim.x;
    const long long n4 = (((size_t)buf & 15) == 0) ? (n >> 2) : 0;     // 128-bit path when the slot is 16-byte aligned; the rest is scalar
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += step)
        reinterpret_cast<float4*>(xb)[i] = reinterpret_cast<const float4*>(buf)[i];
    for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) xb[i] = buf[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int tk = atomicAdd(&cm->xticket, 1u);
        s_last = (tk == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence_system();
        if (threadIdx.x < W) {
            volatile unsigned long long* f = reinterpret_cast<unsigned long long*>(cm->peer[threadIdx.x] + COMM_OFF_XFLAGS) + par * COMM_MAX_WORLD + me;
            *f = seq;
        }
    }
    if (threadIdx.x < W) {
        volatile unsigned long long* mine = reinterpret_cast<unsigned long long*>(cm->peer[me] + COMM_OFF_XFLAGS) + par * COMM_MAX_WORLD + threadIdx.x;
        comm_wait_flag(mine, seq, cm);
    }
    __syncthreads();
    __threadfence_system();
    // pull: all peers' contributions for one float4 are requested before any is used (W independent 16-byte NVLink loads in flight)
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += step) {
        float4 v[COMM_MAX_WORLD];
#pragma unroll
        for (int q = 0; q < COMM_MAX_WORLD; ++q)
            if (q < W) v[q] = __ldcv(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(cm->peer[q] + COMM_OFF_XBUF) + (long long)par * cm->xcap) + i);
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int q = 0; q < COMM_MAX_WORLD; ++q)
            if (q < W) { t.x += v[q].x; t.y += v[q].y; t.z += v[q].z; t.w += v[q].w; }
        reinterpret_cast<float4*>(buf)[i] = t;
    }
    for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        float t = 0.f;
        for (int q = 0; q < W; ++q)
            t += *reinterpret_cast<volatile float*>(reinterpret_cast<float*>(cm->peer[q] + COMM_OFF_XBUF) + (long long)par * cm->xcap + i);
        buf[i] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int tk = atomicAdd(&cm->pad2_, 1u);      // second ticket: the last block to LEAVE bumps the sequence number
        if (tk == gridDim.x - 1) { cm->xticket = 0u; cm->pad2_ = 0u; __threadfence(); cm->seq_x = seq; }
    }
}

// The same all-reduce in the LL form (comm_ll_allreduce4): n % 4 == 0, n <= xll_cap, 16-byte aligned slot.  No block waits for
// another block of its own rank, so the grid is not limited to co-resident blocks.
__global__ void __launch_bounds__(256) k_comm_allreduce_vec_ll(CommDev* cm, Ref bufref, long long n) {
    pdl_begin();
    float4* __restrict__ buf = reinterpret_cast<float4*>(resolve(bufref));
    const unsigned long long seq = cm->seq_x + 1;
    const int par = (int)(seq & 1);
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (n >> 2); i += step)
        buf[i] = comm_ll_allreduce4(cm, (unsigned int)seq, par, i * 4, buf[i]);
    __syncthreads();
    comm_ll_leave(cm, seq);
}

}  // namespace n4j
