// conv_tc.cuh — 3x3 "same" convolution (64 -> 64 channels) as an implicit GEMM on the 5th-generation tensor cores.
//
// Replaces DL4J's ConvolutionLayer forward / backpropGradient reached from ForwardPass.java:81 and
// BackpropagateAdjoint.java:109 (cuDNN / im2col+GEMM in the reference).
//
//   D[pixel, co] = sum_{tap} A[pixel + shift(tap), ci] * W_tap[ci, co]          fp32-faithful via 3xTF32:
//   a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi   with hi = tf32(x), lo = x - hi, accumulated in fp32 in TMEM.
//
// The activation operand lives in global memory as a zero-padded, chunk-major image (tc_common.cuh):
//   img[hl][chunk 0..15][guard + b*(H+2)*(W+2) + ph*(W+2) + pw][4]      hl = 0 (hi) / 1 (lo)
// written by the producing BatchNorm-apply (or BatchNorm-backward) kernel, border pixels stay zero.  One CTA computes
// 128 consecutive padded pixels x 64 output channels: the tile plus a halo of W+3 rows is staged ONCE by 32 bulk copies and
// the nine taps are nine row-shifted shared-memory descriptors into it.  Weights stream per tap (hi+lo = 32 KB) through a
// 4-deep mbarrier ring.  Warp roles: warp 0 = bulk-copy producer (+TMEM alloc), warp 1 = MMA issuer, warps 2-9 = epilogue
// (TMEM -> registers -> compact NHWC rows, optional fused BatchNorm statistics for the next layer).
#pragma once

#include "common.cuh"
#include "layer_kernels.cuh"
#include "tc_common.cuh"

namespace n4j {

struct ConvGeom {
    int B, H, W;      // logical
    int PH, PW;       // padded: H+2, W+2
    int halo;         // PW + 1 rows on each side of a tile
    int guard;        // zero rows in front of / behind the image (>= halo, multiple of 4)
    long long mpad;   // B*PH*PW padded pixels
    long long rows;   // guard + round_up(mpad,128) + guard : rows per chunk of one hi/lo image
    int tiles;        // ceil(mpad / 128)
};

inline ConvGeom make_conv_geom(int B, int H, int W) {
    ConvGeom g;
    g.B = B; g.H = H; g.W = W; g.PH = H + 2; g.PW = W + 2;
    g.halo = g.PW + 1;
    g.guard = (g.halo + 3) & ~3;
    g.mpad = (long long)B * g.PH * g.PW;
    g.tiles = (int)((g.mpad + 127) / 128);
    g.rows = g.guard + (long long)g.tiles * 128 + g.guard;
    return g;
}
inline long long conv_image_floats(const ConvGeom& g) { return 2LL * 16 * g.rows * 4; }   // hi + lo

constexpr int CONV_TC_THREADS = 320;       // producer warp, MMA warp, eight epilogue warps (two per TMEM lane quadrant: 32 columns each)
constexpr int CONV_TC_THREADS_PRO = 512;   // fused operand prologue: ALL sixteen warps build the A tile (latency-bound code: more warps, less work each)
constexpr int conv_tc_threads(int pro) { return pro ? CONV_TC_THREADS_PRO : CONV_TC_THREADS; }
constexpr int CONV_TC_EPI_THREADS = 256;
constexpr int CONV_TC_STAGES = 4;
constexpr int CONV_TC_WSTAGE_BYTES = 2 * 16 * 64 * 16;   // hi+lo image of one tap: 32 KB

inline size_t conv_tc_smem_bytes(const ConvGeom& g) {
    const int rt = 128 + 2 * g.halo;
    return (size_t)2 * 16 * rt * 16 + (size_t)CONV_TC_STAGES * CONV_TC_WSTAGE_BYTES + 256;
}

// weights: DL4J [Co,Ci,3,3] -> per-tap shared-memory images with hi and lo stacked along N:
//   wimg[tap][chunk = k/4][hl*64 + n][k%4]      (one UMMA with N=128 then yields a_hi*b_hi | a_hi*b_lo side by side)
//   forward (mode 0): n = co, k = ci, tap as is              B_tap[ci,co] = W[co,ci,tap]
//   dgrad   (mode 1): n = ci, k = co, tap flipped (8 - tap)  B_tap'[co,ci] = W[co,ci,8-tap']
__global__ void __launch_bounds__(256) k_conv_w_to_tc(const float* __restrict__ w, float* __restrict__ wimg, int mode) {
    pdl_begin();
    const int n_el = 64 * 64 * 9;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_el; i += gridDim.x * blockDim.x) {
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

// small all-reduce by the first four epilogue warps of the conv kernel (named barrier 2 instead of __syncthreads)
struct EpilogueSyncFn { __device__ __forceinline__ void operator()() const { asm volatile("bar.sync 2, 128;" ::: "memory"); } };
__device__ __forceinline__ void comm_allreduce_epilogue(CommDev* cm, double* vals, int n, int et) {
    comm_allreduce_small(cm, vals, n, et, 128, EpilogueSyncFn{});
}

struct ConvTcArgs {
    const float* img;     // operand image (hi then lo)
    const float* wimg;    // weight images [9][2][4096]
    Ref out;              // compact NHWC [B*H*W][64]
    ConvGeom g;
    int act;              // activation of the conv layer (identity in the reference's blocks)
    // optional fused BatchNorm statistics of the OUTPUT (for a following BatchNorm layer): column sums in the epilogue,
    // last CTA finalises mean / invstd and re-arms the accumulators (same contract as k_bn_stats)
    int fuse_stats;
    BnScratch bn;
    float bn_eps;
    long long bn_rows;    // B*H*W
    // fused operand prologue (PRO != 0): all sixteen warps of a 512-thread CTA build the A tile in shared memory from compact NHWC tensors
    //   PRO = 1  forward : A = act(gamma * ((x - mean) * invstd) + beta)                      (BatchNorm apply of the layer in front)
    //   PRO = 2  backward: A = gamma*invstd * ((dy - mean(dy)) - xhat * mean(dy*xhat)),  dy = act'(y) * gscale * g
    //                      (BatchNorm backward of the layer BEHIND the conv whose dgrad this launch computes)
    Ref px, pg;           // x: BatchNorm input, g: gradient w.r.t. the BatchNorm output (PRO 2; its output y is recomputed from x)
    float pgscale;
    const float* pgamma;
    const float* pbeta;
    BnScratch pbn;        // statistics of that BatchNorm (mean, invstd, m_dy, m_dyxh)
    int pact;
    float* pcompact;      // optional: also store the produced tensor compactly (activation for the backward pass / gradient for wgrad)
    Ref pdgamma, pdbeta;  // PRO 2 with deferred statistics: where dgamma / dbeta of that BatchNorm go
    int prearm;           // PRO 1: CTA 0 re-arms pbn.accf_other (the stage combination in front still read that set: k_stage_combine_last MODE 2)
    // fused BatchNorm BACKWARD statistics in the epilogue (EPI == 2, dgrad launches): the output of this launch is the gradient
    // w.r.t. the output of the BatchNorm layer in front of the conv; sums of dy and dy*xhat over the batch, last CTA writes
    // dgamma / dbeta / mean(dy) / mean(dy*xhat) (same contract as k_bn_bwd_stats64).  `bn`, `bn_rows` describe that layer.
    Ref ex;               // its input, compact NHWC (its output's sign is recomputed from it)
    const float* egamma;
    const float* ebeta;
    int eact;             // its activation (ReLU or identity)
    Ref edgamma, edbeta;
    // SUB instantiation (the stem's stride-2 convolutions, outer_net_tc.cuh): the launch computes the stride-1 "same" convolution of the
    // full-resolution image and stores only the pixels a stride-`sub_s` convolution samples, h = oh * sub_s + sub_off (same for w),
    // as compact rows of a [B, sub_H, sub_W] tensor; `act` (ReLU) is applied there.
    int sub_s, sub_off, sub_H, sub_W;
    long long* dbg;       // optional (N4J_CONV_DBG): SM-clock timestamps of CTA 0's phases, see engine.cu: bench_piece
    int dbg_nowstream;    // timing experiment only (N4J_DBG_CONV_NOWSTREAM, wrong results): taps 4-8 reuse the first four weight stages, nothing streams during the MMA phase
};

// Finalisation of the deferred BatchNorm statistics for the operand prologue: one channel per thread of warps 0-1.  Deliberately
// NOT inlined (and fed by value): inlined, the compiler hoists the loop-invariant double-precision reciprocal / division sequences in front of the
// operand loads of ALL warps (measured: the loads were issued 2700 cycles after kernel entry instead of ~1000).
template <int PRO>
__device__ __noinline__ void conv_finalize_stats(double r0, double r1, double r2, double r3, long long rows, float eps, float (*s_fin)[64],
                                                  float* dgamma, float* dbeta, int fast, double inv_rows) {      // everything by value: no local copy of the kernel arguments
    const int c = threadIdx.x;
    if (fast) {      // default mode: the forms of bn_fwd_final / bn_bwd_final with fast_final (layer_kernels.cuh: BnScratch)
        const double m = r0 * inv_rows;
        double var = r1 * inv_rows - m * m;
        if (var < 0) var = 0;
        s_fin[0][c] = (float)m;
        s_fin[1][c] = (float)rsqrt(var + (double)eps);
        if (PRO == 2) {
            s_fin[2][c] = (float)(r2 * inv_rows);
            s_fin[3][c] = (float)(r3 * inv_rows);
            if (dgamma) { dgamma[c] = (float)r3; dbeta[c] = (float)r2; }
        }
        return;
    }
    const double M = (double)rows;
    const double m = r0 / M;
    double var = r1 / M - m * m;
    if (var < 0) var = 0;
    s_fin[0][c] = (float)m;
    s_fin[1][c] = (float)(1.0 / sqrt(var + (double)eps));
    if (PRO == 2) {
        s_fin[2][c] = (float)(r2 / M);
        s_fin[3][c] = (float)(r3 / M);
        if (dgamma) { dgamma[c] = (float)r3; dbeta[c] = (float)r2; }
    }
}

__device__ __forceinline__ void conv_stamp(const ConvTcArgs& a, int slot) {
    if (a.dbg && blockIdx.x == 0) a.dbg[slot] = clock64();
}
__device__ __forceinline__ void conv_stamp_max(const ConvTcArgs& a, int slot) {
    if (a.dbg && blockIdx.x == 0) atomicMax(reinterpret_cast<unsigned long long*>(a.dbg + slot), (unsigned long long)clock64());
}

// SH: sharded (multi-GPU) instantiation — the cross-rank statistics exchange of the deferred scheme is compiled in only there, so the
// single-GPU kernels carry none of its code or registers (measured: +0.3 us per launch when it was a run-time branch).
template <int PASSES, int EPI, int PRO, int SH = 0, int SUB = 0>
__global__ void __launch_bounds__(conv_tc_threads(PRO), 1) k_conv3x3_tc(ConvTcArgs a) {
    if (threadIdx.x == 0) conv_stamp(a, 0);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");     // dependents may be scheduled; they park in their own wait
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const ConvGeom& g = a.g;
    const int rt = 128 + 2 * g.halo;                     // staged rows per chunk
    const uint32_t a_chunk_bytes = (uint32_t)rt * 16;    // one chunk of one hi/lo image
    const uint32_t a_img_bytes = 16 * a_chunk_bytes;
    uint8_t* sA = smem_raw;                              // [2][16][rt][16B]
    uint8_t* sW = sA + 2 * a_img_bytes;                  // [STAGES][2][16][64][16B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sW + CONV_TC_STAGES * CONV_TC_WSTAGE_BYTES);
    uint64_t* w_full = bars;                             // [STAGES]
    uint64_t* w_empty = bars + CONV_TC_STAGES;           // [STAGES]
    uint64_t* a_full = bars + 2 * CONV_TC_STAGES;
    uint64_t* d_full = bars + 2 * CONV_TC_STAGES + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * CONV_TC_STAGES + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long m0 = (long long)blockIdx.x * 128;

    if (threadIdx.x == 0) {
        for (int s = 0; s < CONV_TC_STAGES; ++s) { tc::mbar_init(&w_full[s], 1); tc::mbar_init(&w_empty[s], 1); }
        tc::mbar_init(a_full, PRO == 0 ? 1 : CONV_TC_THREADS_PRO / 32);
        tc::mbar_init(d_full, 1);
        tc::fence_barrier_init();
    }
    if (warp == 0) tc::tmem_alloc<128>(tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    // barrier init and the TMEM allocation above overlap the tail of the producing kernel; global memory is touched from here on
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (threadIdx.x == 0) conv_stamp(a, 1);

    if constexpr (PRO != 0) {
        // ---- fused operand prologue, all sixteen warps: BatchNorm apply (+ReLU) or BatchNorm backward dx, TF32 hi/lo split,
        // zero padding.  The weight ring is already being filled (below) while this runs.  A warp owns blocks of 8 rows x 4
        // channel chunks: per quarter warp the shared-memory stores are 128 contiguous bytes (no bank conflicts) and the
        // global loads are 64 contiguous bytes per row.  All loads of a thread are issued before the first use (one L2
        // latency per thread, not one per row).
        if (threadIdx.x == 0) {
            for (int tap = 0; tap < CONV_TC_STAGES; ++tap) {
                tc::mbar_arrive_expect_tx(&w_full[tap], CONV_TC_WSTAGE_BYTES);
                tc::bulk_g2s(sW + tap * CONV_TC_WSTAGE_BYTES, a.wimg + (long long)tap * 8192, CONV_TC_WSTAGE_BYTES, &w_full[tap]);
            }
        }
        const int chunk = (warp & 3) * 4 + (lane >> 3), c0 = chunk * 4;
        const float* __restrict__ px = resolve(a.px);
        const float* __restrict__ pg = PRO == 2 ? resolve(a.pg) : nullptr;
        // only ReLU / identity are fused (the engine keeps the separate kernels for ELU / tanh): no transcendental code in this kernel
        const bool relu = a.pact == N4J_ACT_RELU;
        // deferred statistics: the producing kernel only accumulated.  64 threads finalise one channel each from the raw double sums
        // (double division / square root are long instruction sequences: once per CTA, not once per thread); the loads of the sums
        // are issued here, the arithmetic runs after the operand loads below are in flight.  CTA 0 publishes dgamma / dbeta, which
        // the last block of the statistics kernel used to do.
        __shared__ float s_fin[4][64];     // mean, invstd, mean(dy), mean(dy*xhat)
        double raw[4] = {0.0, 0.0, 0.0, 0.0};
        if (PRO == 1 && a.prearm && blockIdx.x == 0 && warp >= 2 && warp < 6) a.pbn.accf_other[(threadIdx.x - 64) * ACC_STRIDE] = 0.0;
        unsigned int tagf = 0u, tagb = 0u;
        if (a.pbn.deferred && warp < 2) {      // warp-uniform on purpose: a real branch, not predicated double-precision code in all warps
            raw[0] = a.pbn.accf[threadIdx.x * ACC_STRIDE]; raw[1] = a.pbn.accf[(64 + threadIdx.x) * ACC_STRIDE];
            if (PRO == 2) { raw[2] = a.pbn.accb[threadIdx.x * ACC_STRIDE]; raw[3] = a.pbn.accb[(64 + threadIdx.x) * ACC_STRIDE]; }
            if (SH && a.pbn.cm) {      // sharded: CTA 0 sends this rank's local sums on their way now; the peers' sums are collected right before the finalisation
                tagf = *((volatile unsigned int*)a.pbn.seqf);
                if (PRO == 2) tagb = *((volatile unsigned int*)a.pbn.seqb);
                if (blockIdx.x == 0) {
                    bn_push2(a.pbn.cm, a.pbn.mbf, tagf, threadIdx.x, raw[0], raw[1]);
                    if (PRO == 2) bn_push2(a.pbn.cm, a.pbn.mbb, tagb, threadIdx.x, raw[2], raw[3]);
                }
            }
        }
        float mean[4], inv[4], gam[4], bet[4], mdy[4], mdyxh[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            gam[k] = a.pgamma[c0 + k];
            bet[k] = a.pbeta[c0 + k];
            mdy[k] = 0.f; mdyxh[k] = 0.f;
            if (!a.pbn.deferred) {
                mean[k] = a.pbn.mean[c0 + k]; inv[k] = a.pbn.invstd[c0 + k];
                if (PRO == 2) { mdy[k] = a.pbn.m_dy[c0 + k]; mdyxh[k] = a.pbn.m_dyxh[c0 + k]; }
            }
        }
        const int per = g.PH * g.PW;
        const float inv_pw = 1.0f / (float)g.PW;
        constexpr int U = 5;                         // rows r = (warp >> 2) * 8 + (lane & 7) + 32 u: 160 >= 128 + 2 * halo for W <= 14
        const int rbase = (warp >> 2) * 8 + (lane & 7);
        for (int rr = rbase; rr < rt; rr += 32 * U) {
            int i4[U];
            float4 xv[U], gv[U];
            const int pr0 = (int)m0 - g.halo + rr;     // padded rows fit 31 bits (checked where the tensor path is enabled)
            // image index / position inside the padded image of the first row, advanced by 32 rows per item
            int bq = pr0 >= 0 ? (int)((unsigned)pr0 / (unsigned)per) : -1;
            int rem = pr0 >= 0 ? pr0 - bq * per : pr0;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int r = rr + 32 * u;
                i4[u] = -1;
                if (r < rt && rem >= 0 && bq < g.B) {
                    const int ph = __float2int_rz(((float)rem + 0.5f) * inv_pw), pw = rem - ph * g.PW;
                    if (ph >= 1 && ph <= g.H && pw >= 1 && pw <= g.W) i4[u] = ((bq * g.H + (ph - 1)) * g.W + (pw - 1)) * 16 + chunk;
                }
                if (i4[u] >= 0) {
                    xv[u] = reinterpret_cast<const float4*>(px)[i4[u]];
                    if (PRO == 2) gv[u] = reinterpret_cast<const float4*>(pg)[i4[u]];
                }
                rem += 32;
                if (rem >= 0 && bq < 0) bq = 0;
                while (rem >= per) { rem -= per; ++bq; }
            }
            if (threadIdx.x == 0) conv_stamp(a, 2);
            if (a.pbn.deferred && rr == rbase) {       // first (for W <= 14: only) pass; every thread of the CTA gets here
                if constexpr (SH != 0) {
                    if (a.pbn.cm) {
                        // all sixteen warps collect the peers' cells: thread = (channel, source rank), every rank polled at the same time
                        // (one L2 round trip whatever the number of ranks); warps 0-1 then add them to the local sums in rank order
                        // (the buffer lives in the A-tile region, which nobody writes before the statistics are final)
                        double (*s_peer)[COMM_MAX_WORLD][64][2] = reinterpret_cast<double (*)[COMM_MAX_WORLD][64][2]>(sA);
                        const int gc = threadIdx.x & 63, gq = threadIdx.x >> 6;
                        const int me = a.pbn.cm->rank, W = a.pbn.cm->world;
                        const unsigned int tf = *((volatile unsigned int*)a.pbn.seqf);
                        if (gq < W && gq != me) {
                            const unsigned long long* mbox = reinterpret_cast<const unsigned long long*>(a.pbn.cm->peer[me] + a.pbn.cm->stat_off);
                            bn_poll_peer(a.pbn.cm, mbox + a.pbn.mbf / 8, gq, tf, gc, s_peer[0][gq][gc][0], s_peer[0][gq][gc][1]);
                            if (PRO == 2) bn_poll_peer(a.pbn.cm, mbox + a.pbn.mbb / 8, gq, *((volatile unsigned int*)a.pbn.seqb), gc, s_peer[1][gq][gc][0], s_peer[1][gq][gc][1]);
                        }
                        __syncthreads();
                        if (warp < 2) {
                            double t[4] = {0.0, 0.0, 0.0, 0.0};
                            for (int q = 0; q < W; ++q) {
                                if (q == me) { t[0] += raw[0]; t[1] += raw[1]; t[2] += raw[2]; t[3] += raw[3]; continue; }
                                t[0] += s_peer[0][q][threadIdx.x][0]; t[1] += s_peer[0][q][threadIdx.x][1];
                                if (PRO == 2) { t[2] += s_peer[1][q][threadIdx.x][0]; t[3] += s_peer[1][q][threadIdx.x][1]; }
                            }
                            raw[0] = t[0]; raw[1] = t[1]; raw[2] = t[2]; raw[3] = t[3];
                        }
                    }
                }
                if (warp < 2)
                    conv_finalize_stats<PRO>(raw[0], raw[1], raw[2], raw[3], (SH && a.pbn.cm) ? a.pbn.rows_total : a.pbn.rows, a.pbn.eps, s_fin,
                                             (PRO == 2 && blockIdx.x == 0) ? resolve(a.pdgamma) : nullptr, (PRO == 2 && blockIdx.x == 0) ? resolve(a.pdbeta) : nullptr,
                                             a.pbn.fast_final, a.pbn.inv_rows);
                __syncthreads();
                if (threadIdx.x == 0) conv_stamp(a, 3);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    mean[k] = s_fin[0][c0 + k]; inv[k] = s_fin[1][c0 + k];
                    if (PRO == 2) { mdy[k] = s_fin[2][c0 + k]; mdyxh[k] = s_fin[3][c0 + k]; }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int r = rr + 32 * u;
                if (r < rt) {
                    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (i4[u] >= 0) {
                        const float xx[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w};
                        float oo[4];
                        if (PRO == 1) {
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                            {
                                const float v = __fadd_rn(__fmul_rn(gam[k], __fmul_rn(__fsub_rn(xx[k], mean[k]), inv[k])), bet[k]);
                                oo[k] = (relu && !(v > 0.f)) ? 0.f : v;
                            }
                        } else {
                            const float gg[4] = {gv[u].x, gv[u].y, gv[u].z, gv[u].w};
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float gs = __fmul_rn(a.pgscale, gg[k]);
                                const float xh = __fmul_rn(__fsub_rn(xx[k], mean[k]), inv[k]);
                                // the ReLU mask is recomputed from x with the forward pass's own expression (bit-identical to y > 0),
                                // which saves loading the layer output
                                const float yk = __fadd_rn(__fmul_rn(gam[k], xh), bet[k]);
                                const float dy = (relu && !(yk > 0.f)) ? 0.f : gs;
                                oo[k] = __fmul_rn(__fmul_rn(gam[k], inv[k]), __fsub_rn(__fsub_rn(dy, mdy[k]), __fmul_rn(xh, mdyxh[k])));
                            }
                        }
                        o = make_float4(oo[0], oo[1], oo[2], oo[3]);
                        if (a.pcompact && r >= g.halo && r < g.halo + 128) reinterpret_cast<float4*>(a.pcompact)[i4[u]] = o;   // rows this tile owns
                    }
                    float4 hi, lo;
                    tc::split_tf32(o.x, hi.x, lo.x); tc::split_tf32(o.y, hi.y, lo.y); tc::split_tf32(o.z, hi.z, lo.z); tc::split_tf32(o.w, hi.w, lo.w);
                    *reinterpret_cast<float4*>(sA + chunk * a_chunk_bytes + r * 16) = hi;
                    *reinterpret_cast<float4*>(sA + a_img_bytes + chunk * a_chunk_bytes + r * 16) = lo;
                }
            }
        }
        if (threadIdx.x == 0) conv_stamp(a, 4);
        tc::fence_proxy_async();          // generic-proxy writes above -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) { tc::mbar_arrive(a_full); conv_stamp_max(a, 5); }
    }

    if (warp == 0) {
        if (lane == 0) {
            // ---- producer: activation tile + halo once, then the nine weight taps through the ring ----
            if (PRO == 0) {
                tc::mbar_arrive_expect_tx(a_full, 2 * a_img_bytes);
                const long long row0 = g.guard + m0 - g.halo;
                for (int hl = 0; hl < 2; ++hl)
                    for (int c = 0; c < 16; ++c)
                        tc::bulk_g2s(sA + hl * a_img_bytes + c * a_chunk_bytes, a.img + (((long long)hl * 16 + c) * g.rows + row0) * 4,
                                     a_chunk_bytes, a_full);
            }
            for (int tap = PRO == 0 ? 0 : CONV_TC_STAGES; tap < (a.dbg_nowstream ? CONV_TC_STAGES : 9); ++tap) {   // with a prologue the first STAGES taps are already in flight
                const int s = tap % CONV_TC_STAGES;
                if (tap >= CONV_TC_STAGES) tc::mbar_wait(&w_empty[s], ((tap / CONV_TC_STAGES) - 1) & 1);
                tc::mbar_arrive_expect_tx(&w_full[s], CONV_TC_WSTAGE_BYTES);
                tc::bulk_g2s(sW + s * CONV_TC_WSTAGE_BYTES, a.wimg + (long long)tap * 8192, CONV_TC_WSTAGE_BYTES, &w_full[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ---- MMA issuer: per tap and k-step (K=8):  D[:,0:128] += A_hi * [B_hi | B_lo]   (one N=128 instruction)
            //                                             D[:,0:64]  += A_lo * B_hi            (N=64)
            // the epilogue adds the two 64-column halves.  The a_lo*b_lo term (2^-22 relative) is dropped.
            constexpr uint32_t idesc128 = tc::idesc_tf32(128, 128, 0, 0), idesc64 = tc::idesc_tf32(128, 64, 0, 0);
            const uint32_t sA_u = tc::smem_u32(sA), sW_u = tc::smem_u32(sW);
            tc::mbar_wait(a_full, 0);
            conv_stamp(a, 6);
            uint32_t acc = 0;
            for (int tap = 0; tap < 9; ++tap) {
                const int s = tap % CONV_TC_STAGES;
                if (!a.dbg_nowstream || tap < CONV_TC_STAGES) tc::mbar_wait(&w_full[s], (tap / CONV_TC_STAGES) & 1);
                tc::tc_fence_after();
                const int shift = (tap / 3 - 1) * g.PW + (tap % 3 - 1);
                const uint32_t a_row = (uint32_t)(g.halo + shift) * 16;
                const uint32_t w_base = sW_u + s * CONV_TC_WSTAGE_BYTES;
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    const uint32_t a_hi = sA_u + (2 * kk) * a_chunk_bytes + a_row;
                    const uint32_t a_lo = a_hi + a_img_bytes;
                    const uint32_t b = w_base + (2 * kk) * 2048;
                    const uint64_t dAh = tc::smem_desc(a_hi, a_chunk_bytes, 128), dAl = tc::smem_desc(a_lo, a_chunk_bytes, 128);
                    const uint64_t dB = tc::smem_desc(b, 2048, 128);
                    if (PASSES == 3) {
                        tc::mma_tf32(tmem, dAh, dB, idesc128, acc); acc = 1;
                        tc::mma_tf32(tmem, dAl, dB, idesc64, 1);
                    } else if (PASSES == 1) {
                        tc::mma_tf32(tmem, dAh, dB, idesc64, acc); acc = 1;
                    }
                }
                tc::mma_commit(&w_empty[s]);
            }
            tc::mma_commit(d_full);
            conv_stamp(a, 7);
        }
    } else if (warp < 10) {
        // ---- epilogue, eight warps: TMEM lane = tile row = one padded pixel; a warp reads its lane quadrant (warp & 3, fixed by the
        // hardware) and ONE 32-column half of the 64 output channels, so every quadrant is served by two warps ----
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;            // warps 2-5: channels 0-31, warps 6-9: channels 32-63
        const int row = q * 32 + lane;
        const long long m = m0 + row;
        bool valid = m < g.mpad;
        long long crow = 0;
        if (valid) {
            const int per = g.PH * g.PW;
            const int b = (int)(m / per);
            const int rem = (int)(m - (long long)b * per);
            const int ph = rem / g.PW, pw = rem - ph * g.PW;
            valid = ph >= 1 && ph <= g.H && pw >= 1 && pw <= g.W;
            crow = ((long long)b * g.H + (ph - 1)) * g.W + (pw - 1);
            if constexpr (SUB != 0) {      // keep only the pixels the strided convolution samples
                const int dh = ph - 1 - a.sub_off, dw = pw - 1 - a.sub_off;
                const int oh = dh / a.sub_s, ow = dw / a.sub_s;
                valid = valid && dh >= 0 && dw >= 0 && oh * a.sub_s == dh && ow * a.sub_s == dw && oh < a.sub_H && ow < a.sub_W;
                crow = ((long long)b * a.sub_H + oh) * a.sub_W + ow;
            }
        }
        float* __restrict__ out = resolve(a.out);
        float* sT = reinterpret_cast<float*>(sW);      // [128][65] transpose buffer(s); the weight ring is idle once d_full fires
        float* sT2 = sT + 128 * 65;
        double* sP = reinterpret_cast<double*>(sT2 + 128 * 65);   // [4][128] partial column sums
        const int et = threadIdx.x - 64;               // 0..255
        // EPI 2: while the tensor core works, fetch this pixel's half row of the BatchNorm input / output and keep xhat and the
        // activation mask in registers
        float xh[EPI == 2 ? 32 : 1];
        unsigned int mask = 0u;
        if constexpr (EPI == 2) {
            __shared__ float s_mi[256];      // mean, invstd, gamma, beta of the BatchNorm in front
            if (a.bn.deferred) {
                if (et < 64) bn_fwd_final<SH != 0, true>(a.bn, et, s_mi[et], s_mi[64 + et]);   // all peers requested at once (per-peer polling cost +2 us at eight ranks)
            } else {
                if (et < 64) s_mi[et] = a.bn.mean[et]; else if (et < 128) s_mi[et] = a.bn.invstd[et - 64];
            }
            if (et >= 128 && et < 192) s_mi[et] = a.egamma[et - 128]; else if (et >= 192) s_mi[et] = a.ebeta[et - 192];
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const bool relu = a.eact == N4J_ACT_RELU;
            mask = relu ? 0u : ~0u;
            if (valid) {
                const float4* __restrict__ xr = reinterpret_cast<const float4*>(resolve(a.ex) + crow * 64 + half * 32);
                const float* smean = s_mi + half * 32;
                const float* sinv = s_mi + 64 + half * 32;
                const float* sgam = s_mi + 128 + half * 32;
                const float* sbet = s_mi + 192 + half * 32;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 xv = xr[j];
                    const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int c = 4 * j + k;
                        xh[c] = __fmul_rn(__fsub_rn(xs[k], smean[c]), sinv[c]);
                        // ReLU mask from the forward expression y = gamma * xhat + beta (bit-identical to the stored output's sign)
                        if (relu && __fadd_rn(__fmul_rn(sgam[c], xh[c]), sbet[c]) > 0.f) mask |= 1u << c;
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) xh[j] = 0.f;
                mask = 0u;
            }
        }
        tc::mbar_wait(d_full, 0);
        if (threadIdx.x == 64) conv_stamp(a, 8);
        tc::tc_fence_after();
        {
            float v[32];
            tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + half * 32, v);
            if (PASSES == 3) {
                float u[32];
                tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + 64 + half * 32, u);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] += u[j];
            }
            if constexpr (SUB != 0) {
                if (a.act == N4J_ACT_RELU) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : 0.f;
                }
            }
            if (valid) {   // the ODE block's tensor path only takes identity-activation convolutions (the reference's blocks)
                float4* dst = reinterpret_cast<float4*>(out + crow * 64 + half * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
            if constexpr (EPI == 1) {
#pragma unroll
                for (int j = 0; j < 32; ++j) sT[row * 65 + half * 32 + j] = valid ? v[j] : 0.f;
            }
            if constexpr (EPI == 2) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    sT[row * 65 + half * 32 + j] = ((mask >> j) & 1u) ? v[j] : 0.f;     // dy = act'(y) * g
                    sT2[row * 65 + half * 32 + j] = xh[j];
                }
            }
        }
        if (threadIdx.x == 64) conv_stamp(a, 9);
        if constexpr (EPI != 0) {
            asm volatile("bar.sync 1, 256;" ::: "memory");      // the eight epilogue warps only
            {   // column sums: 256 threads = 64 columns x 4 row quarters, combined through shared memory (one atomic per address and CTA)
                const int c = et & 63, r0 = (et >> 6) * 32;
                double p1 = 0.0, p2 = 0.0;
                if constexpr (EPI == 1) {
                    for (int r = r0; r < r0 + 32; ++r) { const double x = (double)sT[r * 65 + c]; p1 += x; p2 += x * x; }
                } else {
                    for (int r = r0; r < r0 + 32; ++r) { const double dy = (double)sT[r * 65 + c]; p1 += dy; p2 += dy * (double)sT2[r * 65 + c]; }
                }
                sP[(et >> 6) * 128 + c] = p1;
                sP[(et >> 6) * 128 + 64 + c] = p2;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (et < 128) {
                const double tsum = (sP[et] + sP[128 + et]) + (sP[256 + et] + sP[384 + et]);     // et < 64: first sums, else second sums
                if (a.bn.deferred) {      // accumulate only; the consumers finalise, CTA 0 re-arms the next evaluation's set
                    atomicAdd((EPI == 1 ? a.bn.accf : a.bn.accb) + et * ACC_STRIDE, tsum);
                    bn_rearm(EPI == 1 ? a.bn.accf_other : a.bn.accb_other, et);
                    if (SH) bn_bump(a.bn.cm ? (EPI == 1 ? a.bn.seqf : a.bn.seqb) : nullptr, et);
                } else {                  // legacy / multi-GPU: the last CTA finalises (and exchanges the sums between the ranks)
                    atomicAdd(a.bn.acc + et * ACC_STRIDE, tsum);
                    __threadfence();
                    asm volatile("bar.sync 2, 128;" ::: "memory");
                    __shared__ int s_last;
                    if (et == 0) {
                        const unsigned int tk = atomicAdd(a.bn.ticket, 1u);
                        s_last = (tk == gridDim.x - 1);
                    }
                    asm volatile("bar.sync 2, 128;" ::: "memory");
                    if (s_last) {
                        __threadfence();
                        double* tot = reinterpret_cast<double*>(sT);          // 128 doubles, transposition buffer is free again
                        tot[et] = *((volatile double*)(a.bn.acc + et * ACC_STRIDE));
                        a.bn.acc[et * ACC_STRIDE] = 0.0;
                        asm volatile("bar.sync 2, 128;" ::: "memory");
                        if (a.bn.cm) comm_allreduce_epilogue(a.bn.cm, tot, 128, et);
                        const double Mt = a.bn.cm ? (double)a.bn.rows_total : (double)a.bn_rows;
                        if (et < 64) {
                            if constexpr (EPI == 1) {
                                const double mean = tot[et] / Mt;
                                double var = tot[64 + et] / Mt - mean * mean;
                                if (var < 0) var = 0;
                                a.bn.mean[et] = (float)mean;
                                a.bn.invstd[et] = (float)(1.0 / sqrt(var + (double)a.bn_eps));
                            } else {
                                const double sdy = tot[et], sdyxh = tot[64 + et];
                                resolve(a.edgamma)[et] = (float)sdyxh;
                                resolve(a.edbeta)[et] = (float)sdy;
                                a.bn.m_dy[et] = (float)(sdy / Mt);
                                a.bn.m_dyxh[et] = (float)(sdyxh / Mt);
                            }
                        }
                        if (et == 0) *a.bn.ticket = 0u;
                    }
                }
            }
        }
    }
    if (threadIdx.x == 64) conv_stamp(a, 10);
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<128>(tmem);
    if (threadIdx.x == 0) conv_stamp(a, 11);
}

// ---------------------------------------------------------------------------------------------------------------------------
// Operand producers: write the zero-padded chunk-major hi/lo image the tensor kernels read
// ---------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ long long padded_row(const ConvGeom& g, long long crow) {
    const int hw = g.H * g.W;
    const int b = (int)(crow / hw);
    const int rem = (int)(crow - (long long)b * hw);
    const int h = rem / g.W, w = rem - h * g.W;
    return g.guard + ((long long)b * g.PH + h + 1) * g.PW + (w + 1);
}
__device__ __forceinline__ void store_split4(float* __restrict__ img, const ConvGeom& g, int chunk, long long prow, float4 y) {
    float4 hi, lo;
    tc::split_tf32(y.x, hi.x, lo.x); tc::split_tf32(y.y, hi.y, lo.y); tc::split_tf32(y.z, hi.z, lo.z); tc::split_tf32(y.w, hi.w, lo.w);
    float4* p = reinterpret_cast<float4*>(img + ((long long)chunk * g.rows + prow) * 4);
    p[0] = hi;
    reinterpret_cast<float4*>(img + ((long long)(16 + chunk) * g.rows + prow) * 4)[0] = lo;
}

// y = act(gamma*((x-mean)*invstd)+beta) -> operand image (and optionally the compact activation as well)
__global__ void __launch_bounds__(256) k_bn_apply_to_image(Ref xin, const float* __restrict__ gamma, const float* __restrict__ beta, BnScratch sc,
                                                            int act, ConvGeom g, float* __restrict__ img, float* __restrict__ compact) {
    pdl_begin();
    const float* __restrict__ x = resolve(xin);
    const long long n4 = (long long)g.B * g.H * g.W * 16;     // float4 groups: [crow][chunk]
    const long long step = (long long)gridDim.x * blockDim.x;
    __shared__ float s_mean[64], s_inv[64];
    if (threadIdx.x < 64) {
        if (sc.deferred) bn_fwd_final(sc, threadIdx.x, s_mean[threadIdx.x], s_inv[threadIdx.x]);
        else { s_mean[threadIdx.x] = sc.mean[threadIdx.x]; s_inv[threadIdx.x] = sc.invstd[threadIdx.x]; }
    }
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += step) {
        const int chunk = (int)(i & 15);
        const long long crow = i >> 4;
        const float4 v = reinterpret_cast<const float4*>(x)[i];
        const int c = chunk * 4;
        float4 y;
        y.x = act_fwd(act, __fadd_rn(__fmul_rn(gamma[c], __fmul_rn(__fsub_rn(v.x, s_mean[c]), s_inv[c])), beta[c]));
        y.y = act_fwd(act, __fadd_rn(__fmul_rn(gamma[c + 1], __fmul_rn(__fsub_rn(v.y, s_mean[c + 1]), s_inv[c + 1])), beta[c + 1]));
        y.z = act_fwd(act, __fadd_rn(__fmul_rn(gamma[c + 2], __fmul_rn(__fsub_rn(v.z, s_mean[c + 2]), s_inv[c + 2])), beta[c + 2]));
        y.w = act_fwd(act, __fadd_rn(__fmul_rn(gamma[c + 3], __fmul_rn(__fsub_rn(v.w, s_mean[c + 3]), s_inv[c + 3])), beta[c + 3]));
        store_split4(img, g, chunk, padded_row(g, crow), y);
        if (compact) reinterpret_cast<float4*>(compact)[i] = y;
    }
}

// compact NHWC [rows][64] -> operand image (used for gradients that feed dgrad/wgrad and by the unit probe)
__global__ void __launch_bounds__(256) k_compact_to_image(Ref xin, ConvGeom g, float* __restrict__ img) {
    pdl_begin();
    const float* __restrict__ x = resolve(xin);
    const long long n4 = (long long)g.B * g.H * g.W * 16;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += step) {
        const int chunk = (int)(i & 15);
        const long long crow = i >> 4;
        store_split4(img, g, chunk, padded_row(g, crow), reinterpret_cast<const float4*>(x)[i]);
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Weight gradient of the 64->64 3x3 convolution:  dW[tap][ci][co] = sum_{pixels} X[p + shift(tap)][ci] * G[p][co]
// (DL4J ConvolutionLayer.backpropGradient's weight-gradient GEMM).  The contraction runs over pixels, the output is nine
// 64x64 blocks: one CTA per image keeps all 9*64*64 partial sums in registers (144 per thread), stages the zero-padded
// image and its gradient in shared memory, and writes one partial per image; k_wgrad_reduce sums the partials in a fixed
// order (deterministic).  fp32 FMA on the CUDA cores — the tensor-core formulation needs nine 64-column TMEM accumulators
// (576 > 512 columns) and a 147 KB epilogue per CTA, which costs more than it saves at this size; the kernel runs on a side
// branch of the graph, concurrently with the dgrad/BatchNorm chain.
// ---------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1) k_conv_wgrad64(Ref xin, Ref gin, ConvGeom g, float* __restrict__ part, int images_per_cta, int nbuf_avail) {
    pdl_begin();
    extern __shared__ __align__(16) float smw[];
    const int hw = g.H * g.W, pp = g.PH * g.PW;
    // two staging buffers (when they fit): image n+1 arrives by cp.async while image n is being contracted
    const int buf_floats = (pp + hw) * 64;
    const int tci = threadIdx.x >> 4, tco = threadIdx.x & 15;
    float acc[9][4][4];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[t][i][j] = 0.f;
    const int img0 = blockIdx.x * images_per_cta, img1 = min(g.B, (blockIdx.x + 1) * images_per_cta);
    const int nbuf = (img1 - img0 > 1 && nbuf_avail == 2) ? 2 : 1;
    for (int i = threadIdx.x; i < pp * 16; i += 256) {      // zero borders (the interiors are overwritten by every image)
        reinterpret_cast<float4*>(smw)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (nbuf == 2) reinterpret_cast<float4*>(smw + buf_floats)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    const float* __restrict__ xbase = resolve(xin);
    const float* __restrict__ gbase = resolve(gin);
    auto stage = [&](int img, int b) {
        float* Xs = smw + b * buf_floats;           // [PH*PW][64], zero border
        float* Gs = Xs + pp * 64;                   // [H*W][64]
        const float4* x = reinterpret_cast<const float4*>(xbase + (long long)img * hw * 64);
        const float4* gr = reinterpret_cast<const float4*>(gbase + (long long)img * hw * 64);
        for (int i = threadIdx.x; i < hw * 16; i += 256) {
            const int p = i >> 4, c4 = i & 15;
            const int h = p / g.W, w = p - h * g.W;
            const uint32_t dx = (uint32_t)__cvta_generic_to_shared(reinterpret_cast<float4*>(Xs) + ((h + 1) * g.PW + (w + 1)) * 16 + c4);
            const uint32_t dg = (uint32_t)__cvta_generic_to_shared(reinterpret_cast<float4*>(Gs) + i);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dx), "l"(x + i) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dg), "l"(gr + i) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (img0 < img1) stage(img0, 0);
    for (int img = img0; img < img1; ++img) {
        const int b = nbuf == 2 ? (img - img0) & 1 : 0;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();             // image `img` is in buffer b for everybody; buffer b^1 (image img-1) is fully consumed
        if (nbuf == 2 && img + 1 < img1) stage(img + 1, b ^ 1);
        const float* Xs = smw + b * buf_floats;
        const float* Gs = Xs + pp * 64;
        for (int h = 0; h < g.H; ++h) {
            for (int w = 0; w < g.W; ++w) {
                const int center = (h + 1) * g.PW + (w + 1);
                const float4 gv = reinterpret_cast<const float4*>(Gs)[(h * g.W + w) * 16 + tco];
                const float gj[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const int q = center + (t / 3 - 1) * g.PW + (t % 3 - 1);
                    const float4 xv = reinterpret_cast<const float4*>(Xs)[q * 16 + tci];
                    const float xi[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[t][i][j] = fmaf(xi[i], gj[j], acc[t][i][j]);
                }
            }
        }
        if (nbuf == 1 && img + 1 < img1) { __syncthreads(); stage(img + 1, 0); }
    }
    float* __restrict__ out = part + (long long)blockIdx.x * 9 * 4096;
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int i = 0; i < 4; ++i)
            reinterpret_cast<float4*>(out + (t * 64 + tci * 4 + i) * 64)[tco] = make_float4(acc[t][i][0], acc[t][i][1], acc[t][i][2], acc[t][i][3]);
}
inline size_t conv_wgrad_smem_bytes(const ConvGeom& g) { return sizeof(float) * 64 * ((size_t)g.PH * g.PW + (size_t)g.H * g.W); }   // one staging buffer

// dst[i] = sum_b part[b][i]  (fixed order), i over 9*64*64.  Block = 64 float4 columns x 4 partial groups.
// cm != nullptr (sharded path): the sum over the ranks rides on the same launch (LL exchange of the finished float4).
__global__ void __launch_bounds__(256) k_wgrad_reduce(const float* __restrict__ part, Ref dstref, int nparts, CommDev* cm) {
    pdl_begin();
    float* __restrict__ dst = resolve(dstref);
    const unsigned long long seq = cm ? cm->seq_x + 1 : 0ull;
    const int n4 = 9 * 4096 / 4;
    const int col = blockIdx.x * 64 + (threadIdx.x & 63), grp = threadIdx.x >> 6;
    __shared__ float4 sm[4][64];
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    const int per = (nparts + 3) / 4;
    const int b0 = grp * per, b1 = min(nparts, b0 + per);
    if (col < n4) {
#pragma unroll 4
        for (int b = b0; b < b1; ++b) {
            const float4 v = reinterpret_cast<const float4*>(part)[(long long)b * n4 + col];
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
    }
    sm[grp][threadIdx.x & 63] = s;
    __syncthreads();
    if (grp == 0 && col < n4) {
        float4 t = sm[0][threadIdx.x];
#pragma unroll
        for (int q = 1; q < 4; ++q) { const float4 v = sm[q][threadIdx.x]; t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w; }
        if (cm) t = comm_ll_allreduce4(cm, (unsigned int)seq, (int)(seq & 1), (long long)col * 4, t);
        reinterpret_cast<float4*>(dst)[col] = t;
    }
    if (cm) { __syncthreads(); comm_ll_leave(cm, seq); }
}

// BatchNorm backward dx written as the compact gradient AND as the operand image of the following dgrad convolution
__global__ void __launch_bounds__(256) k_bn_bwd_dx_to_image(Ref gin, float gscale, Ref yref, Ref xin, const float* __restrict__ gamma, BnScratch sc,
                                                             int act, ConvGeom g, float* __restrict__ dx_compact, float* __restrict__ img,
                                                             Ref dgamma_ref, Ref dbeta_ref) {
    pdl_begin();
    const float* __restrict__ gg = resolve(gin);
    const float* __restrict__ y = resolve(yref);
    const float* __restrict__ x = resolve(xin);
    const long long n4 = (long long)g.B * g.H * g.W * 16;
    const long long step = (long long)gridDim.x * blockDim.x;
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
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += step) {
        const int chunk = (int)(i & 15);
        const long long crow = i >> 4;
        const float4 gv = reinterpret_cast<const float4*>(gg)[i], yv = reinterpret_cast<const float4*>(y)[i], xv = reinterpret_cast<const float4*>(x)[i];
        const float ga[4] = {gv.x, gv.y, gv.z, gv.w}, ya[4] = {yv.x, yv.y, yv.z, yv.w}, xa[4] = {xv.x, xv.y, xv.z, xv.w};
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = chunk * 4 + k;
            const float inv = s_st[1][c];
            const float dy = act_bwd(act, ya[k], __fmul_rn(gscale, ga[k]));
            const float xh = __fmul_rn(__fsub_rn(xa[k], s_st[0][c]), inv);
            const float kk = __fmul_rn(gamma[c], inv);
            o[k] = __fmul_rn(kk, __fsub_rn(__fsub_rn(dy, s_st[2][c]), __fmul_rn(xh, s_st[3][c])));
        }
        const float4 ov = make_float4(o[0], o[1], o[2], o[3]);
        reinterpret_cast<float4*>(dx_compact)[i] = ov;
        store_split4(img, g, chunk, padded_row(g, crow), ov);
    }
}

}  // namespace n4j
