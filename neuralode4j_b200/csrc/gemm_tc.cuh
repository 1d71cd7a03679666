// gemm_tc.cuh — dense-layer contractions of the RHS on the 5th-generation tensor cores (3xTF32, fp32-faithful).
//
// Replaces DL4J's DenseLayer forward / backpropGradient GEMMs reached from ForwardPass.java:81 and
// BackpropagateAdjoint.java:109 (cuBLAS sgemm / OpenBLAS in the reference).
//
// Operands are pre-split (hi = tf32(x), lo = x - hi) chunk-major images in global memory (tc_common.cuh):
//     img[row / 128][hl*nch + chunk][row % 128][4 floats]     chunk = 4 consecutive features, row = sample (or weight row)
// moved by 1-D bulk copies (cp.async.bulk: one copy per hi / lo A tile, one per stacked B tile) straight into the
// canonical no-swizzle UMMA layout.  (The tensor-map variant, cp.async.bulk.tensor.3d/4d with 16-byte innermost rows, was
// TMA-element-bound: 620 us instead of 219 us for 1024 x 4096 x 4096.)
// One K-major kernel serves all three contractions of a dense layer:
//   C[m, n] = sum_k A[m, k] * W[n, k]
//   forward: A = x image,        W = stacked weight image [nOut][nIn]
//   dgrad  : A = delta image,    W = stacked image of the transposed weights
//   wgrad  : A = delta^T image (rows = out features, K = batch), W = stacked x^T image  ->  dW^T[o, i] = sum_b delta[b, o] x[b, i]
// (an MN-major variant that reads the untransposed images was tried first: with SWIZZLE_NONE descriptors it returned zeros on
//  B200, so the transposed images are produced explicitly by k_rowmajor_to_image_T — one extra pass over x and delta.)
// Tile 128 x 128, K block of 32, 3-stage bulk-copy/mbarrier ring, 3xTF32 as  A_hi*[B_hi | B_lo]  (N = 256) + A_lo*B_hi (N = 128)
// into 256 TMEM columns; the epilogue adds the two halves, applies bias / activation / activation derivative and writes
// row-major fp32 (and, optionally, the split image the next GEMM reads).
#pragma once

#include <cuda.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace n4j {

constexpr int GEMM_TC_THREADS = 192;
constexpr int GEMM_TC_STAGES = 3;
constexpr int GEMM_TC_STAGE_BYTES = 65536;                 // MH = 1: K block 32, A (hi+lo) 32 KB + B (hi+lo) 32 KB
constexpr size_t GEMM_TC_SMEM = (size_t)GEMM_TC_STAGES * GEMM_TC_STAGE_BYTES + 256;
// MH = 2 (256 x 128 tile: two M halves share every B tile): K block 16, A 2 x (hi+lo) x 8 KB + B 16 KB = 48 KB, four stages.
// The kernel is bounded by the L2 -> SM ingest (measured ~42 B/clk per SM with all SMs pulling): 64 KB per 816 MMA cycles with
// MH = 1, 48 KB per 816 cycles with MH = 2.
constexpr int GEMM_TC2_STAGES = 4;
constexpr int GEMM_TC2_STAGE_BYTES = 49152;
constexpr size_t GEMM_TC2_SMEM = (size_t)GEMM_TC2_STAGES * GEMM_TC2_STAGE_BYTES + 256;

namespace tc {
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
                 : "memory");
}
}  // namespace tc

struct GemmTcArgs {
    int M, N;             // logical output extent
    int kblocks;          // K / 32
    int a_nch;            // chunks per hi/lo half of the A image
    Ref out;              // row-major [M][N]
    const float* bias;    // [N] or null (forward)
    int act;              // activation applied to the result (forward)
    Ref dact_y;           // optional: multiply the result by act'(y) with y row-major [M][N] (dgrad through the previous layer's
    int dact;             //           activation); dact = activation kind, dact_y.base == null -> off
    // operand images in global memory, moved with 1-D bulk copies (2 KB per chunk of the A tile, ONE 32 KB copy for the stacked B
    // tile).  The 3-d/4-d tensor-map path needed one TMA element per 16-byte row (4096 per stage) and bounded the kernel at ~4000
    // cycles per K block against 816 cycles of MMA: 620 -> see DESIGN.md.
    const float* a_img;   // [a_rows / 128][2*a_nch][128][4]  (row-tile-major: the chunks of one 128-row tile are contiguous)
    long long a_rows;
    const float* b_img;   // [nblocks][b_nch][256][4]
    int b_nch;
    float* out_img;       // optional: also write the result as a split image [2*ceil(N/4)][img_rows][4] for the next GEMM
    long long out_img_rows;
};

// tmA = activation image (3-d), tmB = stacked image (4-d: {4, 256, nchunks, nblocks}).
template <int MH>
__global__ void __launch_bounds__(GEMM_TC_THREADS, 1) k_gemm_tc(GemmTcArgs a) {
    constexpr int STAGES = MH == 1 ? GEMM_TC_STAGES : GEMM_TC2_STAGES;
    constexpr int STAGE_BYTES = MH == 1 ? GEMM_TC_STAGE_BYTES : GEMM_TC2_STAGE_BYTES;
    constexpr int KCH = MH == 1 ? 8 : 4;                  // 16-byte chunks (4 floats of K) per stage
    constexpr int A_HALF_BYTES = KCH * 2048;              // one hi or lo tile of one M half
    constexpr int B_OFF = MH * 2 * A_HALF_BYTES;          // stacked B tile behind the A tiles
    constexpr int TCOLS = MH * 256;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + STAGES * STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* d_full = bars + 2 * STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_n = (a.N + 127) / 128;
    const int tm = blockIdx.x / tiles_n, tn = blockIdx.x % tiles_n;
    const int m0 = tm * 128 * MH, n0 = tn * 128;
    const bool second = MH == 2 && (long long)m0 + 128 < a.a_rows;      // the A image is padded to 128 rows, not 256

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
        tc::mbar_init(d_full, 1);
        tc::fence_barrier_init();
    }
    if (warp == 0) tc::tmem_alloc<TCOLS>(tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp == 0) {
        if (lane == 0) {
            const int kblocks = a.kblocks * (8 / KCH);          // a.kblocks counts K blocks of 32
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % STAGES;
                if (kb >= STAGES) tc::mbar_wait(&empty[s], ((kb / STAGES) - 1) & 1);
                uint8_t* sA = smem_raw + s * STAGE_BYTES;
                uint8_t* sB = sA + B_OFF;
                tc::mbar_arrive_expect_tx(&full[s], (second || MH == 1) ? STAGE_BYTES : STAGE_BYTES - 2 * A_HALF_BYTES);
#pragma unroll
                for (int h = 0; h < MH; ++h) {
                    if (h == 1 && !second) break;
                    // tile-major image: the KCH chunks of one row tile are contiguous -> one copy per hi / lo tile
                    const float* ah = a.a_img + (((long long)(m0 / 128 + h) * (2 * a.a_nch) + kb * KCH) * 128) * 4;
                    tc::bulk_g2s(sA + (2 * h) * A_HALF_BYTES, ah, A_HALF_BYTES, &full[s]);
                    tc::bulk_g2s(sA + (2 * h + 1) * A_HALF_BYTES, ah + (long long)a.a_nch * 512, A_HALF_BYTES, &full[s]);
                }
                tc::bulk_g2s(sB, a.b_img + ((long long)tn * a.b_nch + kb * KCH) * 1024, KCH * 4096, &full[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t id256 = tc::idesc_tf32(128, 256, 0, 0);
            constexpr uint32_t id128 = tc::idesc_tf32(128, 128, 0, 0);
            uint32_t acc = 0;
            const int kblocks = a.kblocks * (8 / KCH);
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % STAGES;
                tc::mbar_wait(&full[s], (kb / STAGES) & 1);
                tc::tc_fence_after();
                const uint32_t sA = tc::smem_u32(smem_raw + s * STAGE_BYTES), sB = sA + B_OFF;
#pragma unroll
                for (int j = 0; j < KCH / 2; ++j) {
                    // K-major: chunk stride = LBO, 8-row groups packed (SBO = 128)
                    const uint64_t dB = tc::smem_desc(sB + (2 * j) * 4096, 4096, 128);
#pragma unroll
                    for (int h = 0; h < MH; ++h) {
                        if (h == 1 && !second) break;
                        const uint64_t dAh = tc::smem_desc(sA + (2 * h) * A_HALF_BYTES + (2 * j) * 2048, 2048, 128);
                        const uint64_t dAl = tc::smem_desc(sA + (2 * h + 1) * A_HALF_BYTES + (2 * j) * 2048, 2048, 128);
                        tc::mma_tf32(tmem + h * 256, dAh, dB, id256, acc);
                        tc::mma_tf32(tmem + h * 256, dAl, dB, id128, 1);
                    }
                    acc = 1;
                }
                tc::mma_commit(&empty[s]);
            }
            tc::mma_commit(d_full);
        }
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        float* __restrict__ out = resolve(a.out);
        const float* __restrict__ yprev = a.dact_y.base ? resolve(a.dact_y) : nullptr;
        tc::mbar_wait(d_full, 0);
        tc::tc_fence_after();
#pragma unroll 1
        for (int hc = 0; hc < MH * 4; ++hc) {
            const int h = hc >> 2, c32 = hc & 3;
            if (h == 1 && !second) break;
            const int m = m0 + h * 128 + row;
            float v[32], u[32];
            tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + h * 256 + c32 * 32, v);
            tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + h * 256 + 128 + c32 * 32, u);
            const int nb = n0 + c32 * 32;
            if (m < a.M) {
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    const int n = nb + j4 * 4;
                    if (n >= a.N) break;
                    float r[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float x = v[j4 * 4 + e] + u[j4 * 4 + e];
                        if (a.bias) x += a.bias[min(n + e, a.N - 1)];
                        r[e] = act_fwd(a.act, x);
                    }
                    if (yprev) {
                        const float4 y = *reinterpret_cast<const float4*>(yprev + (long long)m * a.N + n);
                        r[0] = act_bwd(a.dact, y.x, r[0]); r[1] = act_bwd(a.dact, y.y, r[1]);
                        r[2] = act_bwd(a.dact, y.z, r[2]); r[3] = act_bwd(a.dact, y.w, r[3]);
                    }
                    *reinterpret_cast<float4*>(out + (long long)m * a.N + n) = make_float4(r[0], r[1], r[2], r[3]);
                    if (a.out_img) {
                        float4 hi, lo;
                        tc::split_tf32(r[0], hi.x, lo.x); tc::split_tf32(r[1], hi.y, lo.y);
                        tc::split_tf32(r[2], hi.z, lo.z); tc::split_tf32(r[3], hi.w, lo.w);
                        const long long nch = (a.N + 3) / 4, ch = n >> 2;
                        const long long base = ((((long long)m >> 7) * (2 * nch) + ch) * 128 + (m & 127)) * 4;
                        *reinterpret_cast<float4*>(a.out_img + base) = hi;
                        *reinterpret_cast<float4*>(a.out_img + base + nch * 512) = lo;
                    }
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<TCOLS>(tmem);
}

// ---- operand image producers ---------------------------------------------------------------------------------------------------
// row-major [rows][cols] (cols % 4 == 0) -> split image [2*cols/4][img_rows][4]; rows beyond `rows` stay zero (written once at alloc)
__global__ void __launch_bounds__(256) k_rowmajor_to_image(Ref xin, float scale, float* __restrict__ img, long long rows, int cols, long long img_rows) {
    pdl_begin();
    const float* __restrict__ x = resolve(xin);
    const int nch = cols >> 2;
    const long long n4 = rows * nch;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += step) {
        const long long r = i / nch;
        const int ch = (int)(i - r * nch);
        float4 v = reinterpret_cast<const float4*>(x)[i];
        v.x = __fmul_rn(v.x, scale); v.y = __fmul_rn(v.y, scale); v.z = __fmul_rn(v.z, scale); v.w = __fmul_rn(v.w, scale);
        float4 hi, lo;
        tc::split_tf32(v.x, hi.x, lo.x); tc::split_tf32(v.y, hi.y, lo.y); tc::split_tf32(v.z, hi.z, lo.z); tc::split_tf32(v.w, hi.w, lo.w);
        const long long base = (((r >> 7) * (2 * nch) + ch) * 128 + (r & 127)) * 4;      // [row tile][hl*nch + chunk][row % 128][4]
        *reinterpret_cast<float4*>(img + base) = hi;
        *reinterpret_cast<float4*>(img + base + (long long)nch * 512) = lo;
    }
}
// row-major [rows][cols] -> split image of the TRANSPOSE: chunk = 4 consecutive rows (samples), image row = column (feature).
//   stacked = 0: plain image   [2*nchT][img_cols][4]          (A operand of the weight-gradient GEMM: delta^T)
//   stacked = 1: stacked image [nblock][nchT][hl*128 + col % 128][4]   (B operand: x^T)
// nchT = img_rows / 4 chunks along the batch; samples beyond `rows` read as zero.
__global__ void __launch_bounds__(256) k_rowmajor_to_image_T(Ref xin, float* __restrict__ img, long long rows, int cols, long long img_rows, long long img_cols,
                                                              int stacked) {
    pdl_begin();
    const float* __restrict__ x = resolve(xin);
    const long long nchT = img_rows >> 2;
    const long long total = nchT * cols;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step) {
        const long long ch = i / cols;
        const int f = (int)(i - ch * cols);
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) { const long long r = ch * 4 + e; v[e] = r < rows ? x[r * cols + f] : 0.f; }
        float4 hi, lo;
        tc::split_tf32(v[0], hi.x, lo.x); tc::split_tf32(v[1], hi.y, lo.y); tc::split_tf32(v[2], hi.z, lo.z); tc::split_tf32(v[3], hi.w, lo.w);
        if (!stacked) {
            const long long base = (((long long)(f >> 7) * (2 * nchT) + ch) * 128 + (f & 127)) * 4;   // tile-major like k_rowmajor_to_image
            *reinterpret_cast<float4*>(img + base) = hi;
            *reinterpret_cast<float4*>(img + base + nchT * 512) = lo;
        } else {
            const long long base = (((long long)(f >> 7) * nchT + ch) * 256 + (f & 127)) * 4;
            *reinterpret_cast<float4*>(img + base) = hi;
            *reinterpret_cast<float4*>(img + base + 128 * 4) = lo;
        }
    }
}

// weights W(k, n) -> stacked weight image [nblock][chunk = k/4][hl*128 + (n % 128)][k % 4], zero-padded to 128 rows per block.
//   transpose = 0: W(k, n) = src[n * K + k]   (forward:  n = out feature, k = in feature, src = DL4J's f-order [nIn,nOut] view)
//   transpose = 1: W(k, n) = src[k * N + n]   (dgrad:    n = in feature,  k = out feature)
__global__ void __launch_bounds__(256) k_weights_to_stacked_image(const float* __restrict__ src, float* __restrict__ img, int K, int N, int transpose) {
    pdl_begin();
    const long long total = (long long)K * N;
    const int nch = K >> 2;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step) {
        int k, n;
        if (!transpose) { n = (int)(i / K); k = (int)(i - (long long)n * K); }
        else { k = (int)(i / N); n = (int)(i - (long long)k * N); }
        float hi, lo;
        tc::split_tf32(src[i], hi, lo);
        const long long base = ((((long long)(n >> 7) * nch + (k >> 2)) * 256) + (n & 127)) * 4 + (k & 3);
        img[base] = hi;
        img[base + 128 * 4] = lo;
    }
}

}  // namespace n4j
