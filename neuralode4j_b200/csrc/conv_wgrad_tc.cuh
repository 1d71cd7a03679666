// conv_wgrad_tc.cuh — weight gradient of the 64->64 3x3 "same" convolution on the 5th-generation tensor cores.
//
// Replaces the weight-gradient half of DL4J's ConvolutionLayer.backpropGradient reached from BackpropagateAdjoint.java:109
// (`current.doBackward`; im2col + GEMM / cuDNN in the reference):
//
//   dW[tap][ci][co] = sum_p X[p + shift(tap)][ci] * G[p][co]         p over zero-padded pixels, contraction = pixels
//
// Both operands are NHWC (channels contiguous, the contraction index is the slow one), i.e. MN-major for the tensor core.
// kind::tf32 takes MN-major operands only in the 128-byte swizzle with 32-byte atoms (SWIZZLE_128B_BASE32B, measured:
// profiles/r02_mn_swizzle_probe.txt; the no-swizzle form returns zeros): a tile is stored as
//   region[hi|lo][32-channel column block][pixel row][128 B],   32-byte chunk j of row r at chunk position j ^ (r & 3)
// and because the swizzle is a function of the ABSOLUTE shared-memory address (base_offset = 0 works for any start row,
// same probe), the nine taps are nine start addresses into ONE staged X tile — no im2col, no per-tap copies.
//
// fp32-faithful via the hi/lo TF32 split, both splits stacked inside one instruction:
//   A = [X_hi ; X_lo]  (M = 128: ci of hi, ci of lo)     B = [G_hi | G_lo]  (N = 128)      one M128 N128 K8 MMA per tap and 8 pixels
//   dW = D[ci][co] + D[ci][64+co] + D[64+ci][co] + D[64+ci][64+co]        (all four partial products, lo*lo included)
//
// Pixels are padded (H+1) x (W+1) with SHARED borders (the right neighbour of the last column is the left border of the next
// row, the row below the last one is the top border of the next image): 64 instead of 81 rows per 7x7 image.  One CTA owns
// 128 padded pixels (two MNIST images) and all nine taps: it stages X (+halo) and G once (fp32 loads, split, swizzled stores
// by all sixteen warps; the padded-row -> pixel table is computed once per CTA), then warp 0 issues the MMAs tap by tap into a
// ring of four 128-column TMEM accumulators while warps 4-7 drain finished taps (TMEM -> registers -> quadrant sum through
// shared memory -> coalesced stores of the CTA's partial).  k_wgrad_reduce sums the per-tile partials in a fixed order
// (deterministic), as before.  The grid is `tiles` CTAs (64 at batch 128) on purpose: the kernel runs on the side branch of the
// adjoint evaluation and must leave >= 81 SMs to the convolutions of the main chain (a 2 x 64-CTA version with the taps split
// over two CTAs per tile was 4 us shorter per launch but made the step SLOWER, 5.03 -> 5.33 ms: it starved the main chain).
#pragma once

#include "common.cuh"
#include "tc_common.cuh"

namespace n4j {

struct WgradTcGeom {
    int B, H, W;
    int PW;          // W + 1
    int PP;          // (H + 1) * (W + 1) padded pixels per image
    int x0;          // staged rows in front of the tile's first pixel (>= PW + 1, even)
    int rows_x;      // 128 + 2 * x0 (multiple of 4: every region stays 512-byte aligned)
    long long mrows; // B * PP
    int tiles;       // ceil(mrows / 128)
};

inline WgradTcGeom make_wgrad_tc_geom(int B, int H, int W) {
    WgradTcGeom g;
    g.B = B; g.H = H; g.W = W;
    g.PW = W + 1;
    g.PP = (H + 1) * (W + 1);
    g.x0 = (g.PW + 1 + 1) & ~1;
    g.rows_x = 128 + 2 * g.x0;
    g.mrows = (long long)B * g.PP;
    g.tiles = (int)((g.mrows + 127) / 128);
    return g;
}

constexpr int WGRAD_TC_THREADS = 512;
constexpr int WGRAD_TC_NBUF = 4;            // TMEM ring: four 128-column accumulators
constexpr int WGRAD_TC_ESTRIDE = 68;        // floats per row of the quadrant-exchange buffer (16-byte aligned rows)

inline size_t conv_wgrad_tc_smem_bytes(const WgradTcGeom& g) {
    return (size_t)4 * g.rows_x * 128 + (size_t)4 * 128 * 128 + (size_t)64 * WGRAD_TC_ESTRIDE * 4 + 256 + (size_t)g.rows_x * 4 + 1024;   // + alignment slack
}

namespace tc {
// Shared-memory matrix descriptor for an MN-major kind::tf32 operand in the SWIZZLE_128B_BASE32B layout (layout type 1):
//   ((8,n),(4,k)) : ((1,LBO),(8,SBO)) in 16-byte units — LBO = distance between 32-channel column blocks, SBO = distance
//   between groups of four pixel rows (512 B when rows are contiguous); base_offset 0 (absolute-address swizzle).
__device__ __forceinline__ uint64_t smem_desc_mn32(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return smem_desc(saddr, lbo_bytes, sbo_bytes) | (1ull << 61);
}
__device__ __forceinline__ void tmem_alloc512(uint32_t* slot_in_smem) { tmem_alloc<512>(slot_in_smem); }
}  // namespace tc

// grid = tiles, block = 512, dynamic shared memory = conv_wgrad_tc_smem_bytes(g)
// part: [tiles][9][64][64] partial weight gradients (every entry is written)
__global__ void __launch_bounds__(WGRAD_TC_THREADS, 1) k_conv_wgrad_tc(Ref xin, Ref gin, WgradTcGeom g, float* __restrict__ part, long long* dbg) {
#define WG_STAMP(slot) do { if (dbg && blockIdx.x == 0) dbg[slot] = clock64(); } while (0)
    if (threadIdx.x == 0) WG_STAMP(0);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    extern __shared__ uint8_t smem_dyn[];
    uint8_t* sm = reinterpret_cast<uint8_t*>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    const uint32_t XR = (uint32_t)g.rows_x * 128u;      // bytes per X region
    constexpr uint32_t GR = 128u * 128u;                // bytes per G region
    uint8_t* sX = sm;                                    // [hi c0][hi c1][lo c0][lo c1][rows_x][128 B]
    uint8_t* sG = sX + 4 * XR;                           // same four regions, 128 rows
    float* sE = reinterpret_cast<float*>(sG + 4 * GR);   // [64][ESTRIDE] quadrant exchange / coalescing buffer
    uint64_t* bars = reinterpret_cast<uint64_t*>(sE + 64 * WGRAD_TC_ESTRIDE);
    uint64_t* full = bars;                               // [NBUF]  accumulator complete
    uint64_t* empty = bars + WGRAD_TC_NBUF;              // [NBUF]  accumulator drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * WGRAD_TC_NBUF);
    int* s_idx = reinterpret_cast<int*>(bars + 2 * WGRAD_TC_NBUF + 2);     // [rows_x] padded row -> compact pixel index * 16, or -1

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x;
    constexpr int tap0 = 0, tap1 = 9;
    const long long m0 = (long long)tile * 128;

    if (threadIdx.x == 0) {
        for (int b = 0; b < WGRAD_TC_NBUF; ++b) { tc::mbar_init(&full[b], 1); tc::mbar_init(&empty[b], 4); }
        tc::fence_barrier_init();
    }
    if (warp == 0) tc::tmem_alloc512(tmem_slot);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (threadIdx.x == 0) WG_STAMP(1);

    // ---- staging, all sixteen warps: thread = (row group of 32, 16-byte channel chunk); every load is issued before the first use ----
    {
        const float4* __restrict__ x = reinterpret_cast<const float4*>(resolve(xin));
        const float4* __restrict__ gr = reinterpret_cast<const float4*>(resolve(gin));
        const int chunk = threadIdx.x & 15, rg = threadIdx.x >> 4;
        const uint32_t cb_off = (uint32_t)(chunk >> 3);                  // 32-channel column block
        const uint32_t j32 = (uint32_t)((chunk & 7) >> 1), sub = (uint32_t)(chunk & 1) * 16u;
        // padded row -> pixel: once per CTA (two integer divisions per row), not once per (row, chunk)
        for (int r = threadIdx.x; r < g.rows_x; r += WGRAD_TC_THREADS) {
            const long long pr = m0 - g.x0 + r;
            int idx = -1;
            if (pr >= 0 && pr < g.mrows) {
                const unsigned p = (unsigned)pr;
                const unsigned b = p / (unsigned)g.PP, rem = p - b * (unsigned)g.PP;
                const unsigned hh = rem / (unsigned)g.PW, ww = rem - hh * (unsigned)g.PW;
                if (hh != 0 && ww != 0) idx = (int)(((b * (unsigned)g.H + hh - 1) * (unsigned)g.W + ww - 1) * 16u);
            }
            s_idx[r] = idx;
        }
        __syncthreads();
        auto compact_index = [&](int r) -> int {                         // r: row of the staged X tile (G row p = X row p + x0)
            const int i = s_idx[r];
            return i < 0 ? -1 : i + chunk;
        };
        auto store_split = [&](uint8_t* base, uint32_t region_bytes, int r, float4 v) {
            float4 hi, lo;
            tc::split_tf32(v.x, hi.x, lo.x); tc::split_tf32(v.y, hi.y, lo.y); tc::split_tf32(v.z, hi.z, lo.z); tc::split_tf32(v.w, hi.w, lo.w);
            const uint32_t off = (uint32_t)r * 128u + ((j32 ^ ((uint32_t)r & 3u)) << 5) + sub;
            *reinterpret_cast<float4*>(base + cb_off * region_bytes + off) = hi;
            *reinterpret_cast<float4*>(base + (2u + cb_off) * region_bytes + off) = lo;
        };
        constexpr int UX = 5, UG = 4;
        float4 gv[UG];
        int gi[UG];
#pragma unroll
        for (int u = 0; u < UG; ++u) {
            gi[u] = compact_index(g.x0 + rg + 32 * u);
            if (gi[u] >= 0) gv[u] = gr[gi[u]];
        }
        for (int rb = 0; rb < g.rows_x; rb += 32 * UX) {                  // one pass for W <= 14
            float4 xv[UX];
            int xi[UX];
#pragma unroll
            for (int u = 0; u < UX; ++u) {
                const int r = rb + rg + 32 * u;
                xi[u] = r < g.rows_x ? compact_index(r) : -1;
                if (xi[u] >= 0) xv[u] = x[xi[u]];
            }
#pragma unroll
            for (int u = 0; u < UX; ++u) {
                const int r = rb + rg + 32 * u;
                if (r < g.rows_x) store_split(sX, XR, r, xi[u] >= 0 ? xv[u] : make_float4(0.f, 0.f, 0.f, 0.f));
            }
        }
#pragma unroll
        for (int u = 0; u < UG; ++u) store_split(sG, GR, rg + 32 * u, gi[u] >= 0 ? gv[u] : make_float4(0.f, 0.f, 0.f, 0.f));
    }
    if (threadIdx.x == 0) WG_STAMP(2);
    tc::fence_proxy_async();          // generic-proxy stores -> visible to the tensor core (async proxy)
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (threadIdx.x == 0) WG_STAMP(3);

    if (warp == 0) {
        if (lane == 0) {
            // ---- MMA issuer: per tap 16 K steps of 8 pixels, D[128 x 128] (+)= [X_hi ; X_lo]^T-major x [G_hi | G_lo] ----
            constexpr uint32_t idesc = tc::idesc_tf32(128, 128, 1, 1);
            const uint32_t sX_u = tc::smem_u32(sX), sG_u = tc::smem_u32(sG);
            for (int tap = tap0, i = 0; tap < tap1; ++tap, ++i) {
                const int buf = i & (WGRAD_TC_NBUF - 1);
                if (i >= WGRAD_TC_NBUF) { tc::mbar_wait(&empty[buf], ((i / WGRAD_TC_NBUF) - 1) & 1); tc::tc_fence_after(); }
                const int shift = (tap / 3 - 1) * g.PW + (tap % 3 - 1);
                const uint32_t a0 = sX_u + (uint32_t)(g.x0 + shift) * 128u;
                const uint32_t d = tmem + (uint32_t)buf * 128u;
#pragma unroll 4
                for (int ks = 0; ks < 16; ++ks) {
                    const uint64_t da = tc::smem_desc_mn32(a0 + (uint32_t)ks * 1024u, XR, 512u);
                    const uint64_t db = tc::smem_desc_mn32(sG_u + (uint32_t)ks * 1024u, GR, 512u);
                    tc::mma_tf32(d, da, db, idesc, ks ? 1u : 0u);
                }
                tc::mma_commit(&full[buf]);
                if (i < 4) WG_STAMP(4 + i); else if (i == 8) WG_STAMP(8);
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ---- epilogue, four warps: TMEM lane quadrant q = rows 32q..32q+31 of D; rows 0-63 belong to X_hi, 64-127 to X_lo ----
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int et = threadIdx.x - 128;            // 0..127
        float* __restrict__ out = part + (long long)tile * 9 * 4096;
        for (int tap = tap0, i = 0; tap < tap1; ++tap, ++i) {
            const int buf = i & (WGRAD_TC_NBUF - 1);
            tc::mbar_wait(&full[buf], (i / WGRAD_TC_NBUF) & 1);
            tc::tc_fence_after();
            if (et == 0) { if (i < 4) WG_STAMP(10 + i); else if (i == 8) WG_STAMP(14); }
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * 128u;
            // quadrant sum: the X_lo rows (lanes 64-127) go through shared memory, the X_hi rows add them and leave the finished
            // 64 x 64 block there.  Columns: [0,64) = G_hi, [64,128) = G_lo; 32 output channels at a time (register budget).
            if (q >= 2) {
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    float a[32], b[32];
                    tc::tmem_ld32(taddr + half * 32, a);
                    tc::tmem_ld32(taddr + 64 + half * 32, b);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<float4*>(sE + (row - 64) * WGRAD_TC_ESTRIDE + half * 32 + 4 * j) =
                            make_float4(a[4 * j] + b[4 * j], a[4 * j + 1] + b[4 * j + 1], a[4 * j + 2] + b[4 * j + 2], a[4 * j + 3] + b[4 * j + 3]);
                }
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&empty[buf]);
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (q < 2) {
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    float a[32], b[32];
                    tc::tmem_ld32(taddr + half * 32, a);
                    tc::tmem_ld32(taddr + 64 + half * 32, b);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4* p = reinterpret_cast<float4*>(sE + row * WGRAD_TC_ESTRIDE + half * 32 + 4 * j);
                        const float4 u = *p;
                        *p = make_float4((a[4 * j] + b[4 * j]) + u.x, (a[4 * j + 1] + b[4 * j + 1]) + u.y, (a[4 * j + 2] + b[4 * j + 2]) + u.z,
                                         (a[4 * j + 3] + b[4 * j + 3]) + u.w);
                    }
                }
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&empty[buf]);       // the accumulator has been read: the issuer may overwrite it
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            float4* __restrict__ dst = reinterpret_cast<float4*>(out + tap * 4096);
#pragma unroll
            for (int k = 0; k < 8; ++k) {                      // 1024 float4 of the block, 128 threads: coalesced 256-byte rows
                const int idx = et + 128 * k;
                dst[idx] = *reinterpret_cast<const float4*>(sE + (idx >> 4) * WGRAD_TC_ESTRIDE + (idx & 15) * 4);
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");    // sE is free for the next tap
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<512>(tmem);
    if (threadIdx.x == 0) WG_STAMP(30);
#undef WG_STAMP
}

}  // namespace n4j
