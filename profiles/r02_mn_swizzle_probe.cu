// PREPARED IN r01, NOT RUN YET (the round's GPU budget was spent) — first thing to run in r02:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o gpurun_out/mn_probe profiles/r02_mn_swizzle_probe.cu && gpurun_out/mn_probe
//
// Question: which shared-memory layout does tcgen05.mma.kind::tf32 accept for an MN-MAJOR operand (the contraction index — pixels,
// for the conv weight gradient — is the slow one, channels are contiguous)?  r01's probe (profiles/r01_mn_major_probe.*) showed an
// all-zero accumulator for the no-swizzle INTERLEAVE form.  CUTLASS (cute/atom/mma_traits_sm100.hpp) lists, in 16-byte units,
//   128B_BASE32B : Swizzle<2,5,2> o ((8,n),(4,k)) : ((1,LBO),(8,SBO))   rows of 128 B (32 channels) per k, 32-byte chunks XOR (k & 3)
//   128B         : Swizzle<3,4,3> o ((8,n),(8,k)) : ((1,LBO),(8,SBO))   rows of 128 B per k, 16-byte chunks XOR (k & 7)
// D[128 x 64] = sum_k A[k][m] * B[k][n]; A is the operand under test (four 32-channel column blocks, [blk][k][32]), B stays in the
// K-major no-swizzle layout the conv kernels use (known good).  Variants: layout type x swizzle function x (LBO, SBO) assignment.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../neuralode4j_b200/csrc/tc_common.cuh"
using namespace n4j;

constexpr int K = 32;      // pixels

__device__ __forceinline__ uint64_t desc_typed(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout_type, uint32_t base_offset) {
    return tc::smem_desc(saddr, lbo, sbo) | ((uint64_t)(base_offset & 7u) << 49) | ((uint64_t)(layout_type & 7u) << 61);
}

// swz: 0 none, 1 = 32-byte chunks ^ (k & 3) [Swizzle<2,5,2>], 2 = 16-byte chunks ^ (k & 7) [Swizzle<3,4,3>]
__global__ void k_probe(const float* A /*[K][128]*/, const float* Bk /*K-major image*/, float* D, int swz, uint32_t layout_type, uint32_t lbo,
                        uint32_t sbo, int k_shift_rows, int zero_base_offset) {
    extern __shared__ __align__(1024) uint8_t sm[];
    float* sA = reinterpret_cast<float*>(sm);                    // 4 blocks x (K + 8 spare rows) x 128 B
    constexpr int ROWS = K + 8;
    float* sB = sA + 4 * ROWS * 32;
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 4 * ROWS * 32; i += blockDim.x) sA[i] = 0.f;
    __syncthreads();
    for (int i = threadIdx.x; i < K * 128; i += blockDim.x) {
        const int k = i / 128, m = i % 128, blk = m / 32, c = m % 32;
        const int row = k + k_shift_rows;                          // the tile may start a few rows into the buffer (tap shift)
        int off = c * 4;                                           // byte offset inside the 128-byte row
        if (swz == 1) off = (((c / 8) ^ (row & 3)) * 32) + (c % 8) * 4;
        if (swz == 2) off = (((c / 4) ^ (row & 7)) * 16) + (c % 4) * 4;
        sA[(blk * ROWS + row) * 32 + off / 4] = A[i];
    }
    for (int i = threadIdx.x; i < 64 * K; i += blockDim.x) sB[i] = Bk[i];
    if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
    if (threadIdx.x < 32) tc::tmem_alloc<64>(&slot);
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = tc::idesc_tf32(128, 64, 1, 0);
        for (int ks = 0; ks < K / 8; ++ks) {
            const uint32_t a0 = tc::smem_u32(sA) + (uint32_t)(ks * 8 + k_shift_rows) * 128u;
            const uint64_t da = desc_typed(a0, lbo, sbo, layout_type, (layout_type && !zero_base_offset) ? ((a0 >> 7) & 7u) : 0u);
            const uint64_t db = tc::smem_desc(tc::smem_u32(sB) + ks * 2 * 64 * 16, 64 * 16, 128);
            tc::mma_tf32(tmem, da, db, idesc, ks ? 1u : 0u);
        }
        tc::mma_commit(&bar);
    }
    __syncthreads();
    if (threadIdx.x < 128) {
        tc::mbar_wait(&bar, 0);
        tc::tc_fence_after();
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int half = 0; half < 2; ++half) {
            float v[32];
            tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + half * 32, v);
            for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * 64 + half * 32 + j] = v[j];
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tc::tmem_dealloc<64>(slot);
}

int main() {
    std::vector<float> a(K * 128), b(64 * K), Bk(64 * K), ref(128 * 64), got(128 * 64);
    srand(1);
    auto tf = [](float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; float y; memcpy(&y, &u, 4); return y; };
    for (auto& x : a) x = tf((rand() % 2001 - 1000) / 500.f);          // a[k][m]
    for (auto& x : b) x = tf((rand() % 2001 - 1000) / 500.f);          // b[n][k]
    for (int n = 0; n < 64; ++n) for (int k = 0; k < K; ++k) Bk[((k / 4) * 64 + n) * 4 + k % 4] = b[n * K + k];
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)a[k * 128 + m] * b[n * K + k]; ref[m * 64 + n] = (float)s; }
    float *dA, *dB, *dD;
    cudaMalloc(&dA, sizeof(float) * K * 128); cudaMalloc(&dB, sizeof(float) * 64 * K); cudaMalloc(&dD, sizeof(float) * 128 * 64);
    cudaMemcpy(dA, a.data(), sizeof(float) * K * 128, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, Bk.data(), sizeof(float) * 64 * K, cudaMemcpyHostToDevice);
    const size_t smem = sizeof(float) * (4 * (K + 8) * 32 + 64 * K) + 2048;
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const uint32_t blk = (K + 8) * 128;      // bytes between two 32-channel column blocks
    struct V { const char* name; int swz; uint32_t type, lbo, sbo; int shift, zero_bo; };
    V vs[] = {{"B128_BASE32B swz<2,5,2> lbo=blk sbo=512", 1, 1, blk, 512, 0},
              {"B128_BASE32B swz<2,5,2> lbo=512 sbo=blk", 1, 1, 512, blk, 0},
              {"B128_BASE32B swz<2,5,2> lbo=blk sbo=1024", 1, 1, blk, 1024, 0},
              {"B128_BASE32B no swizzle in data (control)", 0, 1, blk, 512, 0},
              {"B128 swz<3,4,3> lbo=blk sbo=1024", 2, 2, blk, 1024, 0},
              {"B128 swz<3,4,3> lbo=1024 sbo=blk", 2, 2, 1024, blk, 0},
              {"B128_BASE32B shifted by 1 row (tap shift, base_offset)", 1, 1, blk, 512, 1},
              {"B128_BASE32B shifted by 3 rows", 1, 1, blk, 512, 3},
              {"B128 shifted by 1 row", 2, 2, blk, 1024, 1},
              {"B128_BASE32B shifted by 1 row, base_offset = 0 (absolute-address swizzle?)", 1, 1, blk, 512, 1, 1},
              {"B128 shifted by 1 row, base_offset = 0", 2, 2, blk, 1024, 1, 1}};
    for (auto& v : vs) {
        cudaMemset(dD, 0xFF, sizeof(float) * 128 * 64);
        k_probe<<<1, 128, smem>>>(dA, dB, dD, v.swz, v.type, v.lbo, v.sbo, v.shift, v.zero_bo);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(got.data(), dD, sizeof(float) * 128 * 64, cudaMemcpyDeviceToHost);
        double err = 0, mx = 0; int nz = 0;
        for (int i = 0; i < 128 * 64; ++i) { err = fmax(err, fabs(got[i] - ref[i])); mx = fmax(mx, fabs(ref[i])); nz += got[i] != 0.f; }
        printf("%-58s cuda=%s max err %.4e (ref max %.3f) nonzero %d\n", v.name, cudaGetErrorString(e), err, mx, nz);
        if (e != cudaSuccess) break;
    }
    return 0;
}
