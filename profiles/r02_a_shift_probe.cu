// r02 probe: what do ROW-SHIFTED A operands cost the conv kernels?  The implicit-GEMM convolution reads its nine taps as nine start
// addresses into one staged tile; six of them are shifted by a number of rows that is not a multiple of 8.  In the no-swizzle K-major
// layout ([k/4][row][16 B]) a core matrix (8 rows x 16 B) then straddles two 128-byte lines.  Variants: layout {none, SWIZZLE_128B K-major}
// x row shift {0, 1, 9, 10}: correctness and cycles per M128 N128 K8 instruction (B always aligned, no swizzle).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o scratch/bin/shift_probe profiles/r02_a_shift_probe.cu && scratch/bin/shift_probe
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../neuralode4j_b200/csrc/tc_common.cuh"
using namespace n4j;

constexpr int K = 32;        // one 128-byte swizzle atom of tf32 = 4 K steps
constexpr int ROWS = 128 + 16;

__global__ void k_probe(const float* A /*[ROWS][K] row-major, logical rows*/, const float* B_km, float* D, long long* cycles, int swz, int shift, int reps, int nmode) {
    extern __shared__ __align__(1024) uint8_t sm[];
    float* sA = reinterpret_cast<float*>(sm);                       // none: [K/4][ROWS][4]; swizzled: [ROWS][32 floats]
    float* sB = sA + ROWS * K;                                      // [K/4][128][4]
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < ROWS * K; i += blockDim.x) {
        const int r = i / K, k = i % K;
        if (swz) sA[r * 32 + (((k >> 2) ^ (r & 7)) << 2) + (k & 3)] = A[i];
        else sA[((k >> 2) * ROWS + r) * 4 + (k & 3)] = A[i];
    }
    for (int i = threadIdx.x; i < 128 * K; i += blockDim.x) sB[i] = B_km[i];
    if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
    if (threadIdx.x < 32) tc::tmem_alloc<128>(&slot);
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = slot;
    long long t0 = 0;
    if (threadIdx.x == 0) {
        const uint32_t idesc = tc::idesc_tf32(128, 128, 0, 0);
        t0 = clock64();
        for (int r = 0; r < reps; ++r)
            for (int ks = 0; ks < K / 8; ++ks) {
                uint64_t da;
                if (swz) da = tc::smem_desc(tc::smem_u32(sA) + shift * 128 + ks * 32, 16, 1024) | (2ull << 61);      // SWIZZLE_128B, SBO = 8 rows
                else da = tc::smem_desc(tc::smem_u32(sA) + (ks * 2 * ROWS + shift) * 16, ROWS * 16, 128);
                const uint64_t db = tc::smem_desc(tc::smem_u32(sB) + ks * 2 * 128 * 16, 128 * 16, 128);
                // nmode 0: N128 only; 1: N64 only; 2: the conv kernels' pair per K step (N128 then N64 into the first 64 columns)
                if (nmode != 1) tc::mma_tf32(tmem, da, db, idesc, (r | ks) ? 1u : 0u);
                if (nmode != 0) tc::mma_tf32(tmem, da, db, tc::idesc_tf32(128, 64, 0, 0), (nmode == 2 || (r | ks)) ? 1u : 0u);
            }
        tc::mma_commit(&bar);
    }
    __syncthreads();
    tc::mbar_wait(&bar, 0);
    if (threadIdx.x == 0) *cycles = clock64() - t0;
    tc::tc_fence_after();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int q = 0; q < 4; ++q) {
        float v[32];
        tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + q * 32, v);
        for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * 128 + q * 32 + j] = v[j];
    }
    tc::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tc::tmem_dealloc<128>(tmem);
}

int main() {
    std::vector<float> a(ROWS * K), b(128 * K), Bkm(128 * K), ref(128 * 128), got(128 * 128);
    srand(5);
    auto tf = [](float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; float y; memcpy(&y, &u, 4); return y; };
    for (auto& x : a) x = tf((rand() % 2001 - 1000) / 500.f);
    for (auto& x : b) x = tf((rand() % 2001 - 1000) / 500.f);
    for (int n = 0; n < 128; ++n) for (int k = 0; k < K; ++k) Bkm[((k / 4) * 128 + n) * 4 + k % 4] = b[n * K + k];
    float *dA, *dB, *dD; long long* dC;
    cudaMalloc(&dA, sizeof(float) * ROWS * K); cudaMalloc(&dB, sizeof(float) * 128 * K); cudaMalloc(&dD, sizeof(float) * 128 * 128); cudaMalloc(&dC, sizeof(long long));
    cudaMemcpy(dA, a.data(), sizeof(float) * ROWS * K, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, Bkm.data(), sizeof(float) * 128 * K, cudaMemcpyHostToDevice);
    const size_t smem = sizeof(float) * (ROWS * K + 128 * K) + 2048;
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int swz = 0; swz < 2; ++swz)
        for (int shift : {0, 1, 8, 9, 10})
            for (int reps : {1, 128}) {
                for (int m = 0; m < 128; ++m) for (int n = 0; n < 128; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)a[(m + shift) * K + k] * b[n * K + k]; ref[m * 128 + n] = (float)s; }
                cudaMemset(dD, 0xFF, sizeof(float) * 128 * 128);
                k_probe<<<1, 128, smem>>>(dA, dB, dD, dC, swz, shift, reps, 0);
                cudaError_t e = cudaDeviceSynchronize();
                long long cyc = 0;
                cudaMemcpy(&cyc, dC, sizeof(cyc), cudaMemcpyDeviceToHost);
                cudaMemcpy(got.data(), dD, sizeof(float) * 128 * 128, cudaMemcpyDeviceToHost);
                double err = 0, mx = 0;
                for (int i = 0; i < 128 * 128; ++i) { err = fmax(err, fabs(got[i] - reps * ref[i])); mx = fmax(mx, fabs(reps * ref[i])); }
                if (reps == 1) printf("%-14s shift %2d rows: cuda=%s max err %.3e (ref max %.1f)", swz ? "SWIZZLE_128B" : "no swizzle", shift, cudaGetErrorString(e), err, mx);
                else printf("   %.1f cycles per M128 N128 K8 instruction (err %.2e of %.0f)\n", (double)cyc / (reps * (K / 8)), err, mx);
                if (e != cudaSuccess) return 1;
            }
    for (int nmode = 0; nmode < 3; ++nmode) {
        k_probe<<<1, 128, smem>>>(dA, dB, dD, dC, 0, 9, 128, nmode);
        cudaDeviceSynchronize();
        long long cyc = 0;
        cudaMemcpy(&cyc, dC, sizeof(cyc), cudaMemcpyDeviceToHost);
        printf("per K step, shift 9, no swizzle: %-34s %.1f cycles\n", nmode == 0 ? "M128 N128" : nmode == 1 ? "M128 N64" : "M128 N128 + M128 N64 (conv kernels)", (double)cyc / (128 * (K / 8)));
    }
    return 0;
}
