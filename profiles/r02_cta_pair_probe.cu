// r02 probe: tcgen05.mma.cta_group::2 (CTA pair, M = 256) for kind::tf32 with SWIZZLE_NONE K-major operands — correctness and cycles per
// instruction against cta_group::1 (M = 128), the form the conv kernels use.  Question: how much shorter does the conv's MMA phase get when
// two CTAs share every weight tile (each keeps half of B in its shared memory: 6 KB instead of 8 KB of shared-memory reads per M128·N128·K8
// per SM, half the weight ingest per SM)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o scratch/bin/pair_probe profiles/r02_cta_pair_probe.cu && scratch/bin/pair_probe
// D[256 x 128] = A[256 x K] * B[128 x K]^T.  CTA r of the pair stages A rows [128 r, 128 r + 128) and B rows [64 r, 64 r + 64).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../neuralode4j_b200/csrc/tc_common.cuh"
using namespace n4j;

constexpr int K = 64;        // 8 K steps

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
    asm volatile(
        "{\n\t.reg .b32 r;\n\t"
        "mapa.shared::cluster.u32 r, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [r];\n\t}" ::"r"(tc::smem_u32(bar)), "r"(cta)
        : "memory");
}
__device__ __forceinline__ void mma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void mma_commit_pair(uint64_t* bar) {      // arrives on `bar` (same offset) in both CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(tc::smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* slot) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(slot)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}

// pair != 0: one cta_group::2 stream issued by the leader; pair == 0: every CTA runs its own cta_group::1 stream on A[its 128 rows] x all of B
__global__ void __cluster_dims__(2, 1, 1) k_probe(const float* A_km /*[2][K/4][128][4]*/, const float* B_km /*[K/4][128][4]*/, float* D, long long* cycles,
                                                   int pair, int reps) {
    extern __shared__ __align__(128) uint8_t sm[];
    float* sA = reinterpret_cast<float*>(sm);           // [K/4][128][4]
    float* sB = sA + 128 * K;                           // pair: [K/4][64][4] (this CTA's half); single: [K/4][128][4]
    __shared__ uint64_t ready, done;
    __shared__ uint32_t slot;
    const uint32_t rank = cluster_ctarank();
    for (int i = threadIdx.x; i < 128 * K; i += blockDim.x) sA[i] = A_km[rank * 128 * K + i];
    if (pair) {
        for (int i = threadIdx.x; i < 64 * K; i += blockDim.x) {
            const int k4 = i / 256, r = i % 256;                    // chunk, (row, 4 floats)
            sB[i] = B_km[k4 * 512 + rank * 256 + r];
        }
    } else {
        for (int i = threadIdx.x; i < 128 * K; i += blockDim.x) sB[i] = B_km[i];
    }
    if (threadIdx.x == 0) { tc::mbar_init(&ready, 2); tc::mbar_init(&done, 1); tc::fence_barrier_init(); }
    if (threadIdx.x < 32) { if (pair) tmem_alloc_pair<128>(&slot); else tc::tmem_alloc<128>(&slot); }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    cluster_sync();                                     // both CTAs: barriers initialised, operands staged, TMEM allocated
    tc::tc_fence_after();
    const uint32_t tmem = slot;
    long long t0 = 0;
    if (pair) {
        if (threadIdx.x == 0) mbar_arrive_remote(&ready, 0);        // "my half is staged" -> the leader's barrier
        if (rank == 0 && threadIdx.x == 0) {
            tc::mbar_wait(&ready, 0);
            tc::tc_fence_after();
            const uint32_t idesc = tc::idesc_tf32(256, 128, 0, 0);
            t0 = clock64();
            for (int r = 0; r < reps; ++r)
                for (int ks = 0; ks < K / 8; ++ks) {
                    const uint64_t da = tc::smem_desc(tc::smem_u32(sA) + ks * 2 * 128 * 16, 128 * 16, 128);
                    const uint64_t db = tc::smem_desc(tc::smem_u32(sB) + ks * 2 * 64 * 16, 64 * 16, 128);
                    mma_tf32_pair(tmem, da, db, idesc, (r | ks) ? 1u : 0u);
                }
            mma_commit_pair(&done);
        }
    } else if (threadIdx.x == 0) {
        const uint32_t idesc = tc::idesc_tf32(128, 128, 0, 0);
        t0 = clock64();
        for (int r = 0; r < reps; ++r)
            for (int ks = 0; ks < K / 8; ++ks) {
                const uint64_t da = tc::smem_desc(tc::smem_u32(sA) + ks * 2 * 128 * 16, 128 * 16, 128);
                const uint64_t db = tc::smem_desc(tc::smem_u32(sB) + ks * 2 * 128 * 16, 128 * 16, 128);
                tc::mma_tf32(tmem, da, db, idesc, (r | ks) ? 1u : 0u);
            }
        tc::mma_commit(&done);
    }
    __syncthreads();
    tc::mbar_wait(&done, 0);
    if (threadIdx.x == 0 && (rank == 0 || !pair)) cycles[rank] = clock64() - t0;
    tc::tc_fence_after();
    {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int q = 0; q < 4; ++q) {
            float v[32];
            tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + q * 32, v);
            for (int j = 0; j < 32; ++j) D[(rank * 128 + warp * 32 + lane) * 128 + q * 32 + j] = v[j];
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    cluster_sync();
    if (threadIdx.x < 32) { if (pair) tmem_dealloc_pair<128>(tmem); else tc::tmem_dealloc<128>(tmem); }
}

int main() {
    std::vector<float> a(256 * K), b(128 * K), Akm(256 * K), Bkm(128 * K), ref(256 * 128), got(256 * 128);
    srand(3);
    auto tf = [](float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; float y; memcpy(&y, &u, 4); return y; };
    for (auto& x : a) x = tf((rand() % 2001 - 1000) / 500.f);
    for (auto& x : b) x = tf((rand() % 2001 - 1000) / 500.f);
    for (int m = 0; m < 256; ++m) for (int k = 0; k < K; ++k) Akm[(m / 128) * 128 * K + ((k / 4) * 128 + m % 128) * 4 + k % 4] = a[m * K + k];
    for (int n = 0; n < 128; ++n) for (int k = 0; k < K; ++k) Bkm[((k / 4) * 128 + n) * 4 + k % 4] = b[n * K + k];
    for (int m = 0; m < 256; ++m) for (int n = 0; n < 128; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)a[m * K + k] * b[n * K + k]; ref[m * 128 + n] = (float)s; }
    float *dA, *dB, *dD; long long* dC;
    cudaMalloc(&dA, sizeof(float) * 256 * K); cudaMalloc(&dB, sizeof(float) * 128 * K); cudaMalloc(&dD, sizeof(float) * 256 * 128); cudaMalloc(&dC, 2 * sizeof(long long));
    cudaMemcpy(dA, Akm.data(), sizeof(float) * 256 * K, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, Bkm.data(), sizeof(float) * 128 * K, cudaMemcpyHostToDevice);
    const size_t smem = sizeof(float) * 2 * 128 * K + 256;
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int pair = 0; pair < 2; ++pair)
        for (int reps : {1, 64}) {
            cudaMemset(dD, 0xFF, sizeof(float) * 256 * 128);
            cudaMemset(dC, 0, 2 * sizeof(long long));
            k_probe<<<2, 128, smem>>>(dA, dB, dD, dC, pair, reps);
            cudaError_t e = cudaDeviceSynchronize();
            long long cyc[2] = {0, 0};
            cudaMemcpy(cyc, dC, sizeof(cyc), cudaMemcpyDeviceToHost);
            cudaMemcpy(got.data(), dD, sizeof(float) * 256 * 128, cudaMemcpyDeviceToHost);
            double err = 0, mx = 0;
            for (int i = 0; i < 256 * 128; ++i) { err = fmax(err, fabs(got[i] - reps * ref[i])); mx = fmax(mx, fabs(reps * ref[i])); }
            printf("%s reps %3d: cuda=%s max err %.3e (ref max %.1f), %lld cycles = %.1f per K step of 256 rows x N128 (%s)\n", pair ? "cta_group::2 M256" : "cta_group::1 M128 x2 CTAs",
                   reps, cudaGetErrorString(e), err, mx, cyc[0], (double)cyc[0] / (reps * (K / 8)), pair ? "one instruction for the pair" : "one instruction per CTA, both CTAs in parallel");
            if (e != cudaSuccess) return 1;
        }
    return 0;
}
