// PREPARED IN r01, NOT RUN YET — second experiment of r02:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o gpurun_out/atmem_probe profiles/r02_a_in_tmem_probe.cu && gpurun_out/atmem_probe
//
// Question: how much faster does a kind::tf32 M128 x N128 x K8 MMA stream run when the A operand sits in TENSOR MEMORY instead of
// shared memory?  With SWIZZLE_NONE both operands cost 4 KB of shared-memory reads per instruction (the conv's
// MMA phase is bound by those reads, 116 cycles per K step against 102 tensor cycles).  Swapping the conv's operand roles — stacked
// weights [W_hi; W_lo] as the M = 128 operand in TMEM, pixels as N from the staged tile — would leave 8 KB instead of 14 KB per K step.
// Part 1 checks D[128 x 128] = A[128 x K] * B[128 x K]^T with A written to TMEM by tcgen05.st (row m -> lane m, k -> column k);
// part 2 times REPS back-to-back K-step streams for A-in-smem and A-in-TMEM with clock64 (one CTA; run under `gpurun`).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../neuralode4j_b200/csrc/tc_common.cuh"
using namespace n4j;

constexpr int K = 64;        // 8 K steps
constexpr int REPS = 64;

__device__ __forceinline__ void mma_tf32_a_tmem(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 8 columns from registers (thread t of the warp owns lane base + t)
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(__float_as_uint(v[0])),
                 "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])),
                 "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}

// A_km / B_km: K-major no-swizzle images [k/4][row][4]; A_rm: row-major [m][K] (for the TMEM copy)
__global__ void k_probe(const float* A_km, const float* A_rm, const float* B_km, float* D, long long* cycles, int a_in_tmem, int reps) {
    extern __shared__ __align__(128) uint8_t sm[];
    float* sA = reinterpret_cast<float*>(sm);
    float* sB = sA + 128 * K;
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 128 * K; i += blockDim.x) { sA[i] = A_km[i]; sB[i] = B_km[i]; }
    if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
    if (threadIdx.x < 32) tc::tmem_alloc<256>(&slot);        // columns [0,128): D, [128,128+K): A
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = slot;
    const uint32_t tmem_a = tmem + 128;
    {   // every warp writes its 32 lanes of A: row m = warp * 32 + lane
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, m = warp * 32 + lane;
        for (int k0 = 0; k0 < K; k0 += 8) {
            float v[8];
            for (int j = 0; j < 8; ++j) v[j] = A_rm[m * K + k0 + j];
            tmem_st8(tmem_a + ((uint32_t)(warp * 32) << 16) + k0, v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    long long t0 = 0;
    if (threadIdx.x == 0) {
        const uint32_t idesc = tc::idesc_tf32(128, 128, 0, 0);
        t0 = clock64();
        for (int r = 0; r < reps; ++r)
            for (int ks = 0; ks < K / 8; ++ks) {
                const uint64_t db = tc::smem_desc(tc::smem_u32(sB) + ks * 2 * 128 * 16, 128 * 16, 128);
                if (a_in_tmem) {
                    mma_tf32_a_tmem(tmem, tmem_a + ks * 8, db, idesc, (r | ks) ? 1u : 0u);
                } else {
                    const uint64_t da = tc::smem_desc(tc::smem_u32(sA) + ks * 2 * 128 * 16, 128 * 16, 128);
                    tc::mma_tf32(tmem, da, db, idesc, (r | ks) ? 1u : 0u);
                }
            }
        tc::mma_commit(&bar);
    }
    __syncthreads();
    if (threadIdx.x < 128) {
        tc::mbar_wait(&bar, 0);
        if (threadIdx.x == 0) *cycles = clock64() - t0;
        tc::tc_fence_after();
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int q = 0; q < 4; ++q) {
            float v[32];
            tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + q * 32, v);
            for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * 128 + q * 32 + j] = v[j];
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tc::tmem_dealloc<256>(slot);
}

int main() {
    std::vector<float> a(128 * K), b(128 * K), Akm(128 * K), Bkm(128 * K), ref(128 * 128), got(128 * 128);
    srand(2);
    auto tf = [](float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; float y; memcpy(&y, &u, 4); return y; };
    for (auto& x : a) x = tf((rand() % 2001 - 1000) / 500.f);
    for (auto& x : b) x = tf((rand() % 2001 - 1000) / 500.f);
    for (int m = 0; m < 128; ++m) for (int k = 0; k < K; ++k) { Akm[((k / 4) * 128 + m) * 4 + k % 4] = a[m * K + k]; Bkm[((k / 4) * 128 + m) * 4 + k % 4] = b[m * K + k]; }
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 128; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)a[m * K + k] * b[n * K + k]; ref[m * 128 + n] = (float)s; }
    float *dAk, *dAr, *dB, *dD; long long* dC;
    cudaMalloc(&dAk, sizeof(float) * 128 * K); cudaMalloc(&dAr, sizeof(float) * 128 * K); cudaMalloc(&dB, sizeof(float) * 128 * K);
    cudaMalloc(&dD, sizeof(float) * 128 * 128); cudaMalloc(&dC, sizeof(long long));
    cudaMemcpy(dAk, Akm.data(), sizeof(float) * 128 * K, cudaMemcpyHostToDevice);
    cudaMemcpy(dAr, a.data(), sizeof(float) * 128 * K, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, Bkm.data(), sizeof(float) * 128 * K, cudaMemcpyHostToDevice);
    const size_t smem = sizeof(float) * 2 * 128 * K + 256;
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int mode = 0; mode < 2; ++mode)
        for (int reps : {1, REPS}) {
            cudaMemset(dD, 0xFF, sizeof(float) * 128 * 128);
            k_probe<<<1, 128, smem>>>(dAk, dAr, dB, dD, dC, mode, reps);
            cudaError_t e = cudaDeviceSynchronize();
            long long cyc = 0;
            cudaMemcpy(&cyc, dC, sizeof(cyc), cudaMemcpyDeviceToHost);
            cudaMemcpy(got.data(), dD, sizeof(float) * 128 * 128, cudaMemcpyDeviceToHost);
            double err = 0, mx = 0;
            for (int i = 0; i < 128 * 128; ++i) { err = fmax(err, fabs(got[i] - reps * ref[i])); mx = fmax(mx, fabs(reps * ref[i])); }
            printf("A in %-5s reps %3d: cuda=%s max err %.3e (ref max %.1f), %lld cycles = %.1f per M128 N128 K8 instruction\n", mode ? "TMEM" : "smem", reps,
                   cudaGetErrorString(e), err, mx, cyc, (double)cyc / (reps * (K / 8)));
            if (e != cudaSuccess) return 1;
        }
    return 0;
}
