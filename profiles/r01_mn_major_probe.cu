// MN-major tf32 operand probe for tcgen05.mma (SWIZZLE_NONE): D[128x64] = A[128xK] * B[64xK]^T with A stored MN-major.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../neuralode4j_b200/csrc/tc_common.cuh"
using namespace n4j;

// A (MN-major, no swizzle): [mchunk = m/4 (32)][k (K)][4]   -> element (m,k) at ((m/4)*K + k)*4 + m%4
// B (K-major, no swizzle):  [kchunk = k/4][n (64)][4]        -> element (n,k) at ((k/4)*64 + n)*4 + k%4   (as the conv kernel)
template <int K>
__global__ void k_probe(const float* A, const float* B, float* D, uint32_t lbo, uint32_t sbo, int a_major) {
    extern __shared__ __align__(128) uint8_t sm[];
    float* sA = reinterpret_cast<float*>(sm);
    float* sB = sA + 128 * K;
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 128 * K; i += blockDim.x) sA[i] = A[i];
    for (int i = threadIdx.x; i < 64 * K; i += blockDim.x) sB[i] = B[i];
    if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
    if (threadIdx.x < 32) tc::tmem_alloc<64>(&slot);
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = tc::idesc_tf32(128, 64, a_major, 0);
        for (int ks = 0; ks < K / 8; ++ks) {
            // A: k-step advances 8 k rows = 128 bytes inside every m-chunk
            const uint64_t da = a_major ? tc::smem_desc(tc::smem_u32(sA) + ks * 128, lbo, sbo)
                                        : tc::smem_desc(tc::smem_u32(sA) + ks * 2 * 128 * 16, 128 * 16, 128);
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
    constexpr int K = 16;
    std::vector<float> a(128 * K), b(64 * K), Amn(128 * K), Akm(128 * K), Bk(64 * K), ref(128 * 64), got(128 * 64);
    srand(1);
    auto tf = [](float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; float y; memcpy(&y, &u, 4); return y; };
    for (auto& x : a) x = tf((rand() % 2001 - 1000) / 500.f);
    for (auto& x : b) x = tf((rand() % 2001 - 1000) / 500.f);
    for (int m = 0; m < 128; ++m) for (int k = 0; k < K; ++k) {
        Amn[((m / 4) * K + k) * 4 + m % 4] = a[m * K + k];
        Akm[((k / 4) * 128 + m) * 4 + k % 4] = a[m * K + k];
    }
    for (int n = 0; n < 64; ++n) for (int k = 0; k < K; ++k) Bk[((k / 4) * 64 + n) * 4 + k % 4] = b[n * K + k];
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)a[m * K + k] * b[n * K + k]; ref[m * 64 + n] = (float)s; }
    float *dA, *dB, *dD;
    cudaMalloc(&dA, sizeof(float) * 128 * K); cudaMalloc(&dB, sizeof(float) * 64 * K); cudaMalloc(&dD, sizeof(float) * 128 * 64);
    cudaMemcpy(dB, Bk.data(), sizeof(float) * 64 * K, cudaMemcpyHostToDevice);
    const size_t smem = sizeof(float) * (128 + 64) * K + 256;
    struct V { const char* name; int amaj; uint32_t lbo, sbo; };
    const uint32_t chunk = K * 16;     // bytes of one m-chunk (K rows of 16 B)
    V vs[] = {{"K-major control", 0, 0, 0}, {"MN lbo=128 sbo=chunk", 1, 128, chunk}, {"MN lbo=chunk sbo=128", 1, chunk, 128},
              {"MN lbo=chunk sbo=chunk", 1, chunk, chunk}, {"MN lbo=128 sbo=128", 1, 128, 128}, {"MN lbo=16 sbo=chunk", 1, 16, chunk},
              {"MN lbo=chunk sbo=16", 1, chunk, 16}};
    for (auto& v : vs) {
        cudaMemcpy(dA, v.amaj ? Amn.data() : Akm.data(), sizeof(float) * 128 * K, cudaMemcpyHostToDevice);
        cudaMemset(dD, 0xFF, sizeof(float) * 128 * 64);
        k_probe<K><<<1, 128, smem>>>(dA, dB, dD, v.lbo, v.sbo, v.amaj);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(got.data(), dD, sizeof(float) * 128 * 64, cudaMemcpyDeviceToHost);
        double err = 0, mx = 0; int nz = 0;
        for (int i = 0; i < 128 * 64; ++i) { err = fmax(err, fabs(got[i] - ref[i])); mx = fmax(mx, fabs(ref[i])); nz += got[i] != 0.f; }
        printf("%-26s cuda=%s max err %.4e (ref max %.3f) nonzero %d  sample %.4f vs %.4f | %.4f vs %.4f\n", v.name, cudaGetErrorString(e), err, mx, nz,
               got[0], ref[0], got[5 * 64 + 7], ref[5 * 64 + 7]);
        if (e != cudaSuccess) break;
    }
    return 0;
}
