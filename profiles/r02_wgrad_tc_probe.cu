// r02: standalone check + timing of k_conv_wgrad_tc (neuralode4j_b200/csrc/conv_wgrad_tc.cuh) against a double-precision CPU loop.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -lineinfo -o scratch/bin/wgrad_probe profiles/r02_wgrad_tc_probe.cu
//   scratch/bin/wgrad_probe [B H W]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../neuralode4j_b200/csrc/conv_wgrad_tc.cuh"
using namespace n4j;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

static int run(int B, int H, int W, int iters) {
    const long long npix = (long long)B * H * W;
    std::vector<float> x(npix * 64), g(npix * 64);
    srand(7);
    for (auto& v : x) { v = (rand() % 20001 - 10000) / 7919.f; if (rand() % 3 == 0) v = 0.f; }     // ReLU-like sparsity
    for (auto& v : g) v = (rand() % 20001 - 10000) / 104729.f;
    std::vector<double> ref(9 * 4096, 0.0);
    for (int b = 0; b < B; ++b)
        for (int h = 0; h < H; ++h)
            for (int w = 0; w < W; ++w)
                for (int t = 0; t < 9; ++t) {
                    const int hh = h + t / 3 - 1, ww = w + t % 3 - 1;
                    if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
                    const float* xr = &x[(((long long)b * H + hh) * W + ww) * 64];
                    const float* gr = &g[(((long long)b * H + h) * W + w) * 64];
                    double* d = &ref[t * 4096];
                    for (int ci = 0; ci < 64; ++ci) {
                        const double xv = xr[ci];
                        if (xv == 0.0) continue;
                        for (int co = 0; co < 64; ++co) d[ci * 64 + co] += xv * (double)gr[co];
                    }
                }
    const WgradTcGeom geom = make_wgrad_tc_geom(B, H, W);
    float *dx, *dg, *dpart;
    CK(cudaMalloc(&dx, sizeof(float) * x.size()));
    CK(cudaMalloc(&dg, sizeof(float) * g.size()));
    CK(cudaMalloc(&dpart, sizeof(float) * (size_t)geom.tiles * 9 * 4096));
    CK(cudaMemcpy(dx, x.data(), sizeof(float) * x.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dg, g.data(), sizeof(float) * g.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dpart, 0xFF, sizeof(float) * (size_t)geom.tiles * 9 * 4096));
    const size_t smem = conv_wgrad_tc_smem_bytes(geom);
    CK(cudaFuncSetAttribute(k_conv_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const dim3 grid(geom.tiles);
    k_conv_wgrad_tc<<<grid, WGRAD_TC_THREADS, smem>>>(fixed_ref(dx), fixed_ref(dg), geom, dpart, (long long*)nullptr);
    CK(cudaDeviceSynchronize());
    std::vector<float> part((size_t)geom.tiles * 9 * 4096);
    CK(cudaMemcpy(part.data(), dpart, sizeof(float) * part.size(), cudaMemcpyDeviceToHost));
    double err = 0, mx = 0;
    for (int i = 0; i < 9 * 4096; ++i) {
        double s = 0;
        for (int t = 0; t < geom.tiles; ++t) s += part[(size_t)t * 9 * 4096 + i];
        err = fmax(err, fabs(s - ref[i]));
        mx = fmax(mx, fabs(ref[i]));
    }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 5; ++i) k_conv_wgrad_tc<<<grid, WGRAD_TC_THREADS, smem>>>(fixed_ref(dx), fixed_ref(dg), geom, dpart, (long long*)nullptr);
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) k_conv_wgrad_tc<<<grid, WGRAD_TC_THREADS, smem>>>(fixed_ref(dx), fixed_ref(dg), geom, dpart, (long long*)nullptr);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("B=%d H=%d W=%d tiles=%d grid=%d smem=%zu: max err %.3e (ref max %.3f, rel %.2e), %.2f us per launch\n", B, H, W, geom.tiles, geom.tiles,
           smem, err, mx, err / mx, 1e3 * ms / iters);
    {
        long long* ddbg; long long h[32] = {0};
        cudaMalloc(&ddbg, sizeof(h)); cudaMemset(ddbg, 0, sizeof(h));
        k_conv_wgrad_tc<<<grid, WGRAD_TC_THREADS, smem>>>(fixed_ref(dx), fixed_ref(dg), geom, dpart, ddbg);
        cudaDeviceSynchronize();
        cudaMemcpy(h, ddbg, sizeof(h), cudaMemcpyDeviceToHost);
        printf("  cycles since entry: wait-done %lld | staged %lld | synced %lld | mma issued", h[1] - h[0], h[2] - h[0], h[3] - h[0]);
        for (int i = 0; i < 4; ++i) printf(" %lld", h[4 + i] - h[0]);
        printf(" | tap full");
        for (int i = 0; i < 4; ++i) printf(" %lld", h[10 + i] - h[0]);
        printf(" | exit %lld\n", h[30] - h[0]);
        cudaFree(ddbg);
    }
    cudaFree(dx); cudaFree(dg); cudaFree(dpart);
    return err / mx < 2e-5 ? 0 : 2;
}

int main(int argc, char** argv) {
    if (argc == 4) return run(atoi(argv[1]), atoi(argv[2]), atoi(argv[3]), 50);
    int rc = 0;
    rc |= run(2, 7, 7, 5);
    rc |= run(128, 7, 7, 100);
    rc |= run(3, 5, 6, 5);
    rc |= run(8, 6, 6, 5);
    rc |= run(16, 14, 14, 20);
    printf(rc ? "FAILED\n" : "ALL OK\n");
    return rc;
}
