// tc_probe.cu — developer probe: runs the tensor-core conv kernel on random data and checks it against a CPU loop.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scratch/tc_probe neuralode4j_b200/csrc/tc_probe.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../neuralode4j_b200/csrc/conv_tc.cuh"
#include "solver_kernels.cuh"

using namespace n4j;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

static int run(int B, int H, int W, int mode) {
    const int C = 64;
    ConvGeom g = make_conv_geom(B, H, W);
    const long long rows = (long long)B * H * W;
    std::vector<float> x(rows * C), w(C * C * 9), ref(rows * C), out(rows * C);
    srand(1234 + B);
    for (auto& v : x) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto& v : w) v = ((float)rand() / RAND_MAX * 2.f - 1.f) / 24.f;
    // CPU reference (double): mode 0 forward y[p,co] = sum x[p+s,ci] W[co,ci,tap]; mode 1 dgrad dx[p,ci] = sum g[p-s,co] W[co,ci,tap]
    for (int b = 0; b < B; ++b)
        for (int h = 0; h < H; ++h)
            for (int ww = 0; ww < W; ++ww)
                for (int n = 0; n < C; ++n) {
                    double acc = 0;
                    for (int tap = 0; tap < 9; ++tap) {
                        int dh = tap / 3 - 1, dw = tap % 3 - 1;
                        if (mode == 1) { dh = -dh; dw = -dw; }
                        const int h2 = h + dh, w2 = ww + dw;
                        if (h2 < 0 || h2 >= H || w2 < 0 || w2 >= W) continue;
                        const float* xr = &x[(((long long)b * H + h2) * W + w2) * C];
                        for (int k = 0; k < C; ++k) {
                            const float wv = mode == 0 ? w[(n * C + k) * 9 + tap] : w[(k * C + n) * 9 + tap];
                            acc += (double)xr[k] * (double)wv;
                        }
                    }
                    ref[(((long long)b * H + h) * W + ww) * C + n] = (float)acc;
                }
    float *dx, *dw, *dimg, *dwimg, *dout;
    CK(cudaMalloc(&dx, x.size() * 4)); CK(cudaMalloc(&dw, w.size() * 4)); CK(cudaMalloc(&dout, out.size() * 4));
    CK(cudaMalloc(&dimg, conv_image_floats(g) * 4)); CK(cudaMalloc(&dwimg, 9 * 2 * 4096 * 4));
    CK(cudaMemset(dimg, 0, conv_image_floats(g) * 4));
    CK(cudaMemset(dout, 0xff, out.size() * 4));
    CK(cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dw, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
    k_compact_to_image<<<148, 256>>>(fixed_ref(dx), g, dimg);
    k_conv_w_to_tc<<<64, 256>>>(dw, dwimg, mode);
    CK(cudaDeviceSynchronize());
    ConvTcArgs a{};
    a.img = dimg; a.wimg = dwimg; a.out = fixed_ref(dout); a.g = g; a.act = 0; a.fuse_stats = 0;
    const size_t smem = conv_tc_smem_bytes(g);
    CK(cudaFuncSetAttribute(k_conv3x3_tc<3, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_conv3x3_tc<1, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_conv3x3_tc<0, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_conv3x3_tc<3, 0, 0><<<g.tiles, CONV_TC_THREADS, smem>>>(a);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0;
    long long bad = 0;
    for (size_t i = 0; i < out.size(); ++i) {
        const double e = std::fabs((double)out[i] - (double)ref[i]);
        if (!(e == e)) { ++bad; continue; }
        maxerr = std::max(maxerr, e); maxref = std::max(maxref, (double)std::fabs(ref[i]));
    }
    // timing
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 100;
    float ms = 0, ms1 = 0, ms0 = 0, msE = 0;
    for (int i = 0; i < 5; ++i) k_conv3x3_tc<3, 0, 0><<<g.tiles, CONV_TC_THREADS, smem>>>(a);
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) k_conv3x3_tc<3, 0, 0><<<g.tiles, CONV_TC_THREADS, smem>>>(a);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) k_conv3x3_tc<1, 0, 0><<<g.tiles, CONV_TC_THREADS, smem>>>(a);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms1, e0, e1);
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) k_conv3x3_tc<0, 0, 0><<<g.tiles, CONV_TC_THREADS, smem>>>(a);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms0, e0, e1);
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) k_fill<<<1, 32>>>(dout, 0.f, 1);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&msE, e0, e1);
    printf("   3-pass %.2f us, 1-pass %.2f us, no-mma %.2f us, empty kernel %.2f us\n", ms * 10.f, ms1 * 10.f, ms0 * 10.f, msE * 10.f);
    printf("B=%d H=%d W=%d mode=%d tiles=%d smem=%zu : max|err|=%.3e (max|ref|=%.3f, rel %.2e) nan=%lld  %.2f us/launch\n", B, H, W, mode, g.tiles, smem,
           maxerr, maxref, maxerr / maxref, bad, ms * 1000.f / iters);
    cudaFree(dx); cudaFree(dw); cudaFree(dimg); cudaFree(dwimg); cudaFree(dout);
    return (bad == 0 && maxerr / maxref < 2e-6) ? 0 : 2;
}

int main() {
    int rc = 0;
    rc |= run(3, 7, 7, 0);
    rc |= run(128, 7, 7, 0);
    rc |= run(128, 7, 7, 1);
    rc |= run(16, 6, 6, 1);
    printf(rc == 0 ? "TC PROBE OK\n" : "TC PROBE FAILED\n");
    return rc;
}
