// gemm_probe.cu — developer probe for gemm_tc.cuh (both operand major-nesses) against a CPU loop.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "gemm_tc.cuh"
using namespace n4j;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn enc() { void* p; cudaDriverEntryPointQueryResult q; cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q); return (EncodeTiledFn)p; }
static CUtensorMap imap(float* img, long long rows, long long nch2, int br, int bc) {
    CUtensorMap tm; const cuuint64_t d[3] = {4, (cuuint64_t)rows, (cuuint64_t)nch2}; const cuuint64_t s[2] = {16, (cuuint64_t)rows * 16};
    const cuuint32_t b[3] = {4, (cuuint32_t)br, (cuuint32_t)bc}; const cuuint32_t e[3] = {1, 1, 1};
    CUresult r = enc()(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, img, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) printf("encode failed %d\n", (int)r); return tm; }
static CUtensorMap wmap(float* img, long long nch, long long nb) {
    CUtensorMap tm; const cuuint64_t d[4] = {4, 256, (cuuint64_t)nch, (cuuint64_t)nb}; const cuuint64_t s[3] = {16, 4096, (cuuint64_t)nch * 4096};
    const cuuint32_t b[4] = {4, 256, 8, 1}; const cuuint32_t e[4] = {1, 1, 1, 1};
    CUresult r = enc()(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, img, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) printf("encode failed %d\n", (int)r); return tm; }

static int run(int Brows, int nin, int nout) {
    // x [B][nin], d [B][nout], W [nout][nin]
    std::vector<float> x((size_t)Brows * nin), d((size_t)Brows * nout), w((size_t)nout * nin);
    srand(7); for (auto& v : x) v = rand() / (float)RAND_MAX - 0.5f; for (auto& v : d) v = rand() / (float)RAND_MAX - 0.5f; for (auto& v : w) v = (rand() / (float)RAND_MAX - 0.5f) * 0.1f;
    const long long ir = (Brows + 127) / 128 * 128;
    float *dx, *dd, *dw, *ximg, *dimg, *wsf, *o1, *o2;
    CK(cudaMalloc(&dx, x.size() * 4)); CK(cudaMalloc(&dd, d.size() * 4)); CK(cudaMalloc(&dw, w.size() * 4));
    CK(cudaMalloc(&ximg, 2ll * nin * ir * 4)); CK(cudaMalloc(&dimg, 2ll * nout * ir * 4));
    const long long nbf = (nout + 127) / 128;
    CK(cudaMalloc(&wsf, nbf * (nin / 4) * 256 * 4 * 4)); CK(cudaMalloc(&o1, (size_t)Brows * nout * 4)); CK(cudaMalloc(&o2, (size_t)nout * nin * 4));
    CK(cudaMemset(ximg, 0, 2ll * nin * ir * 4)); CK(cudaMemset(dimg, 0, 2ll * nout * ir * 4)); CK(cudaMemset(wsf, 0, nbf * (nin / 4) * 256 * 16));
    CK(cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dd, d.data(), d.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dw, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
    k_rowmajor_to_image<<<148, 256>>>(fixed_ref(dx), 1.f, ximg, Brows, nin, ir);
    k_rowmajor_to_image<<<148, 256>>>(fixed_ref(dd), 1.f, dimg, Brows, nout, ir);
    k_weights_to_stacked_image<<<148, 256>>>(dw, wsf, nin, nout, 0);
    CK(cudaDeviceSynchronize());
    CK(cudaFuncSetAttribute(k_gemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_TC_SMEM));
    // forward: y[b,o] = sum_i x[b,i] W[o,i]
    { GemmTcArgs a{}; a.M = Brows; a.N = nout; a.kblocks = nin / 32; a.a_nch = nin / 4; a.out = fixed_ref(o1);
      CUtensorMap ta = imap(ximg, ir, 2 * (nin / 4), 128, 8), tb = wmap(wsf, nin / 4, nbf);
      k_gemm_tc<<<((a.M + 127) / 128) * ((a.N + 127) / 128), GEMM_TC_THREADS, GEMM_TC_SMEM>>>(ta, tb, a); CK(cudaDeviceSynchronize()); }
    // wgrad: dWt[o,i] = sum_b d[b,o] x[b,i]
    { const long long ocp = (nout + 127) / 128 * 128, nbb = (nin + 127) / 128;
      float *dT, *xT; CK(cudaMalloc(&dT, 2 * ir * ocp * 4)); CK(cudaMalloc(&xT, nbb * (ir / 4) * 256 * 16));
      CK(cudaMemset(dT, 0, 2 * ir * ocp * 4)); CK(cudaMemset(xT, 0, nbb * (ir / 4) * 256 * 16));
      k_rowmajor_to_image_T<<<148, 256>>>(fixed_ref(dd), dT, Brows, nout, ir, ocp, 0);
      k_rowmajor_to_image_T<<<148, 256>>>(fixed_ref(dx), xT, Brows, nin, ir, 0, 1);
      GemmTcArgs a{}; a.M = nout; a.N = nin; a.kblocks = (int)(ir / 32); a.a_nch = (int)(ir / 4); a.out = fixed_ref(o2);
      CUtensorMap ta = imap(dT, ocp, 2 * (ir / 4), 128, 8), tb = wmap(xT, ir / 4, nbb);
      k_gemm_tc<<<((a.M + 127) / 128) * ((a.N + 127) / 128), GEMM_TC_THREADS, GEMM_TC_SMEM>>>(ta, tb, a); CK(cudaDeviceSynchronize()); }
    std::vector<float> y((size_t)Brows * nout), g((size_t)nout * nin);
    CK(cudaMemcpy(y.data(), o1, y.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(g.data(), o2, g.size() * 4, cudaMemcpyDeviceToHost));
    double e1 = 0, m1 = 0, e2 = 0, m2 = 0;
    for (int b = 0; b < Brows; ++b) for (int o = 0; o < nout; ++o) { double acc = 0; for (int i = 0; i < nin; ++i) acc += (double)x[(size_t)b * nin + i] * w[(size_t)o * nin + i];
        e1 = std::max(e1, std::fabs(acc - y[(size_t)b * nout + o])); m1 = std::max(m1, std::fabs(acc)); }
    for (int o = 0; o < nout; ++o) for (int i = 0; i < nin; ++i) { double acc = 0; for (int b = 0; b < Brows; ++b) acc += (double)d[(size_t)b * nout + o] * x[(size_t)b * nin + i];
        e2 = std::max(e2, std::fabs(acc - g[(size_t)o * nin + i])); m2 = std::max(m2, std::fabs(acc)); }
    printf("B=%d nin=%d nout=%d: fwd(K-major) err %.3e / max %.3f   wgrad(MN-major) err %.3e / max %.3f   sample wgrad got %.5f %.5f %.5f\n", Brows, nin, nout, e1, m1, e2, m2, g[0], g[1], g[nin]);
    return (e1 / m1 < 1e-5 && e2 / m2 < 1e-5) ? 0 : 1;
}
int main() { int rc = 0; rc |= run(128, 128, 128); rc |= run(96, 320, 320); rc |= run(1024, 512, 256); printf(rc ? "GEMM PROBE FAILED\n" : "GEMM PROBE OK\n"); return rc; }
