// common.cuh — shared device/host definitions of the node4j B200 engine.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/node4j.h"

namespace n4j {

// ---------------------------------------------------------------------------------------------------
// Device-resident control block of one solve.  The step controller lives on the GPU: kernels read the
// current step size / buffer rotation from here and the last block of the error-norm kernel updates it,
// so a whole adaptive solve runs without a host round trip (replaces the host reads at
// AdaptiveRungeKuttaSolver.java:78,170 of the reference).
// ---------------------------------------------------------------------------------------------------
struct Ctrl {
    // --- solve state -----------------------------------------------------------------------------
    float t;          // equation.time()
    float h;          // step to try next (already clamped to the end time when is_last)
    float t_end;      // t[1]
    float t_start;    // t[0]
    int backward;     // t.argMax()==0
    int y_cur;        // which of the two state buffers holds y (the other one is yWorking)
    int krow[7];      // physical row of logical stage j in K[7][N]  (FSAL = swap of two indices)
    int krow_prev[7]; // mapping that was valid during the last accepted step (dense output reads it)
    int done;
    int is_last;
    int status;       // 0 ok, N4J_ERR_ILLEGAL_STATE on NaN, N4J_ERR_MAX_STEPS on guard
    float err;
    // --- last accepted step, for the dense-output kernel ------------------------------------------
    int accepted_now; // the trial that just finished was accepted
    float t_prev;     // time before that step
    float h_used;
    int interp_first, interp_cnt;
    int interp_alias; // first accepted step: y_prev aliases the solver state (InterpolatingMultiStepSolver.java:41)
    int steps_this_solve;
    // --- init-step scratch ---------------------------------------------------------------------------
    float init_D;
    // --- reductions (accumulators: separate strided buffer, see ACC_STRIDE) --------------------------------
    double* red_acc;
    unsigned int ticket[4];
    // --- counters (n4j_stats) ----------------------------------------------------------------------------
    long long nfe, accepted, rejected, solves, trials, trials_this_solve;
    int log_n;
    int seg_index;    // which (t0,t1) pair of the time-pair table the next solve integrates
};

struct SolverParams {
    float atol, rtol, hmin, hmax, safety, max_growth, min_reduction;
    int order;
    unsigned flags;
    long long n_norm;      // N of the reference's mean (includes the BatchNorm-stat pad of the adjoint state)
    long long max_trials;
    int log_cap;
};

// Indirect tensor reference: the address is resolved on the device from the control block, which is how
// accept (y <- yWorking) and FSAL (K0 <- K6) cost zero bytes.
struct Ref {
    float* base;
    const int* sel;     // nullptr -> index 0
    long long stride;   // elements between alternatives
    long long off;      // element offset inside the selected alternative
    int flip;           // use (index ^ 1)
};

__device__ __forceinline__ float* resolve(const Ref& r) {
    int i = r.sel ? *r.sel : 0;
    if (r.flip) i ^= 1;
    return r.base + (long long)i * r.stride + r.off;
}

inline Ref fixed_ref(float* p, long long off = 0) { return Ref{p, nullptr, 0, off, 0}; }

// ---------------------------------------------------------------------------------------------------
// Reduction helpers (warp shuffle + block reduce in double)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum NV doubles over the block; result valid in thread 0.  blockDim.x must be a multiple of 32, <= 1024.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV]) {
    __shared__ double sm[NV][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double s = warp_sum(v[k]);
        if (lane == 0) sm[k][warp] = s;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double s = lane < nw ? sm[k][lane] : 0.0;
            s = warp_sum(s);
            v[k] = s;
        }
    }
    __syncthreads();
}

// Cross-block accumulators live ACC_STRIDE doubles (256 bytes) apart: same-line atomics serialise in one L2 slice
// (measured on B200: 148 blocks x 128 double atomics cost +6.3 us on contiguous slots, +1.2 us on 256-byte-strided ones).
constexpr int ACC_STRIDE = 32;
constexpr int ACC_SLOTS = 8;    // scalar reductions are spread over 8 slots per value (block index & 7)

// Publishes this block's partial sums and returns true (block-uniform) in the last block to arrive; in that block
// thread 0 holds the grand totals in v[] and the accumulators are re-armed (zeroed).
template <int NV>
__device__ __forceinline__ bool publish_and_elect(double (&v)[NV], double* acc, unsigned int* ticket) {
    __shared__ int s_last;
    block_sum<NV>(v);
    if (threadIdx.x == 0) {
        const int slot = blockIdx.x & (ACC_SLOTS - 1);
#pragma unroll
        for (int k = 0; k < NV; ++k) atomicAdd(acc + (k * ACC_SLOTS + slot) * ACC_STRIDE, v[k]);
        __threadfence();
        const unsigned int tk = atomicAdd(ticket, 1u);
        s_last = (tk == gridDim.x - 1);
        if (s_last) {
            __threadfence();
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                double t = 0.0;
                for (int q = 0; q < ACC_SLOTS; ++q) {
                    volatile double* a = acc + (k * ACC_SLOTS + q) * ACC_STRIDE;
                    t += *a;
                    *a = 0.0;
                }
                v[k] = t;
            }
            *ticket = 0u;
        }
    }
    __syncthreads();
    return s_last != 0;
}

// ---------------------------------------------------------------------------------------------------
// Activations (every DL4J layer carries one).  Derivatives are expressed through the layer output.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_fwd(int a, float x) {
    switch (a) {
        case N4J_ACT_RELU: return x > 0.f ? x : 0.f;
        case N4J_ACT_ELU: return x > 0.f ? x : __fsub_rn(expf(x), 1.f);
        case N4J_ACT_TANH: return tanhf(x);
        default: return x;
    }
}
__device__ __forceinline__ float act_bwd(int a, float y, float g) {
    switch (a) {
        case N4J_ACT_RELU: return y > 0.f ? g : 0.f;
        case N4J_ACT_ELU: return y > 0.f ? g : __fmul_rn(g, __fadd_rn(y, 1.f));
        case N4J_ACT_TANH: return __fmul_rn(g, __fsub_rn(1.f, __fmul_rn(y, y)));
        default: return g;
    }
}

// ---------------------------------------------------------------------------------------------------
// Host-side error plumbing
// ---------------------------------------------------------------------------------------------------
struct Error {
    int code;
    std::string msg;
};

#define N4J_CUDA(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            throw ::n4j::Error{N4J_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)};        \
    } while (0)

}  // namespace n4j
