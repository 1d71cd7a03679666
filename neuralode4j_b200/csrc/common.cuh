// common.cuh — shared device/host definitions of the node4j B200 engine.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/node4j.h"

namespace n4j {

struct CommDev;

// ---------------------------------------------------------------------------------------------------
// Programmatic dependent launch: every kernel of the engine is launched with programmaticStreamSerialization and starts
// with pdl_begin(): it lets ITS dependents be scheduled right away (their blocks become resident and park in
// griddepcontrol.wait) and then waits until everything it depends on has completed and flushed.  Dependencies stay exactly
// those of plain stream order; what is gained is the launch latency between the ~1000 small dependent kernels of a step.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_begin() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

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
    CommDev* comm;            // nullptr on a single GPU
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
    // multi-GPU: float4 index where the rank-replicated part of the state (theta, a_t) begins; only rank 0 counts it
    long long rep_begin4;
    int count_rep;
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
// Multi-GPU exchange over NVLink peer memory (one process per GPU, CUDA IPC).  The path's couplings are tiny — the
// error-norm scalar, 2*C BatchNorm sums, and the parameter-gradient block — so they are exchanged by a ONE-SHOT
// all-reduce written into the kernels themselves: every rank stores its partial values, each tagged with the exchange's
// sequence number, straight into every peer's mailbox, waits until its own mailbox shows the peers' tags and sums the
// contributions in rank order (bitwise identical totals on all ranks => identical controller decisions).  Two parities
// alternate; a rank can only be one exchange ahead of its slowest peer, so two buffers are enough.  Only the long
// parameter-gradient slots (k_comm_allreduce_vec) still use a flag per rank and pull the data.
// ---------------------------------------------------------------------------------------------------
constexpr int COMM_MAX_WORLD = 8;
constexpr int COMM_SLOT = 256;                                   // doubles per rank per small exchange
constexpr long long COMM_OFF_FLAGS = 0;                          // uint64 flag[2][8]
constexpr long long COMM_OFF_XFLAGS = 256;                       // uint64 xflag[2][8]
constexpr long long COMM_OFF_MBOX = 512;                         // uint64 cell[2][8][COMM_SLOT][2]  (tag<<32 | half of the double)
constexpr long long COMM_OFF_XBUF = COMM_OFF_MBOX + 2LL * COMM_MAX_WORLD * COMM_SLOT * 16;  // float xbuf[2][xcap]

struct CommDev {
    int rank, world;
    int status;                       // != 0: a peer did not show up in time
    int pad_;
    unsigned long long seq_small;     // small exchanges completed
    unsigned long long seq_x;         // vector exchanges completed
    unsigned int xticket;
    unsigned int pad2_;
    long long xcap;                   // floats per parity in xbuf
    long long xll_off;                // LL vector exchange: uint64 cell[2 parities][8 ranks][xll_cap] at this byte offset of every comm region
    long long xll_cap;                // cells (floats) per (parity, rank); 0: no LL buffer
    long long stat_off;               // BatchNorm statistics mailboxes (sharded deferred scheme, layer_kernels.cuh: BnScratch) at this byte offset of every comm region
    char* peer[COMM_MAX_WORLD];       // base of every rank's comm region (peer[rank] is the local one)
};

__device__ __forceinline__ bool comm_wait_flag(volatile unsigned long long* f, unsigned long long seq, CommDev* cm) {
    for (unsigned int spin = 0; spin < (1u << 28); ++spin) {
        if (*f >= seq) return true;
        if ((spin & 1023) == 1023) __nanosleep(64);
    }
    cm->status = 1;
    return false;
}

// Small all-reduce, LL style: every 8-byte cell carries 4 bytes of payload and the 4-byte sequence tag, so a value is
// complete exactly when its two cells (one 16-byte store) show the current tag — no flags, no system fences on the critical path
// (one NVLink store latency per exchange).  Called by `nthr` threads (tid 0..nthr-1) that can synchronise through `sync()`;
// vals[n] (shared memory, n <= COMM_SLOT) in -> global totals out (summed in rank order, identical on every rank).
template <class SyncFn>
__device__ __forceinline__ void comm_allreduce_small(CommDev* cm, double* vals, int n, int tid, int nthr, SyncFn sync) {
    // Latency matters here, not bandwidth: every thread reads the sequence number, rank, world and its peer bases itself (independent
    // loads of one cache line, no broadcast through shared memory in front of the stores), the own contribution never leaves shared
    // memory, and the peers' cells of a value are requested together and re-requested until all carry the tag.
    const unsigned long long seq = *((volatile unsigned long long*)&cm->seq_small) + 1;
    const int me = cm->rank, W = cm->world;
    const unsigned int tag = (unsigned int)seq;
    const int par = (int)(seq & 1);
    // push: cell[(par, me, 2*i + half)] on every peer
    for (int idx = tid; idx < (W - 1) * n; idx += nthr) {
        const int p = idx / n, i = idx - p * n, pr = p + (p >= me ? 1 : 0);
        const unsigned long long bits = (unsigned long long)__double_as_longlong(vals[i]);
        unsigned long long* cell = reinterpret_cast<unsigned long long*>(cm->peer[pr] + COMM_OFF_MBOX) + ((long long)(par * COMM_MAX_WORLD + me) * COMM_SLOT + i) * 2;
        const ulonglong2 c = make_ulonglong2(((unsigned long long)tag << 32) | (bits & 0xFFFFFFFFull), ((unsigned long long)tag << 32) | (bits >> 32));
        asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(cell), "l"(c.x), "l"(c.y) : "memory");
    }
    sync();      // every thread has read seq_small
    if (tid == 0) cm->seq_small = seq;
    // pull: sum in rank order
    const unsigned long long* mbox = reinterpret_cast<const unsigned long long*>(cm->peer[me] + COMM_OFF_MBOX) + (long long)par * COMM_MAX_WORLD * COMM_SLOT * 2;
    for (int i = tid; i < n; i += nthr) {
        unsigned long long c0[COMM_MAX_WORLD], c1[COMM_MAX_WORLD];
        bool ok = false;
        for (unsigned int spin = 0; spin < (1u << 26) && !ok; ++spin) {
            ok = true;
#pragma unroll
            for (int q = 0; q < COMM_MAX_WORLD; ++q) {
                if (q < W && q != me) {
                    const unsigned long long* cell = mbox + ((long long)q * COMM_SLOT + i) * 2;
                    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(c0[q]), "=l"(c1[q]) : "l"(cell) : "memory");
                }
            }
#pragma unroll
            for (int q = 0; q < COMM_MAX_WORLD; ++q)
                if (q < W && q != me && ((unsigned int)(c0[q] >> 32) != tag || (unsigned int)(c1[q] >> 32) != tag)) ok = false;
            if (!ok && (spin & 255) == 255) __nanosleep(32);
        }
        if (!ok) cm->status = 1;
        const double mine = vals[i];
        double t = 0.0;
#pragma unroll
        for (int q = 0; q < COMM_MAX_WORLD; ++q)
            if (q < W) t += q == me ? mine : __longlong_as_double((long long)((c0[q] & 0xFFFFFFFFull) | (c1[q] << 32)));
        vals[i] = t;
    }
    sync();
}
struct BlockSyncFn { __device__ __forceinline__ void operator()() const { __syncthreads(); } };
__device__ __forceinline__ void comm_allreduce_block(CommDev* cm, double* vals, int n) {
    comm_allreduce_small(cm, vals, n, (int)threadIdx.x, (int)blockDim.x, BlockSyncFn{});
}

// LL all-reduce of four consecutive floats of a vector (parameter-gradient slots): every float travels as an 8-byte cell
// {payload, tag}, two cells per 16-byte store, straight into every peer's cell[parity][my rank][i..i+3]; the caller's own cells
// are then polled for the peers' tags and summed in rank order (bitwise identical on all ranks).  No tickets, flags or fences on
// the way: one NVLink store latency.  i is a float index (multiple of 4) below xll_cap; tag = the exchange's sequence number.
__device__ __forceinline__ float4 comm_ll_allreduce4(const CommDev* cm, unsigned int tag, int par, long long i, float4 v) {
    const int me = cm->rank, W = cm->world;
    const unsigned long long hi = (unsigned long long)tag << 32;
    const ulonglong2 c01 = make_ulonglong2(hi | __float_as_uint(v.x), hi | __float_as_uint(v.y));
    const ulonglong2 c23 = make_ulonglong2(hi | __float_as_uint(v.z), hi | __float_as_uint(v.w));
    const long long slot = ((long long)(par * COMM_MAX_WORLD + me) * cm->xll_cap + i) * 8;
#pragma unroll
    for (int q = 0; q < COMM_MAX_WORLD; ++q) {
        if (q < W && q != me) {
            ulonglong2* dst = reinterpret_cast<ulonglong2*>(cm->peer[q] + cm->xll_off + slot);
            dst[0] = c01;
            dst[1] = c23;
        }
    }
    // the cells of ALL peers are requested before the first tag is looked at: one L2 round trip per attempt whatever the number of ranks
    // (peer after peer it was seven dependent round trips at eight ranks, by 9216 threads spinning beside the main chain's convolutions)
    const unsigned long long* base = reinterpret_cast<const unsigned long long*>(cm->peer[me] + cm->xll_off);
    unsigned long long a[COMM_MAX_WORLD][4];
    bool ok = false;
    for (unsigned int spin = 0; spin < (1u << 26) && !ok; ++spin) {
        ok = true;
#pragma unroll
        for (int q = 0; q < COMM_MAX_WORLD; ++q) {
            if (q < W && q != me) {
                const unsigned long long* src = base + ((long long)(par * COMM_MAX_WORLD + q) * cm->xll_cap + i);
                asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a[q][0]), "=l"(a[q][1]) : "l"(src) : "memory");
                asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a[q][2]), "=l"(a[q][3]) : "l"(src + 2) : "memory");
            }
        }
#pragma unroll
        for (int q = 0; q < COMM_MAX_WORLD; ++q)
            if (q < W && q != me &&
                ((unsigned int)(a[q][0] >> 32) != tag || (unsigned int)(a[q][1] >> 32) != tag || (unsigned int)(a[q][2] >> 32) != tag || (unsigned int)(a[q][3] >> 32) != tag))
                ok = false;
        if (!ok && (spin & 15) == 15) __nanosleep(64);
    }
    if (!ok) const_cast<CommDev*>(cm)->status = 1;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int q = 0; q < COMM_MAX_WORLD; ++q) {
        if (q >= W) continue;
        float4 u = v;
        if (q != me)
            u = make_float4(__uint_as_float((unsigned int)a[q][0]), __uint_as_float((unsigned int)a[q][1]), __uint_as_float((unsigned int)a[q][2]),
                            __uint_as_float((unsigned int)a[q][3]));
        t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
    }
    return t;
}
// every block of an LL vector exchange calls this when it leaves: the last one bumps the sequence number for the next exchange
__device__ __forceinline__ void comm_ll_leave(CommDev* cm, unsigned long long seq) {
    if (threadIdx.x == 0) {
        const unsigned int tk = atomicAdd(&cm->pad2_, 1u);
        if (tk == gridDim.x - 1) { cm->pad2_ = 0u; __threadfence(); cm->seq_x = seq; }
    }
}

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
__device__ __forceinline__ bool publish_and_elect(double (&v)[NV], double* acc, unsigned int* ticket, CommDev* cm = nullptr) {
    __shared__ int s_last;
    __shared__ double s_tot[NV];
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
                s_tot[k] = t;
            }
            *ticket = 0u;
        }
    }
    __syncthreads();
    if (s_last && cm) {   // multi-GPU: the last block of every rank exchanges its totals
        comm_allreduce_block(cm, s_tot, NV);
        if (threadIdx.x == 0) {
#pragma unroll
            for (int k = 0; k < NV; ++k) v[k] = s_tot[k];
        }
    }
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
