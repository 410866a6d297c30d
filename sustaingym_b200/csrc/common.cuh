// Shared device/host helpers for the sustaingym_b200 kernels (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/sustaingym_b200.h"

namespace sg {

// ---- error plumbing (thread-local message, no global mutable state) -------------------------
void set_error(const char* fmt, ...);

#define SG_CHECK_ARG(cond, ...)                 \
    do {                                        \
        if (!(cond)) {                          \
            ::sg::set_error(__VA_ARGS__);       \
            return 1;                           \
        }                                       \
    } while (0)

#define SG_CUDA(call)                                                                   \
    do {                                                                                \
        cudaError_t err__ = (call);                                                     \
        if (err__ != cudaSuccess) {                                                     \
            ::sg::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__),  \
                            __FILE__, __LINE__);                                        \
            return 2;                                                                   \
        }                                                                               \
    } while (0)

inline int launch_status(const char* what) {
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) {
        set_error("launch of %s failed: %s", what, cudaGetErrorString(err));
        return 3;
    }
    return 0;
}

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// ---- warp reductions ------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(kFull, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(kFull, v, o));
    return v;
}
__device__ __forceinline__ int warp_sum_int(int v) { return __reduce_add_sync(kFull, v); }

// ---- streaming (read-once / write-once) global accesses: keep them out of L1 ----------------
__device__ __forceinline__ float2 ldg_stream_f2(const float2* p) {
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ldg_stream_f4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream_f2(float2* p, float2 v) {
    asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void stg_stream_f4(float4* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w) : "memory");
}

// ---- aggregate statistics (SG_STATS_LEN doubles; see include/sustaingym_b200.h) --------------
// One call per warp with warp-reduced values; lane 0 issues the atomics.  min/max use the
// order-preserving int64 trick on the double bit pattern.
__device__ __forceinline__ void atomic_min_double(double* addr, double v) {
    unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
    unsigned long long old = *a, assumed;
    do {
        assumed = old;
        if (__longlong_as_double((long long)assumed) <= v) break;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    } while (assumed != old);
}
__device__ __forceinline__ void atomic_max_double(double* addr, double v) {
    unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
    unsigned long long old = *a, assumed;
    do {
        assumed = old;
        if (__longlong_as_double((long long)assumed) >= v) break;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    } while (assumed != old);
}

struct StatsAcc {
    double count = 0, sum = 0, sumsq = 0, mn = 1e300, mx = -1e300, n_done = 0, n_proj = 0, n_maxiter = 0;
    double c0 = 0, c1 = 0, c2 = 0;
    __device__ __forceinline__ void add(double r, bool done, double k0 = 0, double k1 = 0, double k2 = 0) {
        count += 1; sum += r; sumsq += r * r; mn = fmin(mn, r); mx = fmax(mx, r);
        n_done += done ? 1.0 : 0.0; c0 += k0; c1 += k1; c2 += k2;
    }
};

// Block-level flush: every thread holds a StatsAcc; reduce across the block and issue ~11
// atomics per block.  `scratch` must hold 11 * (blockDim.x / 32) doubles of shared memory.
__device__ __forceinline__ void stats_flush_block(const StatsAcc& a, double* stats, double* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    double v[11] = {a.count, a.sum, a.sumsq, a.mn, a.mx, a.n_done, a.n_proj, a.n_maxiter, a.c0, a.c1, a.c2};
#pragma unroll
    for (int i = 0; i < 11; ++i) {
        double r = (i == 3) ? warp_min(v[i]) : (i == 4) ? warp_max(v[i]) : warp_sum(v[i]);
        if (lane == 0) scratch[i * nwarp + warp] = r;
    }
    __syncthreads();
    if (threadIdx.x < 11) {
        const int i = threadIdx.x;
        double r = scratch[i * nwarp];
        for (int w = 1; w < nwarp; ++w) {
            double o = scratch[i * nwarp + w];
            r = (i == 3) ? fmin(r, o) : (i == 4) ? fmax(r, o) : r + o;
        }
        if (i == 3) atomic_min_double(stats + 3, r);
        else if (i == 4) atomic_max_double(stats + 4, r);
        else if (r != 0.0) atomicAdd(stats + i, r);
    }
}

}  // namespace sg
