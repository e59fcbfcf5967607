// Shared device/host helpers for the values_b200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "values_b200.h"

namespace vb {

constexpr int kThreads = 256;

extern thread_local char g_err[512];
extern std::atomic<int64_t> g_launches;

int set_error(int code, const char* fmt, ...);
int check_launch(const char* what);  // cudaGetLastError after a launch; counts it

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ------------------------------------------------------------------ streaming loads
// 128-bit read-once loads: bypass L1 allocation (every probability is read exactly once).
__device__ __forceinline__ uint4 ldg_stream_128(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

template <typename T> struct In;  // input element traits
template <> struct In<float> {
    using acc_t = float;                 // arithmetic type of p * log(p) (reference: input dtype)
    static constexpr int VEC = 4;        // elements per 128-bit load
    __device__ static __forceinline__ void load_vec(const float* p, float (&o)[4]) {
        uint4 r = ldg_stream_128(p);
        o[0] = __uint_as_float(r.x); o[1] = __uint_as_float(r.y);
        o[2] = __uint_as_float(r.z); o[3] = __uint_as_float(r.w);
    }
    __device__ static __forceinline__ float load_one(const float* p) { return __ldg(p); }
};
template <> struct In<double> {
    using acc_t = double;
    static constexpr int VEC = 2;
    __device__ static __forceinline__ void load_vec(const double* p, double (&o)[2]) {
        uint4 r = ldg_stream_128(p);
        o[0] = __hiloint2double((int)r.y, (int)r.x);
        o[1] = __hiloint2double((int)r.w, (int)r.z);
    }
    __device__ static __forceinline__ double load_one(const double* p) { return __ldg(p); }
};
template <> struct In<__nv_bfloat16> {
    using acc_t = float;  // bf16 is widened exactly to fp32, then the fp32 algorithm runs
    static constexpr int VEC = 8;
    __device__ static __forceinline__ void load_vec(const __nv_bfloat16* p, float (&o)[8]) {
        uint4 r = ldg_stream_128(p);
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            o[2 * i] = __uint_as_float(w[i] << 16);
            o[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    __device__ static __forceinline__ float load_one(const __nv_bfloat16* p) {
        return __uint_as_float(((uint32_t) * reinterpret_cast<const uint16_t*>(p)) << 16);
    }
};

// ------------------------------------------------------------------ reductions
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// Block barrier: BAR == 0 is __syncthreads(); BAR > 0 is the named barrier BAR over the first
// kThreads threads only (kernels with an extra producer warp that does not take part).
template <int BAR> __device__ __forceinline__ void block_sync() {
    if constexpr (BAR == 0) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"n"(BAR), "n"(kThreads) : "memory");
}
// Deterministic block sum of K doubles per thread -> valid in thread 0.
template <int K, int BAR = 0>
__device__ __forceinline__ void block_sum(double (&v)[K], double* smem /* [K * 8] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
    block_sync<BAR>();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) smem[k * 8 + warp] = v[k];
    }
    block_sync<BAR>();
    if (threadIdx.x == 0) {
        const int nw = BAR == 0 ? (int)(blockDim.x >> 5) : kThreads / 32;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double s = 0.0;
            for (int w = 0; w < nw; ++w) s += smem[k * 8 + w];
            v[k] = s;
        }
    }
}

// partials [n_items, n_blocks, K] -> out [n_items, K]; fixed summation order.
int launch_reduce_partials(const double* partials, int64_t n_items, int64_t n_blocks, int K,
                           double* out, cudaStream_t stream);

}  // namespace vb
