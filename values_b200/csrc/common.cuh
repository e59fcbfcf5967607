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

// Ring hand-back.  A consumer warp hands a stage back to the copy engine with ONE arrive, by lane 0,
// after a warp barrier that orders every lane's shared-memory reads before it (__syncwarp guarantees
// memory ordering among the participating threads).  compute-sanitizer's racecheck does not connect
// the other 31 lanes' reads to lane 0's arrive and reports the next copy into the stage as a hazard;
// building with -DVALUES_ALL_LANES_ARRIVE makes every lane arrive itself (barrier count x 32; slower),
// which is the build the sanitizer logs under profiles/ were taken with.
#ifdef VALUES_ALL_LANES_ARRIVE
constexpr int kArriveLanes = 32;
#else
constexpr int kArriveLanes = 1;
#endif
__device__ __forceinline__ bool arrives(int lane) { return kArriveLanes == 32 || lane == 0; }

// ------------------------------------------------------------------ async copies (TMA engine) + mbarriers
// Producer / consumer rings: the producer arms a stage's "full" barrier with the byte count and
// issues cp.async.bulk[.tensor] copies that complete on it; consumer warps wait on "full", read
// the stage and hand it back with one arrive per warp on "empty" (arrive is a CTA-scope release:
// the warp's reads of the stage are ordered before the producer's next copy into it).
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// A CTA invalidates its mbarriers before it exits: the next CTA scheduled on the SM finds the same
// shared-memory words, and PTX leaves re-using a word that still holds a live mbarrier object for anything
// but that object undefined (mbarrier.init / mbarrier.inval).  Hygiene: no measured effect.
__device__ __forceinline__ void mbar_inval(uint64_t* bar) {
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                     " selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(a), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void mbar_wait_addr(uint32_t a, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                     " selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(a), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void mbar_arrive_addr(uint32_t a) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 q;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(a) : "memory");
    return q;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}

// 4-D tiled tensor-map copy (cp.async.bulk.tensor -> SASS UTMALDG): box at element coordinates
// (c0 fastest), out-of-range elements arrive as zeros, the full box counts towards complete_tx.
__device__ __forceinline__ void tma_load_4d_addr(uint32_t dst, const void* tmap, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_4d_addr_hint(uint32_t dst, const void* tmap, int c0, int c1, int c2, int c3, uint64_t* bar, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4, %5}], [%6], %7;"
                 ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_load_5d_addr(uint32_t dst, const void* tmap, int c0, int c1, int c2, int c3, int c4, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// partials [n_items, n_blocks, K] -> out [n_items, K]; fixed summation order.
int launch_reduce_partials(const double* partials, int64_t n_items, int64_t n_blocks, int K,
                           double* out, cudaStream_t stream);

}  // namespace vb
