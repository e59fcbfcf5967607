// Library-wide plumbing: error reporting, launch counter, the deterministic second-stage
// reduction shared by K1 and K2a.
#include <stdarg.h>

#include "common.cuh"

namespace vb {

thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_launch(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(VALUES_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return VALUES_OK;
}

// One block per item; thread t sums rows t, t+256, ... then a fixed-order block reduction.
template <int K>
__global__ void __launch_bounds__(kThreads) reduce_partials_kernel(const double* __restrict__ partials,
                                                                   int64_t n_blocks,
                                                                   double* __restrict__ out) {
    __shared__ double red[K * 8];
    const double* src = partials + (int64_t)blockIdx.x * n_blocks * K;
    double acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.0;
    for (int64_t r = threadIdx.x; r < n_blocks; r += kThreads) {
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] += src[r * K + k];
    }
    block_sum<K>(acc, red);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) out[(int64_t)blockIdx.x * K + k] = acc[k];
    }
}

int launch_reduce_partials(const double* partials, int64_t n_items, int64_t n_blocks, int K,
                           double* out, cudaStream_t stream) {
    if (n_items <= 0) return VALUES_OK;
    if (K == 9)
        reduce_partials_kernel<9><<<(unsigned)n_items, kThreads, 0, stream>>>(partials, n_blocks, out);
    else if (K == 5)
        reduce_partials_kernel<5><<<(unsigned)n_items, kThreads, 0, stream>>>(partials, n_blocks, out);
    else if (K == 3)
        reduce_partials_kernel<3><<<(unsigned)n_items, kThreads, 0, stream>>>(partials, n_blocks, out);
    else if (K == 7)     // seg_loss_terms, CT = 2 / 4 / 8
        reduce_partials_kernel<7><<<(unsigned)n_items, kThreads, 0, stream>>>(partials, n_blocks, out);
    else if (K == 13)
        reduce_partials_kernel<13><<<(unsigned)n_items, kThreads, 0, stream>>>(partials, n_blocks, out);
    else if (K == 25)
        reduce_partials_kernel<25><<<(unsigned)n_items, kThreads, 0, stream>>>(partials, n_blocks, out);
    else
        return set_error(VALUES_ERR_INVALID_ARG, "reduce_partials: K=%d", K);
    return check_launch("reduce_partials_kernel");
}

}  // namespace vb

extern "C" int values_abi_version(void) { return VALUES_ABI_VERSION; }
extern "C" const char* values_last_error(void) { return vb::g_err; }
extern "C" int64_t values_launch_count(void) { return vb::g_launches.load(); }
