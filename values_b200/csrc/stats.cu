// K4: whole-map statistics that sit either side of the C2 -> C3 path.
//   count_nonzero   : calculate_foreground_quantile_image        find_threshold.py:11-13
//   radix_histogram : exact np.quantile over all validation voxels (radix select; the
//   min_key_above     histograms of several maps / ranks add up)  find_threshold.py:63-68
//   pair_moments    : compute_ncc (means, ddof=1 std, cross term) evaluation/metrics/ncc.py:9-25
//   calib_bins      : calib_stats 20-bin histogram               evaluation/metrics/ace.py:49-81
//   calib_bins_fused: the per-image body of calibration_error    evaluation/metrics/ace.py:89-135
// All are single HBM sweeps (sizeof(T) bytes per element); integer results are exact and
// order-free (integer atomics), floating sums use fixed-order block-then-grid reductions.
#include "common.cuh"

namespace vb {

constexpr int kStatEPT = 16;  // elements per thread per block (pair_moments / calib kernels)

// ------------------------------------------------------------------ generic element loads
template <typename T> __device__ __forceinline__ T ld_elem(const T* p) { return __ldg(p); }

template <typename T> __device__ __forceinline__ bool is_nonzero(T v) { return v != (T)0; }

// =============================================================================== count_nonzero
// 16-byte vector body + scalar edges; integer count -> one atomicAdd per block (exact).
template <typename T>
__global__ void __launch_bounds__(kThreads) count_nonzero_kernel(const T* __restrict__ x, int64_t n,
                                                                 unsigned long long* __restrict__ out) {
    constexpr int VEC = 16 / sizeof(T);
    __shared__ unsigned long long red[kThreads / 32];
    const uintptr_t addr = reinterpret_cast<uintptr_t>(x);
    int64_t head = (int64_t)(((16 - (addr & 15)) & 15) / sizeof(T));
    if (head > n) head = n;
    const int64_t nvec = (n - head) / VEC;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
    unsigned long long c = 0;
    const uint4* xv = reinterpret_cast<const uint4*>(x + head);
    for (int64_t i = tid; i < nvec; i += nthr) {
        const uint4 r = ldg_stream_128(xv + i);
        T e[VEC];
        memcpy(e, &r, 16);
#pragma unroll
        for (int k = 0; k < VEC; ++k) c += is_nonzero(e[k]) ? 1u : 0u;
    }
    const int64_t tail0 = head + nvec * VEC;
    for (int64_t i = tid; i < head; i += nthr) c += is_nonzero(ld_elem(x + i)) ? 1u : 0u;
    for (int64_t i = tail0 + tid; i < n; i += nthr) c += is_nonzero(ld_elem(x + i)) ? 1u : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long s = 0;
        for (int w = 0; w < kThreads / 32; ++w) s += red[w];
        if (s) atomicAdd(out, s);
    }
}

// =============================================================================== radix select
// Order-preserving keys: NaN -> all ones (sorts last, as np.sort), -0.0 -> key of +0.0.
template <typename T> struct Key;
template <> struct Key<float> {
    using type = uint32_t;
    static constexpr int BITS = 32;
    __device__ static __forceinline__ uint32_t of(float v) {
        if (v != v) return 0xffffffffu;
        uint32_t u = __float_as_uint(v);
        if (u == 0x80000000u) u = 0;
        return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    }
};
template <> struct Key<double> {
    using type = uint64_t;
    static constexpr int BITS = 64;
    __device__ static __forceinline__ uint64_t of(double v) {
        if (v != v) return ~0ull;
        uint64_t u = (uint64_t)__double_as_longlong(v);
        if (u == 0x8000000000000000ull) u = 0;
        return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
    }
};

constexpr int kMaxDigitBits = 11;

// hist[d] += #{ i : key(x_i) >> (BITS - prefix_bits) == prefix  and  (key >> shift) & mask == d }
// Block-private shared histogram in R = 8192 / bins replicas (32 KB: 4 replicas of 2048 bins, 8 of 1024,
// 32 of 256), a lane adds into replica lane % R: the leading digits (sign + exponent) repeat heavily, and the
// replicas keep lanes that hold the same digit on different words.  One shared atomic per matching element
// and no cross-lane dependency; two 16-byte vectors in flight per thread.  (r01 / early r02 aggregated equal
// digits per warp first -- match.any, then two ballot rounds per element: exact, but every element paid a
// chain of votes and shuffles, 0.12-0.18 of the HBM peak on the top digit of fp32 maps.)
constexpr int kHistWords = 8192;
// The maps of one launch: a validation set is many separate images, and a launch per image and digit made the
// exact quantile launch-bound (131 launches for 32 maps: 1.1 ms for 0.27 ms of memory traffic).  The table
// travels as a kernel parameter; blockIdx.y picks the map.
constexpr int kSetMax = 96;
struct MapSet { const void* p[kSetMax]; int64_t n[kSetMax]; };
template <typename T>
__global__ void __launch_bounds__(kThreads) radix_hist_kernel(const __grid_constant__ MapSet set,
                                                              uint64_t prefix, int prefix_bits,
                                                              int shift, int digit_bits,
                                                              unsigned long long* __restrict__ hist,
                                                              const unsigned long long* __restrict__ state) {
    using K = typename Key<T>::type;
    constexpr int VEC = 16 / sizeof(T);
    __shared__ unsigned int sh[kHistWords];
    const T* __restrict__ x = reinterpret_cast<const T*>(set.p[blockIdx.y]);
    const int64_t n = set.n[blockIdx.y];
    if (state) {   // prefix chosen on the device by radix_select_kernel (no host round trip per digit)
        if (state[1] + (unsigned long long)digit_bits > (unsigned long long)Key<T>::BITS) return;   // select flagged an error
        prefix = state[0];
        prefix_bits = (int)state[1];
        shift = Key<T>::BITS - prefix_bits - digit_bits;
    }
    const int nbins = 1 << digit_bits;
    const int rbits = 13 - digit_bits > 5 ? 5 : 13 - digit_bits;      // log2 of the replica count (<= 32)
    const int words = nbins << rbits;
    for (int i = threadIdx.x; i < words; i += kThreads) sh[i] = 0;
    __syncthreads();
    const K mask = (K)(nbins - 1);
    const int pshift = Key<T>::BITS - prefix_bits;
    const unsigned rep = threadIdx.x & ((1u << rbits) - 1u);
    auto add = [&](T v) {
        const K key = Key<T>::of(v);
        if (prefix_bits == 0 || (uint64_t)(key >> pshift) == prefix)
            atomicAdd(&sh[((unsigned)((key >> shift) & mask) << rbits) + rep], 1u);
    };
    auto add_vec = [&](const uint4& r) {
        T e[VEC];
        memcpy(e, &r, 16);
#pragma unroll
        for (int k = 0; k < VEC; ++k) add(e[k]);
    };
    const uintptr_t addr = reinterpret_cast<uintptr_t>(x);
    int64_t head = (int64_t)(((16 - (addr & 15)) & 15) / sizeof(T));
    if (head > n) head = n;
    const int64_t nvec = (n - head) / VEC;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
    const uint4* xv = reinterpret_cast<const uint4*>(x + head);
    int64_t i = tid;
    for (; i + 3 * nthr < nvec; i += 4 * nthr) {
        const uint4 r0 = ldg_stream_128(xv + i), r1 = ldg_stream_128(xv + i + nthr);
        const uint4 r2 = ldg_stream_128(xv + i + 2 * nthr), r3 = ldg_stream_128(xv + i + 3 * nthr);
        add_vec(r0);
        add_vec(r1);
        add_vec(r2);
        add_vec(r3);
    }
    for (; i < nvec; i += nthr) add_vec(ldg_stream_128(xv + i));
    const int64_t tail0 = head + nvec * VEC;
    if (blockIdx.x == 0 && threadIdx.x < 32) {   // < VEC head and < VEC tail elements
        const int lane = threadIdx.x;
        if (lane < head) add(ld_elem(x + lane));
        if (tail0 + lane < n) add(ld_elem(x + tail0 + lane));
    }
    __syncthreads();
    for (int b = threadIdx.x; b < nbins; b += kThreads) {
        unsigned int c = 0;
        for (int r = 0; r < (1 << rbits); ++r) c += sh[(b << rbits) + r];
        if (c) atomicAdd(hist + b, (unsigned long long)c);
    }
}

// One thread block: the bucket of `hist` that holds the wanted rank becomes the next digit of the prefix.
// state = {prefix, prefix_bits, rank within the elements matching the prefix, count in the chosen bucket,
// count in the LAST bucket of the first digit (fp32: NaN count)}; hist is zeroed for the next digit.
__global__ void __launch_bounds__(kThreads) radix_select_kernel(unsigned long long* __restrict__ hist, int digit_bits,
                                                                unsigned long long* __restrict__ state) {
    __shared__ unsigned long long part[kThreads];
    __shared__ unsigned long long cum[kThreads];
    const int nbins = 1 << digit_bits, per = (nbins + kThreads - 1) / kThreads;
    const int tid = threadIdx.x;
    unsigned long long loc = 0;
    for (int i = tid * per; i < min(nbins, (tid + 1) * per); ++i) loc += hist[i];
    part[tid] = loc;
    __syncthreads();
    if (tid == 0) {
        unsigned long long c = 0;
        for (int t = 0; t < kThreads; ++t) { cum[t] = c; c += part[t]; }   // exclusive prefix over threads
        if (state[1] == 0) state[4] = hist[nbins - 1];
        if (state[2] >= c) state[1] = ~0ull;                               // rank out of range
    }
    __syncthreads();
    const unsigned long long rank = state[2];
    if (state[1] != ~0ull && rank >= cum[tid] && rank < cum[tid] + part[tid]) {   // exactly one thread
        unsigned long long below = cum[tid];
        for (int i = tid * per; i < min(nbins, (tid + 1) * per); ++i) {
            const unsigned long long h = hist[i];
            if (rank < below + h) {
                state[0] = (state[0] << digit_bits) | (unsigned long long)i;
                state[1] += (unsigned long long)digit_bits;
                state[2] = rank - below;
                state[3] = h;
                break;
            }
            below += h;
        }
    }
    __syncthreads();
    for (int i = tid; i < nbins; i += kThreads) hist[i] = 0;
}

// out[0] = min(out[0], min key(x_i) > key)   (atomicMin: order-free)
template <typename T>
__global__ void __launch_bounds__(kThreads) min_key_above_kernel(const __grid_constant__ MapSet set,
                                                                 uint64_t key,
                                                                 unsigned long long* __restrict__ out) {
    __shared__ unsigned long long red[kThreads / 32];
    const T* __restrict__ x = reinterpret_cast<const T*>(set.p[blockIdx.y]);
    const int64_t n = set.n[blockIdx.y];
    unsigned long long best = ~0ull;
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nthr) {
        const unsigned long long k = (unsigned long long)Key<T>::of(ld_elem(x + i));
        if (k > key && k < best) best = k;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other < best ? other : best;
    }
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kThreads / 32; ++w) best = red[w] < best ? red[w] : best;
        if (best != ~0ull) atomicMin(out, best);
    }
}

// =============================================================================== pair_moments
// partials[m, blk, 5] = { sum(a-sa), sum(b-sb), sum (a-sa)^2, sum (b-sb)^2, sum (a-sa)(b-sb) }
template <typename TA, typename TB>
__global__ void __launch_bounds__(kThreads) pair_moments_kernel(const TA* __restrict__ a, int64_t stride_a,
                                                                const TB* __restrict__ b, int64_t stride_b,
                                                                int64_t V, int64_t bpm,
                                                                const double* __restrict__ shift,
                                                                double* __restrict__ partials) {
    __shared__ double red[5 * 8];
    const int64_t m = blockIdx.x / bpm;
    const int64_t blk = blockIdx.x - m * bpm;
    const TA* pa = a + m * stride_a;
    const TB* pb = b + m * stride_b;
    const double sa = shift ? shift[2 * m] : 0.0, sb = shift ? shift[2 * m + 1] : 0.0;
    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    const int64_t base = blk * (int64_t)(kThreads * kStatEPT) + threadIdx.x;
#pragma unroll 4
    for (int i = 0; i < kStatEPT; ++i) {
        const int64_t v = base + (int64_t)i * kThreads;
        if (v < V) {
            const double x = (double)ld_elem(pa + v) - sa;
            const double y = (double)ld_elem(pb + v) - sb;
            acc[0] += x; acc[1] += y;
            acc[2] = fma(x, x, acc[2]); acc[3] = fma(y, y, acc[3]); acc[4] = fma(x, y, acc[4]);
        }
    }
    block_sum<5>(acc, red);
    if (threadIdx.x == 0) {
        double* dst = partials + (int64_t)blockIdx.x * 5;
#pragma unroll
        for (int k = 0; k < 5; ++k) dst[k] = acc[k];
    }
}

// =============================================================================== calib_bins
constexpr int kCalibBins = 20;            // n_bins of calib_stats (ace.py:51)
constexpr int kCalibSlots = kCalibBins + 1;  // np.bincount(minlength=len(bins)) -> 21 slots
struct CalibEdges { double e[kCalibSlots]; };

// np.digitize(p, bins) - 1 for increasing bins: largest k with bins[k] <= p (-1 below bins[0]);
// NaN compares false everywhere -> digitize returns len(bins) -> slot 20 here (clamped).
__device__ __forceinline__ int calib_slot(double p, const CalibEdges& ed) {
    if (!(p >= ed.e[0])) return (p != p) ? kCalibSlots - 1 : -1;
    int k = (int)(p * ((double)kCalibBins / ed.e[kCalibBins]));
    k = k < 0 ? 0 : (k > kCalibSlots - 1 ? kCalibSlots - 1 : k);
    while (k > 0 && p < ed.e[k]) --k;
    while (k < kCalibSlots - 1 && p >= ed.e[k + 1]) ++k;
    return k;
}

struct CalibAcc {
    double psum[kCalibSlots];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int k = 0; k < kCalibSlots; ++k) psum[k] = 0.0;
    }
    // count / true-count: warp-aggregated shared integer atomics (exact, order-free);
    // sum of confidences: predicated register accumulators (fixed order -> deterministic)
    __device__ __forceinline__ void add(bool live, int slot, double p, bool is_true,
                                        unsigned int* s_cnt, unsigned int* s_true) {
        live = live && slot >= 0;
        const unsigned active = __ballot_sync(0xffffffffu, live);
        if (live) {
            const unsigned code = (unsigned)slot * 2u + (is_true ? 1u : 0u);
            const unsigned peers = __match_any_sync(active, code);
            if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) {
                atomicAdd(&s_cnt[slot], (unsigned)__popc(peers));
                if (is_true) atomicAdd(&s_true[slot], (unsigned)__popc(peers));
            }
#pragma unroll
            for (int k = 0; k < kCalibSlots; ++k) psum[k] += (k == slot) ? p : 0.0;
        }
    }
    // the same for a voxel that stands for n_live (rater, voxel) pairs, n_true of them correct:
    // confidence and slot depend on the voxel only, so the 21-way predicated add runs once per voxel
    __device__ __forceinline__ void add_n(bool live, int slot, double p, unsigned n_live, unsigned n_true,
                                          unsigned int* s_cnt, unsigned int* s_true) {
        live = live && slot >= 0 && n_live > 0;
        const unsigned active = __ballot_sync(0xffffffffu, live);
        if (live) {
            const unsigned peers = __match_any_sync(active, (unsigned)slot);
            const unsigned tot = __reduce_add_sync(peers, n_live), tru = __reduce_add_sync(peers, n_true);
            if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) {
                atomicAdd(&s_cnt[slot], tot);
                if (tru) atomicAdd(&s_true[slot], tru);
            }
            const double pn = p * (double)n_live;
#pragma unroll
            for (int k = 0; k < kCalibSlots; ++k) psum[k] += (k == slot) ? pn : 0.0;
        }
    }
};

// partials [blocks, 3 * 21] = { count, sum p, sum true } per slot
template <int BAR = 0>
__device__ __forceinline__ void calib_flush(CalibAcc& acc, unsigned int* s_cnt, unsigned int* s_true,
                                            double* red, double* __restrict__ partials) {
    block_sum<kCalibSlots>(acc.psum, red);
    __syncthreads();
    if (threadIdx.x < kCalibSlots) {
        double* dst = partials + (int64_t)blockIdx.x * 3 * kCalibSlots;
        dst[threadIdx.x] = (double)s_cnt[threadIdx.x];
        dst[2 * kCalibSlots + threadIdx.x] = (double)s_true[threadIdx.x];
    }
    if (threadIdx.x == 0) {
        double* dst = partials + (int64_t)blockIdx.x * 3 * kCalibSlots + kCalibSlots;
#pragma unroll
        for (int k = 0; k < kCalibSlots; ++k) dst[k] = acc.psum[k];
    }
}

template <typename TP, typename TL>
__global__ void __launch_bounds__(kThreads) calib_bins_kernel(const TP* __restrict__ prob,
                                                              const TL* __restrict__ correct, int64_t n,
                                                              CalibEdges ed, double* __restrict__ partials) {
    __shared__ unsigned int s_cnt[kCalibSlots], s_true[kCalibSlots];
    __shared__ double red[kCalibSlots * 8];
    if (threadIdx.x < kCalibSlots) { s_cnt[threadIdx.x] = 0; s_true[threadIdx.x] = 0; }
    __syncthreads();
    CalibAcc acc;
    acc.init();
    const int64_t base = (int64_t)blockIdx.x * (kThreads * kStatEPT) + threadIdx.x;
    for (int i = 0; i < kStatEPT; ++i) {
        const int64_t v = base + (int64_t)i * kThreads;
        const bool live = v < n;
        double p = 0.0;
        bool t = false;
        if (live) { p = (double)ld_elem(prob + v); t = ld_elem(correct + v) != (TL)0; }
        acc.add(live, live ? calib_slot(p, ed) : -1, p, t, s_cnt, s_true);
    }
    __syncthreads();
    calib_flush(acc, s_cnt, s_true, red, partials);
}

// Fused per-image body of calibration_error (ace.py:96-121): for rater r and voxel v
//   correct = (ref[r, v] == pred[v]);  skipped when ref[r, v] == ignore_value (if given)
//   conf    = 1 / (1 + exp(-unc[v] * a + b))  -- platt_scale_confid (ace.py:42-46), evaluated in
//             the map's dtype as numpy does (fp32 map -> fp32 arithmetic; exp correctly rounded
//             through fp64)
template <typename T> struct Platt;
template <> struct Platt<float> {
    __device__ static __forceinline__ double conf(float u, double a, double b) {
        const float t = __fadd_rn(__fmul_rn(-u, (float)a), (float)b);
        const float e = (float)exp((double)t);
        return (double)__fdiv_rn(1.0f, __fadd_rn(1.0f, e));
    }
};
template <> struct Platt<double> {
    __device__ static __forceinline__ double conf(double u, double a, double b) {
        const double t = __dadd_rn(__dmul_rn(-u, a), b);
        return 1.0 / (1.0 + exp(t));
    }
};

template <typename T, typename TL>
__global__ void __launch_bounds__(kThreads) calib_fused_kernel(const T* __restrict__ unc,
                                                               const TL* __restrict__ pred,
                                                               const TL* __restrict__ refs, int64_t V,
                                                               int64_t R, double a, double b,
                                                               int has_ignore, long long ignore_value,
                                                               CalibEdges ed, double* __restrict__ partials) {
    __shared__ unsigned int s_cnt[kCalibSlots], s_true[kCalibSlots];
    __shared__ double red[kCalibSlots * 8];
    if (threadIdx.x < kCalibSlots) { s_cnt[threadIdx.x] = 0; s_true[threadIdx.x] = 0; }
    __syncthreads();
    CalibAcc acc;
    acc.init();
    const int64_t base = (int64_t)blockIdx.x * (kThreads * kStatEPT) + threadIdx.x;
    for (int i = 0; i < kStatEPT; ++i) {
        const int64_t v = base + (int64_t)i * kThreads;
        const bool inside = v < V;
        double p = 0.0;
        int slot = -1;
        TL pl = (TL)0;
        if (inside) {
            p = Platt<T>::conf(ld_elem(unc + v), a, b);
            slot = calib_slot(p, ed);
            pl = ld_elem(pred + v);
        }
        unsigned n_live = 0, n_true = 0;
        if (inside) {
            for (int64_t r = 0; r < R; ++r) {
                const TL rl = ld_elem(refs + r * V + v);
                const bool live = !(has_ignore && (long long)rl == ignore_value);
                n_live += live ? 1u : 0u;
                n_true += (live && rl == pl) ? 1u : 0u;
            }
        }
        acc.add_n(inside, slot, p, n_live, n_true, s_cnt, s_true);
    }
    __syncthreads();
    calib_flush(acc, s_cnt, s_true, red, partials);
}

// partials [n_blocks, K] -> out [K]  (K = 63): block k sums column k, rows striped over the threads,
// then a fixed-order block reduction (deterministic)
__global__ void __launch_bounds__(kThreads) calib_reduce_kernel(const double* __restrict__ partials,
                                                                int64_t n_blocks, int K,
                                                                double* __restrict__ out) {
    __shared__ double red[8];
    const int k = blockIdx.x;
    double acc[1] = {0.0};
    for (int64_t r = threadIdx.x; r < n_blocks; r += kThreads) acc[0] += partials[r * K + k];
    block_sum<1>(acc, red);
    if (threadIdx.x == 0) out[k] = acc[0];
}

// =============================================================================== confusion counts
// out[(ia * Nb + ib) * C * C + a * C + b] += #{ v : A[ia, v] == a and B[ib, v] == b }   (labels >= C or
// < 0 are dropped).  Every pair statistic of calculate_ged / torchmetrics' dice (test_3D.py:284-358) is a
// function of these integer matrices.  Block-private shared histogram, warp-aggregated atomics: exact.
constexpr int kMaxConfClasses = 32;

template <typename TL>
__global__ void __launch_bounds__(kThreads) confusion_kernel(const TL* __restrict__ A, int64_t stride_a,
                                                             const TL* __restrict__ B, int64_t stride_b,
                                                             int64_t Nb, int64_t V, int C, int64_t bpp,
                                                             unsigned long long* __restrict__ out) {
    __shared__ unsigned int sh[kMaxConfClasses * kMaxConfClasses];
    const int64_t pair = blockIdx.x / bpp, blk = blockIdx.x - pair * bpp;
    const int64_t ia = pair / Nb, ib = pair - ia * Nb;
    const int cc = C * C;
    for (int i = threadIdx.x; i < cc; i += kThreads) sh[i] = 0;
    __syncthreads();
    const TL* pa = A + ia * stride_a;
    const TL* pb = B + ib * stride_b;
    const int lane = threadIdx.x & 31;
    const int64_t base = blk * (int64_t)(kThreads * kStatEPT) + threadIdx.x;
    for (int i = 0; i < kStatEPT; ++i) {
        const int64_t v = base + (int64_t)i * kThreads;
        bool live = v < V;
        long long a = 0, b = 0;
        if (live) { a = (long long)ld_elem(pa + v); b = (long long)ld_elem(pb + v); }
        live = live && a >= 0 && a < C && b >= 0 && b < C;
        const unsigned active = __ballot_sync(0xffffffffu, live);
        if (live) {
            const unsigned code = (unsigned)(a * C + b);
            const unsigned peers = __match_any_sync(active, code);
            if (lane == __ffs(peers) - 1) atomicAdd(&sh[code], (unsigned)__popc(peers));
        }
    }
    __syncthreads();
    unsigned long long* dst = out + pair * cc;
    for (int i = threadIdx.x; i < cc; i += kThreads)
        if (sh[i]) atomicAdd(dst + i, (unsigned long long)sh[i]);
}

static int stat_grid(int64_t n_units) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t want = ceil_div(n_units, kThreads);
    const int64_t cap = (int64_t)sms * 8;   // 8 resident 256-thread CTAs per SM
    return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

// =============================================================================== seg_loss_terms
// The sums behind calculate_test_metrics (uncertainty_modeling/test_3D.py:250-281): per rater r of a label
// stack [R, V] against ONE probability map [C, V] (the mean softmax prediction),
//   inter[r][c] = sum_v x[c][v] [gt_r[v] == c]      SoftDiceLoss intersect   loss_modules.py:56-66, 86
//   count[r][c] = #{ v : gt_r[v] == c }             (x + onehot).sum = sumx + count            :87
//   sumx[c]     = sum_v x[c][v]                     (stored per rater row)
//   logp[r]     = sum_v log x[gt_r[v]][v]           torch.nn.NLLLoss(torch.log(x), gt) numerator, log in fp64
// in one sweep, fp64 fixed-order block-then-grid reduction.  A label outside [0, C) makes logp NaN (torch
// raises there).  partials [R, blocks, 3 CT + 1]; CT = C rounded up to 2, 4 or 8.
template <typename T, typename L, int CT>
__global__ void __launch_bounds__(kThreads) seg_loss_kernel(const T* __restrict__ probs, int64_t stride_c,
                                                            const L* __restrict__ labels, int64_t stride_r, int C,
                                                            int64_t V, int64_t bpm, double* __restrict__ partials) {
    constexpr int K = 3 * CT + 1;
    __shared__ double red[K * 8];
    const int64_t r = blockIdx.x / bpm;
    const int64_t blk = blockIdx.x - r * bpm;
    const L* lab = labels + r * stride_r;
    double acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.0;
    const int64_t base = blk * (int64_t)(kThreads * kStatEPT) + threadIdx.x;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll 2
    for (int i = 0; i < kStatEPT; ++i) {
        const int64_t v = base + (int64_t)i * kThreads;
        if (v < V) {
            const long long l = (long long)ld_elem(lab + v);
            double xl = qnan;
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                if (c < C) {
                    const double x = (double)ld_elem(probs + c * stride_c + v);
                    const bool hit = l == c;
                    acc[c] += hit ? x : 0.0;
                    acc[CT + c] += hit ? 1.0 : 0.0;
                    acc[2 * CT + c] += x;
                    xl = hit ? x : xl;
                }
            }
            acc[3 * CT] += log(xl);
        }
    }
    block_sum<K>(acc, red);
    if (threadIdx.x == 0) {
        double* dst = partials + (int64_t)blockIdx.x * K;
#pragma unroll
        for (int k = 0; k < K; ++k) dst[k] = acc[k];
    }
}

}  // namespace vb

using namespace vb;

extern "C" int values_count_nonzero(const void* data, int dtype, int64_t n,
                                    unsigned long long* count, void* stream) {
    if (n < 0) return set_error(VALUES_ERR_INVALID_ARG, "count_nonzero: n < 0");
    if (!count) return set_error(VALUES_ERR_INVALID_ARG, "count_nonzero: NULL output");
    if (n == 0) return VALUES_OK;
    if (!data) return set_error(VALUES_ERR_INVALID_ARG, "count_nonzero: NULL input");
    cudaStream_t st = (cudaStream_t)stream;
    int es = dtype == VALUES_U8 ? 1 : (dtype == VALUES_F32 || dtype == VALUES_I32) ? 4 : 8;
    const int grid = stat_grid(ceil_div(n * es, 16));
    switch (dtype) {
        case VALUES_U8:  count_nonzero_kernel<uint8_t><<<grid, kThreads, 0, st>>>((const uint8_t*)data, n, count); break;
        case VALUES_I32: count_nonzero_kernel<int32_t><<<grid, kThreads, 0, st>>>((const int32_t*)data, n, count); break;
        case VALUES_I64: count_nonzero_kernel<int64_t><<<grid, kThreads, 0, st>>>((const int64_t*)data, n, count); break;
        case VALUES_F32: count_nonzero_kernel<float><<<grid, kThreads, 0, st>>>((const float*)data, n, count); break;
        case VALUES_F64: count_nonzero_kernel<double><<<grid, kThreads, 0, st>>>((const double*)data, n, count); break;
        default: return set_error(VALUES_ERR_INVALID_ARG, "count_nonzero: dtype must be u8, i32, i64, f32 or f64");
    }
    return check_launch("count_nonzero_kernel");
}

// launches over a set of maps, kSetMax maps per launch: grid = (CTAs per map, maps)
template <typename Launch>
static int for_each_map_chunk(const void* const* maps, const int64_t* counts, int64_t n_maps, size_t es, int per_sm,
                              const char* what, Launch launch) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    for (int64_t m0 = 0; m0 < n_maps;) {
        MapSet set{};
        int k = 0;
        int64_t nmax = 0;
        for (; m0 < n_maps && k < kSetMax; ++m0) {
            if (counts[m0] < 0) return set_error(VALUES_ERR_INVALID_ARG, "%s: negative size", what);
            if (counts[m0] == 0) continue;
            if (!maps[m0]) return set_error(VALUES_ERR_INVALID_ARG, "%s: NULL input", what);
            set.p[k] = maps[m0]; set.n[k] = counts[m0];
            nmax = counts[m0] > nmax ? counts[m0] : nmax;
            ++k;
        }
        if (k == 0) continue;
        // every CTA in the first wave: per_sm resident CTAs per SM shared by the maps of the launch
        const int64_t want = ceil_div(ceil_div(nmax * (int64_t)es, 64), kThreads);
        const int64_t cap = ((int64_t)sms * per_sm + k - 1) / k;
        const int64_t gx = want < 1 ? 1 : (want > cap ? cap : want);
        launch(set, dim3((unsigned)gx, (unsigned)k));
        const int rc = check_launch(what);
        if (rc) return rc;
    }
    return VALUES_OK;
}

static int radix_histogram_impl(const void* const* maps, const int64_t* counts, int64_t n_maps, int dtype,
                                uint64_t prefix, int prefix_bits, int digit_bits, unsigned long long* hist,
                                const unsigned long long* state, void* stream) {
    if (dtype != VALUES_F32 && dtype != VALUES_F64)
        return set_error(VALUES_ERR_INVALID_ARG, "radix_histogram: dtype must be f32 or f64");
    const int bits = dtype == VALUES_F32 ? 32 : 64;
    if (n_maps < 0 || digit_bits < 1 || digit_bits > kMaxDigitBits || prefix_bits < 0 ||
        prefix_bits + digit_bits > bits)
        return set_error(VALUES_ERR_INVALID_ARG, "radix_histogram: bad sizes (digit_bits 1..%d, "
                         "prefix_bits + digit_bits <= %d)", kMaxDigitBits, bits);
    if (!hist) return set_error(VALUES_ERR_INVALID_ARG, "radix_histogram: NULL histogram");
    if (n_maps == 0) return VALUES_OK;
    if (!maps || !counts) return set_error(VALUES_ERR_INVALID_ARG, "radix_histogram: NULL input");
    cudaStream_t st = (cudaStream_t)stream;
    const int shift = bits - prefix_bits - digit_bits;
    // 32 KB of shared memory per CTA: six resident CTAs per SM
    return for_each_map_chunk(maps, counts, n_maps, bits / 8, 6, "radix_hist_kernel", [&](const MapSet& set, dim3 grid) {
        if (dtype == VALUES_F32)
            radix_hist_kernel<float><<<grid, kThreads, 0, st>>>(set, prefix, prefix_bits, shift, digit_bits, hist, state);
        else
            radix_hist_kernel<double><<<grid, kThreads, 0, st>>>(set, prefix, prefix_bits, shift, digit_bits, hist, state);
    });
}

extern "C" int values_radix_histogram(const void* data, int dtype, int64_t n, uint64_t prefix,
                                      int prefix_bits, int digit_bits, unsigned long long* hist,
                                      void* stream) {
    if (n > 0 && !data) return set_error(VALUES_ERR_INVALID_ARG, "radix_histogram: NULL input");
    return radix_histogram_impl(&data, &n, 1, dtype, prefix, prefix_bits, digit_bits, hist, nullptr, stream);
}

extern "C" int values_radix_histogram_dev(const void* data, int dtype, int64_t n, const unsigned long long* state,
                                          int digit_bits, unsigned long long* hist, void* stream) {
    if (!state) return set_error(VALUES_ERR_INVALID_ARG, "radix_histogram_dev: NULL state");
    if (n > 0 && !data) return set_error(VALUES_ERR_INVALID_ARG, "radix_histogram: NULL input");
    return radix_histogram_impl(&data, &n, 1, dtype, 0, 0, digit_bits, hist, state, stream);
}

extern "C" int values_radix_histogram_set(const void* const* maps_host, const int64_t* counts_host, int64_t n_maps,
                                          int dtype, uint64_t prefix, int prefix_bits,
                                          const unsigned long long* state, int digit_bits,
                                          unsigned long long* hist, void* stream) {
    return radix_histogram_impl(maps_host, counts_host, n_maps, dtype, state ? 0 : prefix, state ? 0 : prefix_bits,
                                digit_bits, hist, state, stream);
}

extern "C" int values_radix_select(unsigned long long* hist, int digit_bits, unsigned long long* state, void* stream) {
    if (!hist || !state || digit_bits < 1 || digit_bits > kMaxDigitBits)
        return set_error(VALUES_ERR_INVALID_ARG, "radix_select: bad arguments");
    radix_select_kernel<<<1, kThreads, 0, (cudaStream_t)stream>>>(hist, digit_bits, state);
    return check_launch("radix_select_kernel");
}

static int min_key_above_impl(const void* const* maps, const int64_t* counts, int64_t n_maps, int dtype, uint64_t key,
                              unsigned long long* out, void* stream) {
    if (dtype != VALUES_F32 && dtype != VALUES_F64)
        return set_error(VALUES_ERR_INVALID_ARG, "min_key_above: dtype must be f32 or f64");
    if (n_maps < 0 || !out) return set_error(VALUES_ERR_INVALID_ARG, "min_key_above: bad arguments");
    if (n_maps == 0) return VALUES_OK;
    if (!maps || !counts) return set_error(VALUES_ERR_INVALID_ARG, "min_key_above: NULL input");
    cudaStream_t st = (cudaStream_t)stream;
    // element loads: sized as if every element were 64 bytes' worth of work per thread quarter
    return for_each_map_chunk(maps, counts, n_maps, 16, 8, "min_key_above_kernel", [&](const MapSet& set, dim3 grid) {
        if (dtype == VALUES_F32)
            min_key_above_kernel<float><<<grid, kThreads, 0, st>>>(set, key, out);
        else
            min_key_above_kernel<double><<<grid, kThreads, 0, st>>>(set, key, out);
    });
}

extern "C" int values_min_key_above(const void* data, int dtype, int64_t n, uint64_t key,
                                    unsigned long long* out, void* stream) {
    if (n < 0) return set_error(VALUES_ERR_INVALID_ARG, "min_key_above: bad arguments");
    if (n > 0 && !data) return set_error(VALUES_ERR_INVALID_ARG, "min_key_above: NULL input");
    return min_key_above_impl(&data, &n, 1, dtype, key, out, stream);
}

extern "C" int values_min_key_above_set(const void* const* maps_host, const int64_t* counts_host, int64_t n_maps,
                                        int dtype, uint64_t key, unsigned long long* out, void* stream) {
    return min_key_above_impl(maps_host, counts_host, n_maps, dtype, key, out, stream);
}

extern "C" size_t values_pair_moments_workspace_bytes(int64_t M, int64_t V) {
    if (M <= 0 || V <= 0) return 0;
    return (size_t)(M * ceil_div(V, kThreads * kStatEPT) * 5) * sizeof(double);
}

extern "C" int values_pair_moments(const void* a, int dtype_a, int64_t stride_a, const void* b,
                                   int dtype_b, int64_t stride_b, int64_t M, int64_t V,
                                   const double* shift, double* out, void* workspace,
                                   size_t workspace_bytes, void* stream) {
    if (M < 0 || V < 0) return set_error(VALUES_ERR_INVALID_ARG, "pair_moments: bad sizes");
    if ((dtype_a != VALUES_F32 && dtype_a != VALUES_F64) || (dtype_b != VALUES_F32 && dtype_b != VALUES_F64))
        return set_error(VALUES_ERR_INVALID_ARG, "pair_moments: dtypes must be f32 or f64");
    if (M == 0) return VALUES_OK;
    if (!a || !b || !out) return set_error(VALUES_ERR_INVALID_ARG, "pair_moments: NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (V == 0) {
        if (cudaMemsetAsync(out, 0, (size_t)M * 5 * sizeof(double), st) != cudaSuccess)
            return set_error(VALUES_ERR_CUDA, "pair_moments: memset failed");
        return VALUES_OK;
    }
    const size_t need = values_pair_moments_workspace_bytes(M, V);
    if (!workspace || workspace_bytes < need)
        return set_error(VALUES_ERR_WORKSPACE, "pair_moments: workspace %zu < %zu", workspace_bytes, need);
    const int64_t bpm = ceil_div(V, kThreads * kStatEPT);
    if (bpm * M > 0x7fffffffLL) return set_error(VALUES_ERR_UNSUPPORTED, "pair_moments: grid too large");
    const unsigned grid = (unsigned)(bpm * M);
    double* partials = reinterpret_cast<double*>(workspace);
    if (dtype_a == VALUES_F32 && dtype_b == VALUES_F32)
        pair_moments_kernel<float, float><<<grid, kThreads, 0, st>>>((const float*)a, stride_a, (const float*)b, stride_b, V, bpm, shift, partials);
    else if (dtype_a == VALUES_F32)
        pair_moments_kernel<float, double><<<grid, kThreads, 0, st>>>((const float*)a, stride_a, (const double*)b, stride_b, V, bpm, shift, partials);
    else if (dtype_b == VALUES_F32)
        pair_moments_kernel<double, float><<<grid, kThreads, 0, st>>>((const double*)a, stride_a, (const float*)b, stride_b, V, bpm, shift, partials);
    else
        pair_moments_kernel<double, double><<<grid, kThreads, 0, st>>>((const double*)a, stride_a, (const double*)b, stride_b, V, bpm, shift, partials);
    int rc = check_launch("pair_moments_kernel");
    if (rc) return rc;
    return launch_reduce_partials(partials, M, bpm, 5, out, st);
}

extern "C" size_t values_calib_bins_workspace_bytes(int64_t n) {
    if (n <= 0) return 0;
    return (size_t)(ceil_div(n, kThreads * kStatEPT) * 3 * kCalibSlots) * sizeof(double);
}

static int calib_edges(const double* edges_host, int n_bins, CalibEdges& ed) {
    if (n_bins != kCalibBins || !edges_host)
        return set_error(VALUES_ERR_UNSUPPORTED, "calib_bins: n_bins must be %d (ace.py:51) with %d edges",
                         kCalibBins, kCalibSlots);
    for (int k = 0; k < kCalibSlots; ++k) {
        ed.e[k] = edges_host[k];
        if (k && !(ed.e[k] > ed.e[k - 1]))
            return set_error(VALUES_ERR_INVALID_ARG, "calib_bins: edges must increase");
    }
    return VALUES_OK;
}

static int calib_finish(double* partials, int64_t blocks, double* out, cudaStream_t st) {
    int rc = check_launch("calib kernel");
    if (rc) return rc;
    calib_reduce_kernel<<<3 * kCalibSlots, kThreads, 0, st>>>(partials, blocks, 3 * kCalibSlots, out);
    return check_launch("calib_reduce_kernel");
}

extern "C" int values_calib_bins(const void* prob, int dtype, const void* correct, int label_dtype,
                                 int64_t n, const double* edges_host, int n_bins, double* out,
                                 void* workspace, size_t workspace_bytes, void* stream) {
    CalibEdges ed;
    int rc = calib_edges(edges_host, n_bins, ed);
    if (rc) return rc;
    if (n < 0 || !out) return set_error(VALUES_ERR_INVALID_ARG, "calib_bins: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        if (cudaMemsetAsync(out, 0, 3 * kCalibSlots * sizeof(double), st) != cudaSuccess)
            return set_error(VALUES_ERR_CUDA, "calib_bins: memset failed");
        return VALUES_OK;
    }
    if (!prob || !correct) return set_error(VALUES_ERR_INVALID_ARG, "calib_bins: NULL input");
    const size_t need = values_calib_bins_workspace_bytes(n);
    if (!workspace || workspace_bytes < need)
        return set_error(VALUES_ERR_WORKSPACE, "calib_bins: workspace %zu < %zu", workspace_bytes, need);
    const int64_t blocks = ceil_div(n, kThreads * kStatEPT);
    if (blocks > 0x7fffffffLL) return set_error(VALUES_ERR_UNSUPPORTED, "calib_bins: grid too large");
    double* partials = reinterpret_cast<double*>(workspace);
#define VB_CALIB(TP, TL) calib_bins_kernel<TP, TL><<<(unsigned)blocks, kThreads, 0, st>>>((const TP*)prob, (const TL*)correct, n, ed, partials)
    if (dtype == VALUES_F32 && label_dtype == VALUES_U8) VB_CALIB(float, uint8_t);
    else if (dtype == VALUES_F32 && label_dtype == VALUES_I32) VB_CALIB(float, int32_t);
    else if (dtype == VALUES_F32 && label_dtype == VALUES_I64) VB_CALIB(float, int64_t);
    else if (dtype == VALUES_F64 && label_dtype == VALUES_U8) VB_CALIB(double, uint8_t);
    else if (dtype == VALUES_F64 && label_dtype == VALUES_I32) VB_CALIB(double, int32_t);
    else if (dtype == VALUES_F64 && label_dtype == VALUES_I64) VB_CALIB(double, int64_t);
    else return set_error(VALUES_ERR_INVALID_ARG, "calib_bins: prob f32/f64, labels u8/i32/i64");
#undef VB_CALIB
    return calib_finish(partials, blocks, out, st);
}

extern "C" int values_calib_bins_fused(const void* unc, int dtype, const void* pred_seg,
                                       const void* ref_segs, int label_dtype, int64_t V, int64_t R,
                                       double a, double b, int has_ignore, int64_t ignore_value,
                                       const double* edges_host, int n_bins, double* out,
                                       void* workspace, size_t workspace_bytes, void* stream) {
    CalibEdges ed;
    int rc = calib_edges(edges_host, n_bins, ed);
    if (rc) return rc;
    if (V < 0 || R < 0 || !out) return set_error(VALUES_ERR_INVALID_ARG, "calib_bins_fused: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (V == 0 || R == 0) {
        if (cudaMemsetAsync(out, 0, 3 * kCalibSlots * sizeof(double), st) != cudaSuccess)
            return set_error(VALUES_ERR_CUDA, "calib_bins_fused: memset failed");
        return VALUES_OK;
    }
    if (!unc || !pred_seg || !ref_segs) return set_error(VALUES_ERR_INVALID_ARG, "calib_bins_fused: NULL input");
    const size_t need = values_calib_bins_workspace_bytes(V);
    if (!workspace || workspace_bytes < need)
        return set_error(VALUES_ERR_WORKSPACE, "calib_bins_fused: workspace %zu < %zu", workspace_bytes, need);
    const int64_t blocks = ceil_div(V, kThreads * kStatEPT);
    if (blocks > 0x7fffffffLL) return set_error(VALUES_ERR_UNSUPPORTED, "calib_bins_fused: grid too large");
    double* partials = reinterpret_cast<double*>(workspace);
#define VB_CALIBF(T, TL) calib_fused_kernel<T, TL><<<(unsigned)blocks, kThreads, 0, st>>>((const T*)unc, (const TL*)pred_seg, (const TL*)ref_segs, V, R, a, b, has_ignore, (long long)ignore_value, ed, partials)
    if (dtype == VALUES_F32 && label_dtype == VALUES_U8) VB_CALIBF(float, uint8_t);
    else if (dtype == VALUES_F32 && label_dtype == VALUES_I32) VB_CALIBF(float, int32_t);
    else if (dtype == VALUES_F32 && label_dtype == VALUES_I64) VB_CALIBF(float, int64_t);
    else if (dtype == VALUES_F64 && label_dtype == VALUES_U8) VB_CALIBF(double, uint8_t);
    else if (dtype == VALUES_F64 && label_dtype == VALUES_I32) VB_CALIBF(double, int32_t);
    else if (dtype == VALUES_F64 && label_dtype == VALUES_I64) VB_CALIBF(double, int64_t);
    else return set_error(VALUES_ERR_INVALID_ARG, "calib_bins_fused: map f32/f64, labels u8/i32/i64");
#undef VB_CALIBF
    return calib_finish(partials, blocks, out, st);
}

extern "C" int values_confusion_counts(const void* labels_a, int64_t Na, int64_t stride_a,
                                       const void* labels_b, int64_t Nb, int64_t stride_b,
                                       int label_dtype, int64_t V, int n_classes,
                                       unsigned long long* out, void* stream) {
    if (Na < 0 || Nb < 0 || V < 0 || n_classes < 1 || n_classes > kMaxConfClasses)
        return set_error(VALUES_ERR_INVALID_ARG, "confusion_counts: bad sizes (1 <= n_classes <= %d)", kMaxConfClasses);
    if (!out) return set_error(VALUES_ERR_INVALID_ARG, "confusion_counts: NULL output");
    if (Na == 0 || Nb == 0 || V == 0) return VALUES_OK;
    if (!labels_a || !labels_b) return set_error(VALUES_ERR_INVALID_ARG, "confusion_counts: NULL input");
    const int64_t bpp = ceil_div(V, kThreads * kStatEPT);
    if (bpp * Na * Nb > 0x7fffffffLL) return set_error(VALUES_ERR_UNSUPPORTED, "confusion_counts: grid too large");
    const unsigned grid = (unsigned)(bpp * Na * Nb);
    cudaStream_t st = (cudaStream_t)stream;
    switch (label_dtype) {
        case VALUES_U8:
            confusion_kernel<uint8_t><<<grid, kThreads, 0, st>>>((const uint8_t*)labels_a, stride_a, (const uint8_t*)labels_b, stride_b, Nb, V, n_classes, bpp, out);
            break;
        case VALUES_I32:
            confusion_kernel<int32_t><<<grid, kThreads, 0, st>>>((const int32_t*)labels_a, stride_a, (const int32_t*)labels_b, stride_b, Nb, V, n_classes, bpp, out);
            break;
        case VALUES_I64:
            confusion_kernel<int64_t><<<grid, kThreads, 0, st>>>((const int64_t*)labels_a, stride_a, (const int64_t*)labels_b, stride_b, Nb, V, n_classes, bpp, out);
            break;
        default: return set_error(VALUES_ERR_INVALID_ARG, "confusion_counts: labels must be u8, i32 or i64");
    }
    return check_launch("confusion_kernel");
}

static int seg_loss_ct(int C) { return C <= 2 ? 2 : (C <= 4 ? 4 : 8); }

extern "C" size_t values_seg_loss_workspace_bytes(int64_t R, int C, int64_t V) {
    if (R <= 0 || V <= 0 || C <= 0 || C > 8) return 0;
    return (size_t)(R * ceil_div(V, kThreads * kStatEPT) * (3 * seg_loss_ct(C) + 1)) * sizeof(double);
}

template <typename T, typename L>
static void launch_seg_loss(const void* probs, int64_t stride_c, const void* labels, int64_t stride_r, int C,
                            int64_t V, int64_t bpm, unsigned grid, double* partials, cudaStream_t st) {
    const T* p = (const T*)probs;
    const L* l = (const L*)labels;
    switch (seg_loss_ct(C)) {
        case 2: seg_loss_kernel<T, L, 2><<<grid, kThreads, 0, st>>>(p, stride_c, l, stride_r, C, V, bpm, partials); break;
        case 4: seg_loss_kernel<T, L, 4><<<grid, kThreads, 0, st>>>(p, stride_c, l, stride_r, C, V, bpm, partials); break;
        default: seg_loss_kernel<T, L, 8><<<grid, kThreads, 0, st>>>(p, stride_c, l, stride_r, C, V, bpm, partials); break;
    }
}

extern "C" int values_seg_loss_terms(const void* probs, int dtype, int64_t stride_c, const void* labels,
                                     int label_dtype, int64_t stride_r, int64_t R, int C, int64_t V,
                                     double* out, void* workspace, size_t workspace_bytes, void* stream) {
    if (R < 0 || V < 0 || C < 1) return set_error(VALUES_ERR_INVALID_ARG, "seg_loss_terms: bad sizes");
    if (C > 8) return set_error(VALUES_ERR_UNSUPPORTED, "seg_loss_terms: C = %d > 8 classes", C);
    if (dtype != VALUES_F32 && dtype != VALUES_F64)
        return set_error(VALUES_ERR_INVALID_ARG, "seg_loss_terms: probabilities must be f32 or f64");
    if (R == 0) return VALUES_OK;
    if (!out) return set_error(VALUES_ERR_INVALID_ARG, "seg_loss_terms: NULL output");
    const int K = 3 * seg_loss_ct(C) + 1;
    cudaStream_t st = (cudaStream_t)stream;
    if (V == 0) {
        if (cudaMemsetAsync(out, 0, (size_t)R * K * sizeof(double), st) != cudaSuccess)
            return set_error(VALUES_ERR_CUDA, "seg_loss_terms: memset failed");
        return VALUES_OK;
    }
    if (!probs || !labels) return set_error(VALUES_ERR_INVALID_ARG, "seg_loss_terms: NULL input");
    const size_t need = values_seg_loss_workspace_bytes(R, C, V);
    if (!workspace || workspace_bytes < need)
        return set_error(VALUES_ERR_WORKSPACE, "seg_loss_terms: workspace %zu < %zu", workspace_bytes, need);
    const int64_t bpm = ceil_div(V, kThreads * kStatEPT);
    if (bpm * R > 0x7fffffffLL) return set_error(VALUES_ERR_UNSUPPORTED, "seg_loss_terms: grid too large");
    const unsigned grid = (unsigned)(bpm * R);
    double* partials = reinterpret_cast<double*>(workspace);
#define VB_SEGLOSS(T) do { switch (label_dtype) { \
        case VALUES_U8: launch_seg_loss<T, uint8_t>(probs, stride_c, labels, stride_r, C, V, bpm, grid, partials, st); break; \
        case VALUES_I32: launch_seg_loss<T, int32_t>(probs, stride_c, labels, stride_r, C, V, bpm, grid, partials, st); break; \
        case VALUES_I64: launch_seg_loss<T, int64_t>(probs, stride_c, labels, stride_r, C, V, bpm, grid, partials, st); break; \
        default: return set_error(VALUES_ERR_INVALID_ARG, "seg_loss_terms: labels must be u8, i32 or i64"); } } while (0)
    if (dtype == VALUES_F32) VB_SEGLOSS(float); else VB_SEGLOSS(double);
#undef VB_SEGLOSS
    int rc = check_launch("seg_loss_kernel");
    if (rc) return rc;
    return launch_reduce_partials(partials, R, bpm, K, out, st);
}
