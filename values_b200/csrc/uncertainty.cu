// K1: fused N x C reduction -> predictive entropy, expected entropy, mutual information,
// arg-max of the mean (and of every sample), plus the image-level / threshold score
// numerators, in ONE sweep over the softmax stack.
//
// Reference semantics (uncertainty_modeling/test_3D.py:486-518), reproduced per voxel:
//   m_c   = (sum_n p[n,c]) / N            sequential sum in the input dtype, true division
//   PE    = -sum_c m_c*log(m_c)           term in the input dtype, fp32 accumulator, class order,
//   H_n   = -sum_c p[n,c]*log(p[n,c])     NaN terms (0*log0, log of negatives) are skipped
//   EE    = (sum_n H_n) / N               fp32, sample order
//   MI    = PE - EE
// fp64 stacks (the reference's 3-D path) keep exactly this order: one fp32 accumulator H_n per sample,
// classes in index order, EE their sequential fp32 sum.  The fp32 / bf16 kernels do NOT: they stream
// class-outer / sample-inner and add every p*log(p) of a class into one fp32 accumulator, the classes
// then into E, with a polynomial (fp32) or MUFU (bf16) logarithm -- the same N*C terms in another
// order.  EE, and so MI = PE - EE, therefore differ from the reference by rounding only: measured
// <= 1.2e-6 absolute (8 ulp of EE) on the BASELINE shapes, bounded by ~N*C*2^-24*max|p log p|; the
// parity tests hold 1e-5 |ref| + 1e-6, and tests/test_gpu_parity_counts.py counts what that does to
// threshold masks (0-2 voxels per 2.1 M at the median threshold).
// HBM-bound: algorithmic bytes per voxel = N*C*sizeof(T) + 3*4 (+1 arg-max byte).
#include "common.cuh"
#include "tma_host.cuh"

namespace vb {

// Kernel choice per call (values_uncertainty_fused `variant`; 0 = automatic, the others exist so that
// tests can run two implementations of the same arithmetic against each other -- every variant gives
// bit-identical outputs): 1 = register-stream kernel instead of the bulk-copy ring, 2 = ring with
// 4-row stages where 8-row stages are the default, 3 = 8-row stages x 3 at two CTAs per SM,
// 4 = (fp64 stacks) sample-outer kernel instead of the class-outer ring.
enum { K1_AUTO = 0, K1_STREAM = 1, K1_RING4 = 2, K1_RING8X3 = 3, K1_SAMPLE_OUTER = 4, K1_VARIANTS = 5 };

struct K1Params {
    const void* probs;
    int64_t N, C, V, sb, sn, sc;
    int64_t blocks_per_vol;
    float* pe;
    float* ee;
    float* mi;
    uint8_t* amax;
    uint8_t* samax;
    double* partials;  // [B, blocks_per_vol, 9] or nullptr
    double* scores;    // row (b, map) at scores[(3 b + map) * score_stride + 0..2]: written by the last CTA of each volume
    int64_t score_stride;
    int variant, tiles_per_cta;   // per-call kernel choice (K1_*), tiles per CTA (0 = automatic)
    unsigned int* counters;  // [B] arrival tickets (zeroed before the launch)
    float thr_f[3];    // thresholds rounded to fp32: numpy compares an fp32 map in fp32
    int has_thr;
    int need_ent;
    int iter;          // voxel tiles per CTA (stream kernel)
    float inv_n;       // RN(1/N)
    int64_t so;        // elements between consecutive volumes in pe / ee / mi (>= V)
};

// ---- raw vector loads of VEC elements (16 / 8 / smaller bytes)
template <typename T, int VEC>
__device__ __forceinline__ void load_elems(const T* p, typename In<T>::acc_t (&o)[VEC]) {
    constexpr int BYTES = VEC * (int)sizeof(T);
    if constexpr (VEC == In<T>::VEC) {
        In<T>::load_vec(p, o);
    } else if constexpr (BYTES == 8 && sizeof(T) == 2) {  // 4 x bf16
        uint2 r;
        asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];"
                     : "=r"(r.x), "=r"(r.y) : "l"(p));
        o[0] = __uint_as_float(r.x << 16); o[1] = __uint_as_float(r.x & 0xffff0000u);
        o[2] = __uint_as_float(r.y << 16); o[3] = __uint_as_float(r.y & 0xffff0000u);
    } else {
#pragma unroll
        for (int j = 0; j < VEC; ++j) o[j] = In<T>::load_one(p + j);
    }
}

// log(p) for finite normal p > 0, <= 1.06 ulp over every mantissa (exhaustive check in
// tools/fit_log_poly.py).  CUDA's logf spends ~35 issue slots per element on special cases
// (zero, negative, denormal, inf); here the caller's `p > 0` predicate already covers what the
// NaN-skip rule needs, which keeps K1 under the HBM roofline instead of issue-bound:
//   p = m * 2^e, m in [2/3, 4/3), f = m - 1;  log(p) = e*ln2 + f + f^2*Q(f), Q degree 7.
// Denormals give a finite wrong log; times p (< 1.2e-38) the term is below any tolerance.
__device__ __forceinline__ float fast_logf(float p) {
    const int i = __float_as_int(p);
    const int e = (i - 0x3f2aaaab) & 0xff800000;      // exponent field, scaled by 2^23
    const float f = __int_as_float(i - e) - 1.0f;
    float q = 0.13979104161262512f;
    q = fmaf(q, f, -0.15397904813289642f);
    q = fmaf(q, f, 0.14004801213741302f);
    q = fmaf(q, f, -0.1641434133052826f);
    q = fmaf(q, f, 0.20010659098625183f);
    q = fmaf(q, f, -0.2500789761543274f);
    q = fmaf(q, f, 0.3333320617675781f);
    q = fmaf(q, f, -0.49999934434890747f);
    const float r = fmaf(q * f, f, f);
    return fmaf((float)e, 8.262958317573066e-08f /* ln2 * 2^-23 */, r);
}

// ------------------------------------------------------------------ fp64 log for the fp64 path
// The reference computes p*log(p) in the INPUT dtype and adds it into an fp32 accumulator
// (test_3D.py:490-506); for fp64 stacks (its 3D path) CUDA's log(double) -- ~1 ulp, every special
// case handled inline -- was what bounded the kernel (0.28 of the HBM peak).  The terms only have
// to survive the rounding into fp32, so a table-driven log with <= 5e-15 relative error does:
//   p = m * 2^k, m in [~sqrt(2)/2, ~sqrt(2)): the split point has a zero low word, so k, the table
//   index and the mantissa rebias are 32-bit integer operations on the HIGH word of p alone;
//   interval i = top 9 mantissa bits of m (513 intervals); r = fma(m, inv_i, -1) with inv_i ~ 1/centre_i
//   rounded to 12 bits (|r| <= 2^-9; the two intervals around 1 use inv = 1, so r = m - 1 exactly and
//   the log keeps its relative accuracy near p = 1);
//   log(p) = k*ln2 + (-log(inv_i)) + log1p(r),  log1p by its series to r^5.
// Checked against a 120-bit reference over 5e5 values from 1e-300 to 8 incl. 1 +- 1e-15
// (tools/check_log64.py, which also writes log64_table.inc): max relative error 4.7e-15, i.e. a term
// changes its fp32 rounding about once in 1e7.
__device__ const double2 kLog64Tab[513] = {
#include "log64_table.inc"
};
constexpr unsigned int kLog64Lo = 0x3FE6A09Eu;     // high word of the split point
constexpr int kLog64Base = 0x7FCD4;                // kLog64Lo >> 11: table index of the first interval

// exponent, rebiased mantissa and table index from the high word (any bit pattern gives an index inside the table)
__device__ __forceinline__ void log64_split(int hi, int& k, int& mhi, int& idx) {
    k = (int)((unsigned int)hi - kLog64Lo) >> 20;
    mhi = (int)((unsigned int)hi - ((unsigned int)k << 20));
    idx = (mhi >> 11) - kLog64Base;
}
// (double)k without the XU-pipe conversion: 2^52 + 2^31 + k is exact in the low word of a double
__device__ __forceinline__ double log64_k(int k) {
    return __hiloint2double(0x43300000, k ^ (int)0x80000000) - 4503601774854144.0;
}
__device__ __forceinline__ double log64_finish(int k, double m, double2 t) {
    const double r = fma(m, t.x, -1.0);
    const double r2 = r * r;
    // log1p(r) = r + r^2 (-1/2 + r/3 + r^2 (-1/4 + r/5)): Estrin, three dependent levels instead of five
    const double q = fma(r2, fma(0.2, r, -0.25), fma(1.0 / 3.0, r, -0.5));
    const double l1p = fma(r2, q, r);
    return fma(log64_k(k), 0.6931471805599453, t.y + l1p);
}
// any input: zero, negative, NaN, inf and subnormal values take log()
__device__ __forceinline__ double fast_log_f64(double p) {
    const int hi = __double2hiint(p);
    if ((unsigned int)(hi - 0x00100000) >= 0x7fe00000u) return log(p);
    int k, mhi, idx;
    log64_split(hi, k, mhi, idx);
    return log64_finish(k, __hiloint2double(mhi, __double2loint(p)), __ldg(&kLog64Tab[idx]));
}

// the same, table read from the CTA's shared-memory copy (`tab` = its shared-memory address)
__device__ __forceinline__ double fast_log_f64_tab(double p, uint32_t tab) {
    const int hi = __double2hiint(p);
    if ((unsigned int)(hi - 0x00100000) >= 0x7fe00000u) return log(p);
    int k, mhi, idx;
    log64_split(hi, k, mhi, idx);
    double2 t;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(t.x), "=d"(t.y) : "r"(tab + (uint32_t)idx * 16u));
    return log64_finish(k, __hiloint2double(mhi, __double2loint(p)), t);
}

// acc += p*log(p), NaN terms skipped (test_3D.py:490-494, 500-504).  A term is NaN exactly when
// p is 0 (0 * -inf), negative (log -> NaN) or NaN, i.e. when !(p > 0); +inf is NOT skipped.
__device__ __forceinline__ void accum_term(float& acc, float p) {
    const float t = fmaf(p, fast_logf(p), acc);
    acc = (p > 0.f) ? t : acc;  // select, not a branch
}
__device__ __forceinline__ void accum_term(float& acc, double p) {
    const double t = p * fast_log_f64(p);
    acc = (t == t) ? (float)((double)acc + t) : acc;  // add in fp64, round into fp32
}

// The N x C entropy terms of the fp64 ring kernel: no special-case branch per element.  The table-driven
// log gives a FINITE value for every bit pattern (log64_split keeps the index inside the table), so for
// finite p >= +0 the term p * log(p) is the reference's contribution (exactly 0 for p = 0, below any fp32
// rounding for subnormals).  Inputs the NaN-skip rule treats differently -- negative, -0, inf, NaN -- are
// detected from the unsigned maximum of the high words and the voxel's entropy is then recomputed by
// k1_entropy_exact_samples (cold).  `tab` = shared-memory address of the table copy.
__device__ __forceinline__ void accum_term_fast(float& acc, double p, unsigned int& hi_max, uint32_t tab) {
    const int hi = __double2hiint(p);
    hi_max = max(hi_max, (unsigned int)hi);
    int k, mhi, idx;
    log64_split(hi, k, mhi, idx);
    double2 t;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(t.x), "=d"(t.y) : "r"(tab + (uint32_t)idx * 16u));
    const double lg = log64_finish(k, __hiloint2double(mhi, __double2loint(p)), t);
    acc = (float)((double)acc + p * lg);   // add in fp64, round into fp32 (test_3D.py:500-506 on an fp64 stack)
}

template <typename A>
__device__ __forceinline__ void argmax_update(A v, int c, A& best, int& idx) {
    // first max wins; NaN counts as maximal (np.argmax / torch.argmax)
    if (v > best || (v != v && best == best)) { best = v; idx = c; }
}

template <int VEC>
__device__ __forceinline__ void store_u8(uint8_t* dst, const int (&idx)[VEC]) {
    if constexpr (VEC == 8) {
        uint2 w;
        w.x = idx[0] | (idx[1] << 8) | (idx[2] << 16) | (idx[3] << 24);
        w.y = idx[4] | (idx[5] << 8) | (idx[6] << 16) | (idx[7] << 24);
        *reinterpret_cast<uint2*>(dst) = w;
    } else if constexpr (VEC == 4) {
        *reinterpret_cast<uint32_t*>(dst) =
            idx[0] | (idx[1] << 8) | (idx[2] << 16) | (idx[3] << 24);
    } else if constexpr (VEC == 2) {
        *reinterpret_cast<uint16_t*>(dst) = (uint16_t)(idx[0] | (idx[1] << 8));
    } else {
#pragma unroll
        for (int j = 0; j < VEC; ++j) dst[j] = (uint8_t)idx[j];
    }
}

template <int VEC>
__device__ __forceinline__ void store_f32(float* dst, const float (&v)[VEC]) {
    if constexpr (VEC % 4 == 0) {
#pragma unroll
        for (int j = 0; j < VEC; j += 4)
            *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else if constexpr (VEC == 2) {
        *reinterpret_cast<float2*>(dst) = make_float2(v[0], v[1]);
    } else {
#pragma unroll
        for (int j = 0; j < VEC; ++j) dst[j] = v[j];
    }
}

// Epilogue shared by both kernels: maps from the class sums S and the entropy sum E.
template <typename A, int VEC, int CT, typename GetS>
__device__ __forceinline__ void k1_epilogue(const K1Params& prm, int64_t b, int64_t v0, int C,
                                            GetS getS, const float (&E)[VEC], double (&part)[9]) {
    const A invN_den = (A)prm.N;
    float pe[VEC], ee[VEC], mi[VEC];
    int idx[VEC];
    A best[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) { pe[j] = 0.f; idx[j] = 0; }
    const int nc = CT > 0 ? CT : C;
#pragma unroll
    for (int c = 0; c < nc; ++c) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            const A m = getS(c, j) / invN_den;  // mean: true division (torch.mean / np.mean)
            if (c == 0) best[j] = m; else argmax_update(m, c, best[j], idx[j]);
            if (CT > 0 || prm.need_ent) accum_term(pe[j], m);
        }
    }
    const float nf = (float)prm.N;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        pe[j] = -pe[j];
        ee[j] = -E[j] / nf;
        mi[j] = pe[j] - ee[j];
    }
    const int64_t o = b * prm.so + v0;
    if (prm.pe) store_f32<VEC>(prm.pe + o, pe);
    if (prm.ee) store_f32<VEC>(prm.ee + o, ee);
    if (prm.mi) store_f32<VEC>(prm.mi + o, mi);
    if (prm.amax) store_u8<VEC>(prm.amax + b * prm.V + v0, idx);
    if (prm.partials) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            const double m3[3] = {(double)pe[j], (double)ee[j], (double)mi[j]};
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                part[3 * k] += m3[k];
                const float mk = k == 0 ? pe[j] : k == 1 ? ee[j] : mi[j];
                if (prm.has_thr && mk >= prm.thr_f[k]) { part[3 * k + 1] += m3[k]; part[3 * k + 2] += 1.0; }
            }
        }
    }
}

// Block partial -> workspace; the LAST CTA of a volume to arrive (ticket counter) sums all of
// that volume's partials in a fixed order, so the scores do not depend on CTA scheduling and
// no second kernel launch is needed.
template <int BAR = 0>
__device__ __forceinline__ void k1_write_partials(const K1Params& prm, double (&part)[9]) {
    __shared__ double red[9 * 8];
    __shared__ int s_last;
    block_sum<9, BAR>(part, red);
    const int64_t b = blockIdx.x / prm.blocks_per_vol;
    if (threadIdx.x == 0) {
        double* dst = prm.partials + (int64_t)blockIdx.x * 9;
#pragma unroll
        for (int k = 0; k < 9; ++k) dst[k] = part[k];
        __threadfence();
        const unsigned int ticket = atomicAdd(prm.counters + b, 1u);
        s_last = ticket == (unsigned int)(prm.blocks_per_vol - 1);
    }
    block_sync<BAR>();
    if (!s_last) return;
    __threadfence();
    const double* src = prm.partials + b * prm.blocks_per_vol * 9;
    double acc[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] = 0.0;
    for (int64_t r = threadIdx.x; r < prm.blocks_per_vol; r += kThreads) {
#pragma unroll
        for (int k = 0; k < 9; ++k) acc[k] += __ldcg(src + r * 9 + k);
    }
    block_sync<BAR>();
    block_sum<9, BAR>(acc, red);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 9; ++k) prm.scores[(3 * b + k / 3) * prm.score_stride + k % 3] = acc[k];
    }
}

// =========================================================================== K1 stream kernel
// One generic kernel for any N and C: the (class, sample) rows of a voxel vector are walked
// class-outer / sample-inner, so the only state that lives across rows is the running class
// sum S, the entropy sum and the arg-max -- a handful of registers regardless of C.  Rows
// are fetched U at a time, one batch ahead of the arithmetic (register double buffer), with
// 128-bit read-once loads.  fp32 / bf16 inputs use packed f32x2 arithmetic (FFMA2 / FADD2 /
// FMUL2, sm_100): the kernel is bounded by HBM only if p*log(p) costs ~10 issue slots per
// element, which the packed degree-7 polynomial does (11) and the scalar one (20) does not.
//
// NaN-skip rule without a per-element select: for finite p >= +0 the polynomial log is finite,
// so p*log(p) is already the reference's contribution (0 for p == 0).  Elements the rule would
// treat differently (negative, -0, NaN, inf) are detected from the OR of the sign bits and the
// finiteness of the class sums; a thread that saw one recomputes its entropy sum with the
// exact per-element select (k1_entropy_exact) -- a cold path softmax outputs never take.

template <typename T, int VEC> struct Raw {
    static constexpr int BYTES = VEC * (int)sizeof(T);
    static constexpr int WORDS = BYTES >= 4 ? BYTES / 4 : 1;
    uint32_t w[WORDS];
};

// 128-bit read-once load that also marks the line evict-first in L2: the softmax stack is
// touched exactly once, while the maps K1 writes are re-read by K2b and should stay resident.
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
template <typename T, int VEC>
__device__ __forceinline__ void load_raw_stream(const T* p, Raw<T, VEC>& r, uint64_t pol) {
    static_assert(Raw<T, VEC>::BYTES == 16, "vector path only");
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]) : "l"(p), "l"(pol));
}

template <typename T, int VEC>
__device__ __forceinline__ void load_raw(const T* p, Raw<T, VEC>& r) {
    constexpr int BYTES = Raw<T, VEC>::BYTES;
    if constexpr (BYTES == 16) {
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]) : "l"(p));
    } else if constexpr (BYTES == 8) {
        asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];"
                     : "=r"(r.w[0]), "=r"(r.w[1]) : "l"(p));
    } else if constexpr (BYTES == 4) {
        asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r.w[0]) : "l"(p));
    } else {
        static_assert(BYTES == 2, "unsupported vector width");
        uint16_t h;
        asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(h) : "l"(p));
        r.w[0] = h;
    }
}

template <int VEC> __device__ __forceinline__ void unpack(const Raw<float, VEC>& r, float (&o)[VEC]) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) o[j] = __uint_as_float(r.w[j]);
}
template <int VEC> __device__ __forceinline__ void unpack(const Raw<double, VEC>& r, double (&o)[VEC]) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) o[j] = __hiloint2double((int)r.w[2 * j + 1], (int)r.w[2 * j]);
}
template <int VEC>
__device__ __forceinline__ void unpack(const Raw<__nv_bfloat16, VEC>& r, float (&o)[VEC]) {
    if constexpr (VEC == 1) {
        o[0] = __uint_as_float(r.w[0] << 16);
    } else {
#pragma unroll
        for (int j = 0; j < VEC / 2; ++j) {
            o[2 * j] = __uint_as_float(r.w[j] << 16);
            o[2 * j + 1] = __uint_as_float(r.w[j] & 0xffff0000u);
        }
    }
}
// OR of the sign bits of every element of a row (bit 31 set <=> some element has its sign set)
template <int VEC> __device__ __forceinline__ uint32_t sign_or(const Raw<float, VEC>& r) {
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < VEC; ++j) s |= r.w[j];
    return s;
}
// bf16: two elements per word, signs at bits 15 and 31 -- the words are OR-ed as they are (one LOP3 per two
// words instead of a shift and an OR per word) and bit 15 is folded onto bit 31 once per voxel (fold_sign)
template <int VEC> __device__ __forceinline__ uint32_t sign_or(const Raw<__nv_bfloat16, VEC>& r) {
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < Raw<__nv_bfloat16, VEC>::WORDS; ++j) s |= r.w[j];
    return s;
}
template <typename T> __device__ __forceinline__ uint32_t fold_sign(uint32_t bad) {
    if constexpr (sizeof(T) == 2) return bad | (bad << 16);
    else return bad;
}
template <int VEC> __device__ __forceinline__ uint32_t sign_or(const Raw<double, VEC>&) { return 0; }

__device__ __forceinline__ float2 splat2(float a) { return make_float2(a, a); }

// acc += p*log(p) for two fp32 elements, packed; same polynomial as fast_logf, so every
// finite p > 0 gets bit-identical terms on the packed and the scalar (exact) path.
__device__ __forceinline__ void plogp_pair(float& acc0, float& acc1, float p0, float p1) {
    const int i0 = __float_as_int(p0), i1 = __float_as_int(p1);
    const int e0 = (i0 - 0x3f2aaaab) & 0xff800000;
    const int e1 = (i1 - 0x3f2aaaab) & 0xff800000;
    const float2 f = __fadd2_rn(make_float2(__int_as_float(i0 - e0), __int_as_float(i1 - e1)),
                                splat2(-1.0f));
    float2 q = __ffma2_rn(splat2(0.13979104161262512f), f, splat2(-0.15397904813289642f));
    q = __ffma2_rn(q, f, splat2(0.14004801213741302f));
    q = __ffma2_rn(q, f, splat2(-0.1641434133052826f));
    q = __ffma2_rn(q, f, splat2(0.20010659098625183f));
    q = __ffma2_rn(q, f, splat2(-0.2500789761543274f));
    q = __ffma2_rn(q, f, splat2(0.3333320617675781f));
    q = __ffma2_rn(q, f, splat2(-0.49999934434890747f));
    const float2 r = __ffma2_rn(__fmul2_rn(q, f), f, f);
    const float2 lg = __ffma2_rn(make_float2((float)e0, (float)e1), splat2(8.262958317573066e-08f), r);
    const float2 a = __ffma2_rn(make_float2(p0, p1), lg, make_float2(acc0, acc1));
    acc0 = a.x; acc1 = a.y;
}

// bf16 inputs (tolerance 1e-3): log2 on the SFU (MUFU.LG2, abs err 2^-22), entropy kept in
// log2 units until the epilogue.  max(p, FLT_MIN) keeps 0*log(0) at 0 without a select.
__device__ __forceinline__ float lg2_ftz(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Branch-free first-max-wins arg-max update; NaN counts as maximal (np.argmax / torch.argmax).
template <typename A>
__device__ __forceinline__ void argmax_update_sel(A v, int c, A& best, int& idx) {
    const bool take = (v > best) | ((v != v) & (best == best));
    best = take ? v : best;
    idx = take ? c : idx;
}

// m = S / N, correctly rounded.  fp32: q0 = S*y, r = fma(-q0, N, S), q = fma(r, y, q0) with
// y = RN(1/N) -- the refinement CUDA's own IEEE division ends with -- packed two at a time.
// outside the range where the residual fma is exact (tiny non-zero, huge, inf, NaN): true division
__device__ __forceinline__ bool mean_needs_div(float s) {
    const float a = fabsf(s);
    return !(a < 1e30f) | ((a < 1e-30f) & (a != 0.f));
}
__device__ __noinline__ float mean_div_slow(float s, float n) { return s / n; }  // cold
template <int VEC>
__device__ __forceinline__ void class_mean(const float (&S)[VEC], float Nf, float inv_n, float (&m)[VEC]) {
    if constexpr (VEC == 1) {
        const float q0 = S[0] * inv_n;
        m[0] = fmaf(fmaf(-q0, Nf, S[0]), inv_n, q0);
        if (mean_needs_div(S[0])) m[0] = mean_div_slow(S[0], Nf);
    } else {
#pragma unroll
        for (int j = 0; j < VEC; j += 2) {
            const float2 s2 = make_float2(S[j], S[j + 1]);
            const float2 q0 = __fmul2_rn(s2, splat2(inv_n));
            const float2 r = __ffma2_rn(q0, splat2(-Nf), s2);
            const float2 q = __ffma2_rn(r, splat2(inv_n), q0);
            m[j] = q.x; m[j + 1] = q.y;
            if (mean_needs_div(S[j])) m[j] = mean_div_slow(S[j], Nf);
            if (mean_needs_div(S[j + 1])) m[j + 1] = mean_div_slow(S[j + 1], Nf);
        }
    }
}
// fp64: the same refinement (DMUL + 2 DFMA instead of a ~25-instruction DDIV); checked against true
// division for N up to 255 over 17 binades around the exponent guards (fractions on the host, and
// tests/test_gpu_parity.py::test_fp64_class_mean_is_true_division on the device)
__device__ __noinline__ double mean_div_slow(double s, double n) { return s / n; }  // cold
template <int VEC>
__device__ __forceinline__ void class_mean(const double (&S)[VEC], double Nf, float, double (&m)[VEC]) {
    const double y = 1.0 / Nf;   // loop-invariant: hoisted out of the voxel loop
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        const double q0 = S[j] * y;
        m[j] = fma(fma(-q0, Nf, S[j]), y, q0);
        const unsigned int e = ((unsigned int)__double2hiint(S[j]) >> 20) & 0x7ffu;   // biased exponent
        if ((e - 200u) >= 1600u && S[j] != 0.0) m[j] = mean_div_slow(S[j], Nf);        // tiny, huge, inf, NaN
    }
}

// PE[j] += m*log(m) with the NaN-skip select (exact for every m, packed where possible)
template <int VEC>
__device__ __forceinline__ void pe_terms(float (&PE)[VEC], const float (&m)[VEC]) {
    if constexpr (VEC == 1) {
        accum_term(PE[0], m[0]);
    } else {
#pragma unroll
        for (int j = 0; j < VEC; j += 2) {
            float t0 = PE[j], t1 = PE[j + 1];
            plogp_pair(t0, t1, m[j], m[j + 1]);
            PE[j] = (m[j] > 0.f) ? t0 : PE[j];
            PE[j + 1] = (m[j + 1] > 0.f) ? t1 : PE[j + 1];
        }
    }
}
template <int VEC>
__device__ __forceinline__ void pe_terms(float (&PE)[VEC], const double (&m)[VEC]) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) accum_term(PE[j], m[j]);
}
// fp64 ring kernel: the log table is already in shared memory (the global-memory lookup of accum_term was
// 9 % of the kernel's stall samples, ncu r02u); same arithmetic, bit-identical
template <int VEC>
__device__ __forceinline__ void pe_terms_tab(float (&PE)[VEC], const double (&m)[VEC], uint32_t tab) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        const double t = m[j] * fast_log_f64_tab(m[j], tab);
        PE[j] = (t == t) ? (float)((double)PE[j] + t) : PE[j];
    }
}
template <int VEC>
__device__ __forceinline__ void pe_terms_tab(float (&PE)[VEC], const float (&m)[VEC], uint32_t) { pe_terms<VEC>(PE, m); }

template <typename T> struct Math;
template <> struct Math<float> {
    static constexpr bool kFlagged = true;     // fast path + flag + exact recompute
    static constexpr float kScale = 1.0f;      // entropy accumulators are in nats
    template <int VEC>
    static __device__ __forceinline__ void rows(float (&e)[VEC], const float (&p)[VEC]) {
        if constexpr (VEC == 1) {
            e[0] = fmaf(p[0], fast_logf(p[0]), e[0]);
        } else {
#pragma unroll
            for (int j = 0; j < VEC; j += 2) plogp_pair(e[j], e[j + 1], p[j], p[j + 1]);
        }
    }
};
template <> struct Math<__nv_bfloat16> {
    static constexpr bool kFlagged = true;
    static constexpr float kScale = 0.693147180559945309f;  // accumulators are in bits
    template <int VEC>
    static __device__ __forceinline__ void rows(float (&e)[VEC], const float (&p)[VEC]) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) e[j] = fmaf(p[j], lg2_ftz(fmaxf(p[j], 1.17549435e-38f)), e[j]);
    }
};
template <> struct Math<double> {
    static constexpr bool kFlagged = false;    // exact per-element select, always
    static constexpr float kScale = 1.0f;
    template <int VEC>
    static __device__ __forceinline__ void rows(float (&e)[VEC], const double (&p)[VEC]) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) accum_term(e[j], p[j]);
    }
};

// Cold path: entropy sum of one voxel vector with the exact NaN-skip select per element.
template <typename T, int VEC>
__device__ __noinline__ void k1_entropy_exact(const T* base, int N, int C, int64_t sn, int64_t sc,
                                              float* E_out) {
    using A = typename In<T>::acc_t;
    float E[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) E[j] = 0.f;
    for (int c = 0; c < C; ++c) {
        float e[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) e[j] = 0.f;
        for (int n = 0; n < N; ++n) {
            Raw<T, VEC> r;
            load_raw<T, VEC>(base + c * sc + n * sn, r);
            A p[VEC];
            unpack(r, p);
#pragma unroll
            for (int j = 0; j < VEC; ++j) accum_term(e[j], p[j]);
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j) E[j] += e[j];
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) E_out[j] = E[j];
}

// Cold path of the fp64 ring kernel: the reference's own order -- one fp32 accumulator per sample,
// classes in index order, exact NaN-skip select, the samples' entropies added in index order.
template <typename T, int VEC>
__device__ __noinline__ void k1_entropy_exact_samples(const T* base, int N, int C, int64_t sn, int64_t sc,
                                                      float* E_out) {
    using A = typename In<T>::acc_t;
    float E[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) E[j] = 0.f;
    for (int n = 0; n < N; ++n) {
        float h[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) h[j] = 0.f;
        for (int c = 0; c < C; ++c) {
            Raw<T, VEC> r;
            load_raw<T, VEC>(base + c * sc + n * sn, r);
            A p[VEC];
            unpack(r, p);
#pragma unroll
            for (int j = 0; j < VEC; ++j) accum_term(h[j], p[j]);
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j) E[j] += h[j];
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) E_out[j] = E[j];
}

// SH mode of the ring kernel: element j of a thread is VEC-strided through the tile (voxel tid + 256 j), read
// from shared memory element by element into the same Raw words the vector path fills
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
template <typename T, int VEC> __device__ __forceinline__ void lds_strided(uint32_t a, Raw<T, VEC>& r) {
    constexpr uint32_t kStep = (uint32_t)kThreads * (uint32_t)sizeof(T);
    if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) r.w[j] = lds32(a + j * kStep);
    } else if constexpr (sizeof(T) == 8) {
#pragma unroll
        for (int j = 0; j < VEC; ++j)
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r.w[2 * j]), "=r"(r.w[2 * j + 1]) : "r"(a + j * kStep));
    } else {
#pragma unroll
        for (int j = 0; j < VEC; j += 2) {
            uint16_t lo, hi;
            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(lo) : "r"(a + j * kStep));
            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(hi) : "r"(a + (j + 1) * kStep));
            r.w[j / 2] = (uint32_t)lo | ((uint32_t)hi << 16);
        }
    }
}
// elements past the ragged end of a volume's last tile read as +0 (no term, no flag)
template <typename T, int VEC> __device__ __forceinline__ void mask_strided(Raw<T, VEC>& r, int tid, int nvalid) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        if (tid + kThreads * j < nvalid) continue;
        if constexpr (sizeof(T) == 4) r.w[j] = 0u;
        else if constexpr (sizeof(T) == 8) { r.w[2 * j] = 0u; r.w[2 * j + 1] = 0u; }
        else r.w[j / 2] &= (j & 1) ? 0x0000ffffu : 0xffff0000u;
    }
}

// S += p, packed two at a time for fp32
template <int VEC> __device__ __forceinline__ void add_rows(float (&S)[VEC], const float (&p)[VEC]) {
    if constexpr (VEC % 2 == 0) {
#pragma unroll
        for (int j = 0; j < VEC; j += 2) {
            const float2 r = __fadd2_rn(make_float2(S[j], S[j + 1]), make_float2(p[j], p[j + 1]));
            S[j] = r.x; S[j + 1] = r.y;
        }
    } else {
#pragma unroll
        for (int j = 0; j < VEC; ++j) S[j] += p[j];
    }
}
template <int VEC> __device__ __forceinline__ void add_rows(double (&S)[VEC], const double (&p)[VEC]) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) S[j] += p[j];
}

// FULL: N % U == 0, every batch holds exactly U rows (no per-row guards).
template <typename T, int VEC, int U, int MINB, bool FULL>
__global__ void __launch_bounds__(kThreads, MINB) k1_stream_kernel(const K1Params prm) {
    using A = typename In<T>::acc_t;
    using M = Math<T>;
    const int64_t b = blockIdx.x / prm.blocks_per_vol;
    const int64_t blk = blockIdx.x - b * prm.blocks_per_vol;
    const int N = (int)prm.N, C = (int)prm.C;
    const int64_t snb = prm.sn * (int64_t)sizeof(T), scb = prm.sc * (int64_t)sizeof(T);  // bytes
    const A Nf = (A)prm.N;
    const uint64_t policy = l2_evict_first_policy();
    double psum[6];
    int pcnt[3];
#pragma unroll
    for (int k = 0; k < 6; ++k) psum[k] = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) pcnt[k] = 0;

    for (int it = 0; it < prm.iter; ++it) {
        const int64_t v0 = ((blk * prm.iter + it) * kThreads + threadIdx.x) * VEC;
        if (v0 >= prm.V) break;
        const T* base = reinterpret_cast<const T*>(prm.probs) + b * prm.sb + v0;
        A S[VEC], best[VEC];
        float e[VEC], E[VEC], PE[VEC];
        int idx[VEC];
        float Sacc[VEC];  // sum of the class sums: non-finite iff some input was +-inf / NaN
        uint32_t bad = 0;
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            S[j] = (A)0; e[j] = 0.f; E[j] = 0.f; PE[j] = 0.f; idx[j] = 0; best[j] = (A)0; Sacc[j] = 0.f;
        }

        // Rows are fetched in batches of <= U samples of ONE class (a batch never straddles a
        // class boundary, so the class epilogue sits at two code sites only); the load cursor
        // runs one batch ahead of the arithmetic across class boundaries.
        const char* lrow = reinterpret_cast<const char*>(base);  // load cursor: next batch
        int ln = 0, lc = 0;
        auto issue = [&](Raw<T, VEC> (&buf)[U]) {
            if (lc < C) {
                const char* r = lrow;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (FULL || ln + u < N) {
                        if constexpr (Raw<T, VEC>::BYTES == 16)
                            load_raw_stream<T, VEC>(reinterpret_cast<const T*>(r), buf[u], policy);
                        else
                            load_raw<T, VEC>(reinterpret_cast<const T*>(r), buf[u]);
                    }
                    r += snb;
                }
                lrow = r;
                ln += U;
                if (ln >= N) { lrow += scb - (int64_t)ln * snb; ln = 0; ++lc; }
            }
        };
        int n = 0, c = 0;       // consume cursor
        auto consume = [&](const Raw<T, VEC> (&buf)[U]) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (FULL || n + u < N) {
                    A p[VEC];
                    unpack(buf[u], p);
                    if (M::kFlagged) bad |= sign_or(buf[u]);
                    add_rows<VEC>(S, p);
                    M::rows(e, p);
                }
            }
            n += U;
            if (n >= N) {  // class c complete: mean, arg-max, PE term (uniform branch)
                A m[VEC];
                class_mean<VEC>(S, Nf, prm.inv_n, m);
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    if (c == 0) best[j] = m[j]; else argmax_update_sel(m[j], c, best[j], idx[j]);
                    if (M::kFlagged) Sacc[j] += (float)S[j];  // stays finite iff every class sum is
                    S[j] = (A)0;
                    E[j] += e[j];
                    e[j] = 0.f;
                }
                pe_terms<VEC>(PE, m);
                n = 0; ++c;
            }
        };
        Raw<T, VEC> bufA[U], bufB[U];
        issue(bufA);
        while (true) {
            issue(bufB);
            consume(bufA);
            if (c >= C) break;
            issue(bufA);
            consume(bufB);
            if (c >= C) break;
        }
        if (M::kFlagged) {
#pragma unroll
            for (int j = 0; j < VEC; ++j)
                bad |= ((__float_as_uint(Sacc[j]) & 0x7f800000u) == 0x7f800000u) ? 0x80000000u : 0u;
        }
        bad = fold_sign<T>(bad);
        if (M::kFlagged && (bad & 0x80000000u)) {
            k1_entropy_exact<T, VEC>(base, N, C, prm.sn, prm.sc, E);
#pragma unroll
            for (int j = 0; j < VEC; ++j) E[j] = E[j] / M::kScale;  // back to accumulator units
        }
        float pe[VEC], ee[VEC], mi[VEC];
        const float nf = (float)prm.N;
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            pe[j] = -PE[j];
            ee[j] = -(M::kScale == 1.0f ? E[j] : E[j] * M::kScale) / nf;
            mi[j] = pe[j] - ee[j];
        }
        const int64_t o = b * prm.so + v0;
        if (prm.pe) store_f32<VEC>(prm.pe + o, pe);
        if (prm.ee) store_f32<VEC>(prm.ee + o, ee);
        if (prm.mi) store_f32<VEC>(prm.mi + o, mi);
        if (prm.amax) store_u8<VEC>(prm.amax + b * prm.V + v0, idx);
        if (prm.partials) {
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                const float m3[3] = {pe[j], ee[j], mi[j]};
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    psum[2 * k] += (double)m3[k];
                    if (prm.has_thr && m3[k] >= prm.thr_f[k]) { psum[2 * k + 1] += (double)m3[k]; ++pcnt[k]; }
                }
            }
        }
    }
    if (prm.partials) {
        double part[9];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            part[3 * k] = psum[2 * k]; part[3 * k + 1] = psum[2 * k + 1]; part[3 * k + 2] = (double)pcnt[k];
        }
        k1_write_partials(prm, part);
    }
}

// =========================================================================== K1 bulk-copy kernel
// Same arithmetic as k1_stream_kernel, different data movement: one producer thread streams
// the (tile, class, sample) rows of the CTA -- 4 KB each, the 16-byte vectors of its 256
// consumer threads side by side -- into a shared-memory ring with cp.async.bulk (the TMA
// engine, L2 evict-first), completion signalled on an mbarrier per stage; consumer warps wait on
// "full", read their 16 bytes, hand the slot back on "empty" (one arrive per warp) and do the
// arithmetic.  The loads in flight (kTmaStages x 4 KB per CTA) no longer depend on registers
// or on how far the arithmetic has got, and run ahead across tiles of the CTA.

// RS rows (samples of one class) share a stage and one mbarrier round trip; N % RS == 0.
// NS > 0 (fp64 stacks): the number of samples is the compile-time NS and every sample keeps its own
// fp32 entropy accumulator H[n], so each H_n still adds its classes in index order and EE is their
// sequential fp32 sum -- the reference's order (test_3D.py:499-507) -- while the rows stream
// class-outer through the ring like the fp32 path.  Bit-identical to the sample-outer k1_smem_kernel.
//
// SH (stacks whose rows are not 16-byte aligned: a voxel count that is not a multiple of the vector, odd
// strides, an offset base): every (class, sample) row of a tile starts at its own 16-byte phase, so no
// thread can read its four voxels of every row as one aligned vector.  The producer copies the aligned
// 16-byte granules that cover the row's 4 KB (a slot is 16 bytes longer), and a thread owns the voxels
// tid, tid + 256, tid + 512, ... of the tile instead of VEC consecutive ones: its element of row r sits at
// slot + phase_r + (tid + 256 j) * sizeof(T) -- element loads, lane-consecutive and conflict-free -- and
// the maps are written with lane-consecutive element stores.  Same arithmetic per voxel, so maps and
// arg-max are bit-identical to the other kernels; the fp64 score sums group the voxels differently.
// The granules before the first and after the last element of the stack are read as well: the host
// checks that they lie inside the allocation (dispatch_k1, allocation_range).
// RAG (sample counts no stage size divides: N = 7, 11, 13, 14, 17, ...): the last stage of a class holds the
// N % RS rows that are left -- `rows` below is uniform, and a compile-time RS without it -- instead of
// single-row stages with an mbarrier round trip per row.
template <typename T, int VEC, int MINB, int RS, int kTmaStages, int NS = 0, bool SH = false, bool RAG = false>
__global__ void __launch_bounds__(kThreads + 32, MINB) k1_tma_kernel(const K1Params prm) {
    using A = typename In<T>::acc_t;
    using M = Math<T>;
    static_assert(VEC * sizeof(T) == 16, "vector path only");
    static_assert(!RAG || NS == 0, "ragged stages: fp32 / bf16 stacks");
    constexpr int kRowBytes = kThreads * 16;
    constexpr int kRowPitch = kRowBytes + (SH ? 16 : 0);
    constexpr int kStageBytes = RS * kRowPitch;
    extern __shared__ __align__(128) unsigned char ring[];            // [kTmaStages][RS][kRowBytes]
    __shared__ __align__(8) uint64_t full_bar[kTmaStages], empty_bar[kTmaStages];
    __shared__ __align__(16) double2 s_log_tab[NS > 0 ? 513 : 1];     // fp64 stacks: the log table, 8 KB
    const int tid = threadIdx.x;
    const int64_t b = blockIdx.x / prm.blocks_per_vol;
    const int64_t blk = blockIdx.x - b * prm.blocks_per_vol;
    const int N = (int)prm.N, C = (int)prm.C;
    const int64_t snb = prm.sn * (int64_t)sizeof(T), scb = prm.sc * (int64_t)sizeof(T);
    if constexpr (NS > 0) {
        for (int i = tid; i < 513; i += kThreads + 32) s_log_tab[i] = kLog64Tab[i];
    }
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kTmaStages; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, kThreads / 32 * kArriveLanes); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const char* vol = reinterpret_cast<const char*>(reinterpret_cast<const T*>(prm.probs) + b * prm.sb);
    const int64_t tile_vox = (int64_t)kThreads * VEC;

    if (tid >= kThreads) {   // ---------------- producer warp: one elected lane issues every copy
        if (tid == kThreads) {
            const uint64_t policy = l2_evict_first_policy();
            int stage = 0;
            uint32_t phase = 0;
            for (int it = 0; it < prm.iter; ++it) {
                const int64_t t0 = (blk * prm.iter + it) * tile_vox;
                if (t0 >= prm.V) break;
                const uint32_t bytes = (uint32_t)(min(tile_vox, prm.V - t0) * (int64_t)sizeof(T));
                const char* row_c = vol + t0 * (int64_t)sizeof(T);
                for (int c = 0; c < C; ++c, row_c += scb) {
                    const char* row = row_c;
                    for (int n = 0; n < N; n += RS) {
                        mbar_wait(empty_bar + stage, phase ^ 1u);
                        const int rows = RAG ? min(RS, N - n) : RS;
                        if constexpr (SH) {
                            // the 16-byte granules covering [row, row + bytes) of every row of the stage
                            uint32_t total = 0;
                            const char* r2 = row;
#pragma unroll
                            for (int u = 0; u < RS; ++u, r2 += snb)
                                if (u < rows) total += (((uint32_t)reinterpret_cast<uintptr_t>(r2) & 15u) + bytes + 15u) & ~15u;
                            mbar_expect_tx(full_bar + stage, total);
#pragma unroll
                            for (int u = 0; u < RS; ++u, row += snb) {
                                if (u >= rows) continue;
                                const uint32_t a = (uint32_t)reinterpret_cast<uintptr_t>(row) & 15u;
                                bulk_g2s(ring + stage * kStageBytes + u * kRowPitch, row - a, (a + bytes + 15u) & ~15u,
                                         full_bar + stage, policy);
                            }
                        } else {
                            mbar_expect_tx(full_bar + stage, bytes * rows);
#pragma unroll
                            for (int u = 0; u < RS; ++u, row += snb)
                                if (u < rows) bulk_g2s(ring + stage * kStageBytes + u * kRowBytes, row, bytes, full_bar + stage, policy);
                        }
                        if (++stage == kTmaStages) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
        return;   // consumers synchronise among themselves with named barrier 1
    }

    // ---------------- consumers
    const A Nf = (A)prm.N;
    const int lane = tid & 31;
    // Shared-memory addresses and per-tile flags live in registers the compiler cannot re-derive
    // (the empty asm makes them opaque): at the 72-register cap it otherwise rebuilds them from
    // %tid / %cluster_ctaid and a 64-bit compare in EVERY stage -- ~20 of 420 instructions, and K1 is
    // issue-bound.
    uint32_t ring_tid = smem_u32(ring) + (uint32_t)tid * (SH ? (uint32_t)sizeof(T) : 16u);
    // SH: 16-byte phase of the rows, from the low address bits (tiles are whole multiples of 4 KB apart)
    const uint32_t ph_vol = (uint32_t)reinterpret_cast<uintptr_t>(vol), ph_sn = (uint32_t)snb, ph_sc = (uint32_t)scb;
    uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
    int is_lane0 = arrives(lane);
    uint32_t log_tab = smem_u32(s_log_tab);
    asm volatile("" : "+r"(ring_tid), "+r"(full0), "+r"(empty0), "+r"(is_lane0), "+r"(log_tab));
    double psum[6];
    int pcnt[3];
#pragma unroll
    for (int k = 0; k < 6; ++k) psum[k] = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) pcnt[k] = 0;
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < prm.iter; ++it) {
        const int64_t t0 = (blk * prm.iter + it) * tile_vox;
        if (t0 >= prm.V) break;                     // uniform: the producer stops at the same tile
        const int64_t v0 = SH ? t0 + tid : t0 + (int64_t)tid * VEC;
        int active = v0 < prm.V;                    // only the last tile of a volume is ragged
        asm volatile("" : "+r"(active));
        // SH: voxels of the tile (a thread's element j is voxel tid + 256 j: past a ragged end it is zeroed)
        const int nvalid = SH ? (int)min(tile_vox, prm.V - t0) : 0;
        const bool ragged = SH && nvalid < (int)tile_vox;
        A S[VEC], best[VEC];
        float e[VEC], E[VEC], PE[VEC], Sacc[VEC];
        int idx[VEC];
        uint32_t bad = 0;
        unsigned int hi_max = 0;                    // fp64 stacks: largest high word seen (negative / inf / NaN inputs)
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            S[j] = (A)0; e[j] = 0.f; E[j] = 0.f; PE[j] = 0.f; idx[j] = 0; best[j] = (A)0; Sacc[j] = 0.f;
        }
        float H[NS > 0 ? NS : 1][VEC];
#pragma unroll
        for (int n = 0; n < (NS > 0 ? NS : 1); ++n) {
#pragma unroll
            for (int j = 0; j < VEC; ++j) H[n][j] = 0.f;
        }
        for (int c = 0; c < C; ++c) {
            auto stage_body = [&](const int n) {
                mbar_wait_addr(full0 + stage * 8, phase);
                // rows of a stage are consumed SB at a time (registers), the stage is one mbarrier round trip
                constexpr int SB = RS > 5 ? RS / 2 : RS;
                static_assert(RS % SB == 0, "sub-batches");
                const int rows = RAG ? min(RS, N - n) : RS;
#pragma unroll
                for (int h = 0; h < RS; h += SB) {
                    Raw<T, VEC> raw[SB];
#pragma unroll
                    for (int u = 0; u < SB; ++u) {
                        if (RAG && h + u >= rows) continue;
                        if constexpr (SH) {
                            const uint32_t a = (ph_vol + (uint32_t)c * ph_sc + (uint32_t)(n + h + u) * ph_sn) & 15u;
                            lds_strided<T, VEC>(ring_tid + stage * kStageBytes + (h + u) * kRowPitch + a, raw[u]);
                            if (ragged) mask_strided<T, VEC>(raw[u], tid, nvalid);
                        } else {
                            const uint4 q = lds128(ring_tid + stage * kStageBytes + (h + u) * kRowBytes);
                            raw[u].w[0] = q.x; raw[u].w[1] = q.y; raw[u].w[2] = q.z; raw[u].w[3] = q.w;
                        }
                    }
                    if (active) {
#pragma unroll
                        for (int u = 0; u < SB; ++u) {
                            if (RAG && h + u >= rows) continue;
                            A p[VEC];
                            unpack(raw[u], p);
                            if (M::kFlagged) bad |= sign_or(raw[u]);
                            add_rows<VEC>(S, p);
                            if constexpr (NS > 0) {
#pragma unroll
                                for (int j = 0; j < VEC; ++j) accum_term_fast(H[n + h + u][j], p[j], hi_max, log_tab);
                            } else {
                                M::rows(e, p);
                            }
                        }
                    }
                }
                // The slot may only be handed back once the LDS above have RETURNED (an arrive
                // issued right after them races with the next bulk copy -- measured).  The
                // empty asm pins the arrive behind arithmetic that consumed every row of the
                // stage, and that arithmetic cannot issue before the loads complete.
                asm volatile("" ::"f"(NS > 0 ? H[n + RS - 1][0] : e[0]), "r"(bad) : "memory");
                __syncwarp();
                if (is_lane0) mbar_arrive_addr(empty0 + stage * 8);
                if (++stage == kTmaStages) { stage = 0; phase ^= 1u; }
            };
            if constexpr (NS > 0) {
#pragma unroll
                for (int n = 0; n < NS; n += RS) stage_body(n);     // unrolled: H[] is indexed statically
            } else {
                for (int n = 0; n < N; n += RS) stage_body(n);
            }
            if (active) {   // class c complete: mean, arg-max, PE term
                A m[VEC];
                class_mean<VEC>(S, Nf, prm.inv_n, m);
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    if (c == 0) best[j] = m[j]; else argmax_update_sel(m[j], c, best[j], idx[j]);
                    if (M::kFlagged) Sacc[j] += (float)S[j];
                    S[j] = (A)0;
                    E[j] += e[j];
                    e[j] = 0.f;
                }
                if constexpr (NS > 0) pe_terms_tab<VEC>(PE, m, log_tab); else pe_terms<VEC>(PE, m);
            }
        }
        if (!active) continue;
        if constexpr (NS > 0) {   // EE numerator: the samples' entropies added in index order (fp32)
#pragma unroll
            for (int n = 0; n < NS; ++n) {
#pragma unroll
                for (int j = 0; j < VEC; ++j) E[j] += H[n][j];
            }
            if (hi_max >= 0x7ff00000u) {   // a negative, -0, inf or NaN input: the exact select, sample by sample
                if constexpr (SH) {
#pragma unroll
                    for (int j = 0; j < VEC; ++j)
                        if (tid + kThreads * j < nvalid)
                            k1_entropy_exact_samples<T, 1>(reinterpret_cast<const T*>(vol) + v0 + kThreads * j, N, C,
                                                           prm.sn, prm.sc, &E[j]);
                } else {
                    k1_entropy_exact_samples<T, VEC>(reinterpret_cast<const T*>(vol) + v0, N, C, prm.sn, prm.sc, E);
                }
            }
        }
        if (M::kFlagged) {
#pragma unroll
            for (int j = 0; j < VEC; ++j)
                bad |= ((__float_as_uint(Sacc[j]) & 0x7f800000u) == 0x7f800000u) ? 0x80000000u : 0u;
            bad = fold_sign<T>(bad);
            if (bad & 0x80000000u) {
                if constexpr (SH) {
#pragma unroll
                    for (int j = 0; j < VEC; ++j)
                        if (tid + kThreads * j < nvalid)
                            k1_entropy_exact<T, 1>(reinterpret_cast<const T*>(vol) + v0 + kThreads * j, N, C, prm.sn,
                                                   prm.sc, &E[j]);
                } else {
                    k1_entropy_exact<T, VEC>(reinterpret_cast<const T*>(vol) + v0, N, C, prm.sn, prm.sc, E);
                }
#pragma unroll
                for (int j = 0; j < VEC; ++j) E[j] = E[j] / M::kScale;
            }
        }
        float pe[VEC], ee[VEC], mi[VEC];
        const float nf = (float)prm.N;
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            pe[j] = -PE[j];
            ee[j] = -(M::kScale == 1.0f ? E[j] : E[j] * M::kScale) / nf;
            mi[j] = pe[j] - ee[j];
        }
        const int64_t o = b * prm.so + v0;
        if constexpr (SH) {   // element stores, lane-consecutive: no alignment asked of the maps
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                if (ragged && tid + kThreads * j >= nvalid) continue;
                if (prm.pe) prm.pe[o + kThreads * j] = pe[j];
                if (prm.ee) prm.ee[o + kThreads * j] = ee[j];
                if (prm.mi) prm.mi[o + kThreads * j] = mi[j];
                if (prm.amax) prm.amax[b * prm.V + v0 + kThreads * j] = (uint8_t)idx[j];
            }
        } else {
            if (prm.pe) store_f32<VEC>(prm.pe + o, pe);
            if (prm.ee) store_f32<VEC>(prm.ee + o, ee);
            if (prm.mi) store_f32<VEC>(prm.mi + o, mi);
            if (prm.amax) store_u8<VEC>(prm.amax + b * prm.V + v0, idx);
        }
        if (prm.partials) {
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                if (SH && ragged && tid + kThreads * j >= nvalid) continue;
                const float m3[3] = {pe[j], ee[j], mi[j]};
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    psum[2 * k] += (double)m3[k];
                    if (prm.has_thr && m3[k] >= prm.thr_f[k]) { psum[2 * k + 1] += (double)m3[k]; ++pcnt[k]; }
                }
            }
        }
    }
    if (prm.partials) {
        double part[9];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            part[3 * k] = psum[2 * k]; part[3 * k + 1] = psum[2 * k + 1]; part[3 * k + 2] = (double)pcnt[k];
        }
        k1_write_partials<1>(prm, part);
    }
    // every consumer has passed its last full-barrier wait, so the producer lane has issued its last copy and
    // every copy has landed: the barriers are idle and can be invalidated (see mbar_inval)
    asm volatile("bar.sync 1, %0;" ::"n"(kThreads) : "memory");
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kTmaStages; ++s) { mbar_inval(full_bar + s); mbar_inval(empty_bar + s); }
    }
}

// ---- arg-max only (no entropy maps, no scores): data_carrier_3D.py:253-285 (mean_seg / pred_seg at save
// time), test_2D.py:119-127 (per-sample arg-max incl. the zero channel), the Dice / GED inputs of f2.
// A segmented sweep over the stack's (N x C) rows of one 16-byte voxel vector, U rows per batch with the
// next batch already in flight (register double buffer, read-once evict-first loads):
//   MEAN = false  per-sample arg-max: rows sample-outer / class-inner, a segment is one sample (C rows),
//                 state = running best + index; the u8 map of sample n leaves when its segment closes;
//   MEAN = true   arg-max of the mean: rows class-outer / sample-inner, a segment is one class (N rows),
//                 state = class sum (sequential over n, as every other kernel) -> mean by the correctly
//                 rounded class_mean -> running best over the classes.
// First maximum wins, NaN counts as maximal (np.argmax / torch.argmax); results are those of k1_smem_kernel
// (which this replaces for aligned stacks: 0.26-0.61 of the HBM peak, a shared-memory round trip per element).
template <typename T, int VEC, int U, bool MEAN>
__global__ void __launch_bounds__(kThreads) k1_argmax_kernel(const K1Params prm) {
    using A = typename In<T>::acc_t;
    static_assert(VEC * sizeof(T) == 16 || VEC == 1, "16-byte vectors, or single elements for unaligned stacks");
    const int64_t b = blockIdx.x / prm.blocks_per_vol;
    const int64_t blk = blockIdx.x - b * prm.blocks_per_vol;
    const int64_t v0 = (blk * kThreads + threadIdx.x) * VEC;
    if (v0 >= prm.V) return;
    const int L = MEAN ? (int)prm.N : (int)prm.C;            // rows per segment
    const int nseg = MEAN ? (int)prm.C : (int)prm.N;
    const int64_t s_in = MEAN ? prm.sn : prm.sc;             // row to row inside a segment
    const int64_t s_out = (MEAN ? prm.sc : prm.sn) - (int64_t)L * s_in;   // last row of a segment to the next segment
    const int R = L * nseg;
    const uint64_t pol = l2_evict_first_policy();
    const T* lp = reinterpret_cast<const T*>(prm.probs) + b * prm.sb + v0;   // load cursor
    int lpos = 0;
    Raw<T, VEC> buf[2][U];
    auto issue = [&](Raw<T, VEC> (&dst)[U], int r0) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (r0 + u < R) {
                if constexpr (VEC * sizeof(T) == 16) load_raw_stream<T, VEC>(lp, dst[u], pol);
                else load_raw<T, VEC>(lp, dst[u]);
                lp += s_in;
                if (++lpos == L) { lpos = 0; lp += s_out; }
            }
        }
    };
    A acc[VEC], best[VEC];
    int idx[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) { acc[j] = (A)0; best[j] = (A)0; idx[j] = 0; }
    const A Nf = (A)prm.N;
    int pos = 0, seg = 0;
    auto consume = [&](const Raw<T, VEC> (&src)[U], int r0) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (r0 + u >= R) break;
            A p[VEC];
            unpack(src[u], p);
            if constexpr (MEAN) {
                if (pos == 0) {
#pragma unroll
                    for (int j = 0; j < VEC; ++j) acc[j] = p[j];
                } else {
                    add_rows<VEC>(acc, p);
                }
            } else {
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    if (pos == 0) { best[j] = p[j]; idx[j] = 0; }
                    else argmax_update_sel(p[j], pos, best[j], idx[j]);
                }
            }
            if (++pos == L) {
                pos = 0;
                if constexpr (MEAN) {
                    A m[VEC];
                    class_mean<VEC>(acc, Nf, prm.inv_n, m);
#pragma unroll
                    for (int j = 0; j < VEC; ++j) {
                        if (seg == 0) { best[j] = m[j]; idx[j] = 0; }
                        else argmax_update_sel(m[j], seg, best[j], idx[j]);
                    }
                } else {
                    store_u8<VEC>(prm.samax + (b * prm.N + seg) * prm.V + v0, idx);
                }
                ++seg;
            }
        }
    };
    issue(buf[0], 0);
    for (int r0 = 0; r0 < R; r0 += 2 * U) {
        issue(buf[1], r0 + U);
        consume(buf[0], r0);
        issue(buf[0], r0 + 2 * U);
        consume(buf[1], r0 + U);
    }
    if constexpr (MEAN) store_u8<VEC>(prm.amax + b * prm.V + v0, idx);
}

template <typename T, int VEC, bool MEAN>
static int launch_argmax(K1Params prm, int64_t B, cudaStream_t st) {
    prm.blocks_per_vol = ceil_div(ceil_div(prm.V, VEC), kThreads);
    const int64_t grid = prm.blocks_per_vol * B;
    if (grid > 0x7fffffffLL) return set_error(VALUES_ERR_UNSUPPORTED, "grid too large");
    k1_argmax_kernel<T, VEC, (VEC == 1 ? 8 : 4), MEAN><<<(unsigned)grid, kThreads, 0, st>>>(prm);
    return check_launch(MEAN ? "k1_argmax_kernel<mean>" : "k1_argmax_kernel<sample>");
}

// ---- class sums in shared memory (any C); thread-private columns, conflict-free
template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads) k1_smem_kernel(const K1Params prm) {
    using A = typename In<T>::acc_t;
    struct alignas(sizeof(A) * VEC >= 16 ? 16 : sizeof(A) * VEC) Pack { A v[VEC]; };
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Pack* S = reinterpret_cast<Pack*>(smem_raw);
    const int C = (int)prm.C;
    const int64_t b = blockIdx.x / prm.blocks_per_vol;
    const int64_t blk = blockIdx.x - b * prm.blocks_per_vol;
    const int64_t v0 = (blk * kThreads + threadIdx.x) * VEC;
    double part[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) part[k] = 0.0;
    if (v0 < prm.V) {
        const T* base = reinterpret_cast<const T*>(prm.probs) + b * prm.sb + v0;
        float E[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) E[j] = 0.f;
        for (int64_t n = 0; n < prm.N; ++n) {
            float H[VEC];
            A best[VEC];
            int idx[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j) { H[j] = 0.f; idx[j] = 0; }
            const T* pn = base + n * prm.sn;
#pragma unroll 4
            for (int c = 0; c < C; ++c) {
                A p[VEC];
                load_elems<T, VEC>(pn + c * prm.sc, p);
                Pack s;
                if (n == 0) {
#pragma unroll
                    for (int j = 0; j < VEC; ++j) s.v[j] = p[j];
                } else {
                    s = S[c * kThreads + threadIdx.x];
#pragma unroll
                    for (int j = 0; j < VEC; ++j) s.v[j] += p[j];
                }
                S[c * kThreads + threadIdx.x] = s;
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    if (prm.need_ent) accum_term(H[j], p[j]);
                    if (prm.samax) {
                        if (c == 0) best[j] = p[j]; else argmax_update(p[j], c, best[j], idx[j]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < VEC; ++j) E[j] += H[j];
            if (prm.samax) store_u8<VEC>(prm.samax + (b * prm.N + n) * prm.V + v0, idx);
        }
        k1_epilogue<A, VEC, 0>(prm, b, v0, C, [&](int c, int j) -> A {
            return S[c * kThreads + threadIdx.x].v[j];
        }, E, part);
    }
    if (prm.partials) k1_write_partials(prm, part);
}

// ---- 1 - max_c p  (test_3D.py:521-525)
template <typename T>
__global__ void __launch_bounds__(kThreads) msr_kernel(const T* __restrict__ probs, int64_t C,
                                                       int64_t V, int64_t sb, int64_t sc,
                                                       T* __restrict__ out, int64_t blocks_per_vol) {
    using A = typename In<T>::acc_t;
    const int64_t b = blockIdx.x / blocks_per_vol;
    const int64_t v = (blockIdx.x - b * blocks_per_vol) * kThreads + threadIdx.x;
    if (v >= V) return;
    const T* p = probs + b * sb + v;
    A m = In<T>::load_one(p);
    for (int64_t c = 1; c < C; ++c) {
        const A x = In<T>::load_one(p + c * sc);
        m = (x > m || x != x) ? x : m;  // torch.max propagates NaN
    }
    const A r = (A)1 - m;
    if constexpr (sizeof(T) == 2) out[b * V + v] = __float2bfloat16_rn(r);
    else out[b * V + v] = r;
}

// the same on 16-byte voxel vectors, four class rows in flight per thread (the element kernel above: 0.34-0.59
// of the HBM peak on 256^3 / 2048^2 stacks)
template <typename T>
__global__ void __launch_bounds__(kThreads) msr_vec_kernel(const T* __restrict__ probs, int64_t C, int64_t V,
                                                           int64_t sb, int64_t sc, T* __restrict__ out,
                                                           int64_t blocks_per_vol) {
    using A = typename In<T>::acc_t;
    constexpr int VEC = In<T>::VEC;
    const int64_t b = blockIdx.x / blocks_per_vol;
    const int64_t v = ((blockIdx.x - b * blocks_per_vol) * kThreads + threadIdx.x) * VEC;
    if (v >= V) return;
    const T* p = probs + b * sb + v;
    const uint64_t pol = l2_evict_first_policy();
    A m[VEC];
    {
        Raw<T, VEC> r;
        load_raw_stream<T, VEC>(p, r, pol);
        unpack(r, m);
    }
    for (int64_t c = 1; c < C; c += 4) {
        Raw<T, VEC> r[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (c + u < C) load_raw_stream<T, VEC>(p + (c + u) * sc, r[u], pol);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (c + u >= C) break;
            A x[VEC];
            unpack(r[u], x);
#pragma unroll
            for (int j = 0; j < VEC; ++j) m[j] = (x[j] > m[j] || x[j] != x[j]) ? x[j] : m[j];   // torch.max propagates NaN
        }
    }
    T* o = out + b * V + v;
    if constexpr (sizeof(T) == 2) {
        uint32_t w[VEC / 2];
#pragma unroll
        for (int j = 0; j < VEC; j += 2) {
            const __nv_bfloat162 h = __floats2bfloat162_rn((A)1 - m[j], (A)1 - m[j + 1]);
            memcpy(&w[j / 2], &h, 4);
        }
        *reinterpret_cast<uint4*>(o) = make_uint4(w[0], w[1], w[2], w[3]);
    } else if constexpr (sizeof(T) == 4) {
        *reinterpret_cast<float4*>(o) = make_float4(1.f - m[0], 1.f - m[1], 1.f - m[2], 1.f - m[3]);
    } else {
        *reinterpret_cast<double2*>(o) = make_double2(1.0 - m[0], 1.0 - m[1]);
    }
}

template <typename T>
static int launch_msr(const void* probs, int64_t B, int64_t C, int64_t V, int64_t sb, int64_t sc, void* out,
                      cudaStream_t st) {
    constexpr int VEC = In<T>::VEC;
    const bool vec = V % VEC == 0 && sb % VEC == 0 && sc % VEC == 0 &&
                     reinterpret_cast<uintptr_t>(probs) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0;
    const int64_t bpv = ceil_div(vec ? V / VEC : V, kThreads);
    if (bpv * B > 0x7fffffffLL) return set_error(VALUES_ERR_UNSUPPORTED, "grid too large");
    const unsigned grid = (unsigned)(bpv * B);
    if (vec) msr_vec_kernel<T><<<grid, kThreads, 0, st>>>((const T*)probs, C, V, sb, sc, (T*)out, bpv);
    else msr_kernel<T><<<grid, kThreads, 0, st>>>((const T*)probs, C, V, sb, sc, (T*)out, bpv);
    return check_launch("msr_kernel");
}

template <typename T, int VEC>
static int launch_smem(const K1Params& prm, int64_t grid, cudaStream_t st) {
    using A = typename In<T>::acc_t;
    const size_t smem = (size_t)prm.C * kThreads * VEC * sizeof(A);
    if (smem > 227 * 1024) return 1;  // caller retries with a narrower vector
    auto kern = k1_smem_kernel<T, VEC>;
    if (smem > 48 * 1024) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
            cudaSuccess)
            return set_error(VALUES_ERR_CUDA, "cudaFuncSetAttribute(smem=%zu) failed", smem);
    }
    kern<<<(unsigned)grid, kThreads, smem, st>>>(prm);
    return check_launch("k1_smem_kernel");
}

// tiles per CTA of the stream kernel: amortise the block reduction over several tiles once
// the grid is many waves deep (148 SMs x MINB CTAs), keep one tile per CTA for small jobs.
// Stacks with few rows per voxel (cfg1/2: N*C = 10) do little work per tile, so the fixed cost per
// CTA (setup, score reduction, ticket) wants more tiles: measured on cfg2, 4 -> 16 tiles per CTA
// lifts K1 from 0.67 to 0.72 of the HBM peak, while cfg5 (N*C = 64) and cfg4 (200) peak at 4.
static int choose_iter(int64_t total_tiles, int minb, int64_t rows, bool f64 = false) {
    const int64_t resident = 148LL * minb;
    // fp64 tiles hold half the voxels and every CTA copies the 8 KB log table first: 16 tiles per CTA
    // measured +3 % on 256^3 N=8 C=2 (0.640 -> 0.663) and 128^3 N=16 C=4 (0.714 -> 0.733)
    int want = (f64 || rows <= 12) ? 16 : (rows <= 24 ? 8 : 4);
    while (want > 4 && total_tiles < 8 * want * resident) want >>= 1;
    if (want > 4) return want;
    if (total_tiles >= 32 * resident) return 4;
    if (total_tiles >= 12 * resident) return 2;
    return 1;
}

template <typename T, int VEC, int U, int MINB, bool FULL>
static int launch_stream(K1Params& prm, int64_t B, cudaStream_t st) {
    const int64_t tiles = ceil_div(ceil_div(prm.V, VEC), kThreads);
    prm.iter = prm.tiles_per_cta > 0 ? prm.tiles_per_cta : choose_iter(tiles * B, MINB, prm.N * prm.C);
    prm.blocks_per_vol = ceil_div(tiles, prm.iter);
    const int64_t grid = prm.blocks_per_vol * B;
    if (grid > 0x7fffffffLL) return set_error(VALUES_ERR_UNSUPPORTED, "grid too large");
    k1_stream_kernel<T, VEC, U, MINB, FULL><<<(unsigned)grid, kThreads, 0, st>>>(prm);
    return check_launch("k1_stream_kernel");
}

template <typename T, int VEC, int MINB, int RS, int STAGES, int NS = 0, bool SH = false, bool RAG = false>
static int launch_tma(K1Params& prm, int64_t B, cudaStream_t st) {
    const int64_t tiles = ceil_div(ceil_div(prm.V, VEC), kThreads);
    prm.iter = prm.tiles_per_cta > 0 ? prm.tiles_per_cta : choose_iter(tiles * B, MINB, prm.N * prm.C, sizeof(T) == 8);
    prm.blocks_per_vol = ceil_div(tiles, prm.iter);
    const int64_t grid = prm.blocks_per_vol * B;
    if (grid > 0x7fffffffLL) return set_error(VALUES_ERR_UNSUPPORTED, "grid too large");
    auto kern = k1_tma_kernel<T, VEC, MINB, RS, STAGES, NS, SH, RAG>;
    const size_t smem = (size_t)STAGES * RS * (kThreads * 16 + (SH ? 16 : 0));
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return set_error(VALUES_ERR_CUDA, "cudaFuncSetAttribute(k1_tma_kernel) failed");
    kern<<<(unsigned)grid, kThreads + 32, smem, st>>>(prm);
    return check_launch("k1_tma_kernel");
}

// aligned: every row, stride and map is 16-byte aligned (vector loads / stores); shiftable: it is not, but
// the 16-byte granules around the stack lie inside its allocation, so the ring kernel's SH mode can run
template <typename T>
static int dispatch_k1(K1Params& prm, int64_t B, bool aligned, bool shiftable, cudaStream_t st) {
    constexpr int NV = In<T>::VEC;
    const int64_t V = prm.V;
    // The per-sample arg-max wants the rows sample-outer, everything else class-outer: it is its own sweep
    // (1.0 of the HBM peak), and the maps / scores / mean arg-max follow through the ring as if it had not
    // been asked for -- two sweeps at the roofline instead of one shared-memory kernel at 0.26-0.47 of it.
    if (prm.samax && prm.variant != K1_SAMPLE_OUTER) {
        // (unaligned stacks: the same sweep on single elements, lane-consecutive)
        const int rc = aligned ? launch_argmax<T, NV, false>(prm, B, st) : launch_argmax<T, 1, false>(prm, B, st);
        if (rc) return rc;
        prm.samax = nullptr;
        if (!prm.need_ent && !prm.amax) return VALUES_OK;
    }
    // fp64 stacks (the reference's 3D path, raw overlap sums with large magnitudes) keep the
    // reference's exact accumulation order: sample-outer kernel below.
    if (prm.need_ent && !prm.samax && sizeof(T) != 8) {  // fp32 / bf16: class-outer stream kernel
        if (!aligned) {
            if constexpr (sizeof(T) != 8) {
                // rows at every 16-byte phase: the ring kernel with element-strided ownership (127^3, N = 16,
                // C = 4: 0.81 of the HBM peak; the scalar kernel below 0.40; 8-row stages measured: no gain)
                if (shiftable && prm.variant != K1_STREAM) {
                    if (prm.N % 4 == 0) return launch_tma<T, NV, 3, 4, 4, 0, true>(prm, B, st);
                    if (prm.N % 5 == 0) return launch_tma<T, NV, 3, 5, 3, 0, true>(prm, B, st);
                    if (prm.N == 2) return launch_tma<T, NV, 3, 2, 8, 0, true>(prm, B, st);
                    if (prm.N == 1) return launch_tma<T, NV, 3, 1, 8, 0, true>(prm, B, st);
                    return launch_tma<T, NV, 3, 4, 4, 0, true, true>(prm, B, st);   // ragged last stage
                }
            }
            return launch_stream<T, 1, 4, 4, false>(prm, B, st);
        }
        // batch = U samples of one class; unguarded when U divides N (N = 4k: MC-dropout/TTA
        // 8, 16; N = 5k: the reference's 5-member ensembles and N = 10)
        // Occupancy beats register comfort here (measured on B200, profiles/r01b_k1_variants.txt):
        // 3 CTAs/SM with a few cold-path spills reaches 0.84 of the HBM peak, 2 CTAs/SM 0.66.
        const int v = prm.variant;
        if constexpr (sizeof(T) != 8) {
            // default: bulk-copy ring, RS samples per stage (one mbarrier round trip per stage)
            if (v != K1_STREAM) {
                // N = 8k (MC-dropout / TTA 8, 16): 8 rows per stage, consumed 4 at a time -- half the
                // mbarrier round trips and loop overhead per element (K1 is issue-bound, not HBM-bound:
                // 70 % issue slots); measured on cfg5 0.90 -> 0.94 of the HBM peak at 24 volumes per
                // launch, equal at 8, outputs bit-identical
                if (prm.N % 8 == 0 && v == K1_RING8X3) return launch_tma<T, NV, 2, 8, 3>(prm, B, st);
                if (prm.N % 8 == 0 && v != K1_RING4) return launch_tma<T, NV, 3, 8, 2>(prm, B, st);
                if (prm.N % 4 == 0) return launch_tma<T, NV, 3, 4, 4>(prm, B, st);
                if (prm.N % 5 == 0) return launch_tma<T, NV, 3, 5, 3>(prm, B, st);
                if (prm.N % 3 == 0) return launch_tma<T, NV, 3, 3, 5>(prm, B, st);
                if (prm.N == 2) return launch_tma<T, NV, 3, 2, 8>(prm, B, st);
                if (prm.N == 1) return launch_tma<T, NV, 3, 1, 8>(prm, B, st);
                // N = 7, 11, 13, 14, 17, ...: 4-row stages with a ragged last one (single-row stages: 0.65-0.72)
                return launch_tma<T, NV, 3, 4, 4, 0, false, true>(prm, B, st);
            }
        }
        if (prm.N % 4 == 0) return launch_stream<T, NV, 4, 3, true>(prm, B, st);
        if (prm.N % 5 == 0) return launch_stream<T, NV, 5, 3, true>(prm, B, st);
        return launch_stream<T, NV, 4, 3, false>(prm, B, st);
    }
    if constexpr (sizeof(T) == 8) {
        // fp64 stacks with the sample counts of the reference's configs (5-member ensembles, N = 10
        // MC dropout, TTA 8 / 16): class-outer ring kernel with one fp32 accumulator per sample
        if (prm.need_ent && !prm.samax && aligned && prm.variant != K1_SAMPLE_OUTER) {
            if (prm.N == 16) return launch_tma<T, NV, 2, 4, 4, 16>(prm, B, st);
            if (prm.N == 8) return launch_tma<T, NV, 3, 4, 4, 8>(prm, B, st);
            if (prm.N == 10) return launch_tma<T, NV, 2, 5, 3, 10>(prm, B, st);
            if (prm.N == 5) return launch_tma<T, NV, 3, 5, 3, 5>(prm, B, st);
            // other small ensembles / TTA counts (the shared-memory kernel below: 0.24-0.32 of the HBM peak)
            if (prm.N == 2) return launch_tma<T, NV, 3, 2, 8, 2>(prm, B, st);
            if (prm.N == 3) return launch_tma<T, NV, 3, 3, 5, 3>(prm, B, st);
            if (prm.N == 4) return launch_tma<T, NV, 3, 4, 4, 4>(prm, B, st);
            if (prm.N == 6) return launch_tma<T, NV, 3, 3, 5, 6>(prm, B, st);
            if (prm.N == 12) return launch_tma<T, NV, 2, 4, 4, 12>(prm, B, st);
            if (prm.N == 20) return launch_tma<T, NV, 2, 5, 3, 20>(prm, B, st);
        }
        // odd voxel counts (rows 8 bytes off): the same kernels in SH mode
        if (prm.need_ent && !prm.samax && !aligned && shiftable && prm.variant != K1_SAMPLE_OUTER) {
            if (prm.N == 16) return launch_tma<T, NV, 2, 4, 4, 16, true>(prm, B, st);
            if (prm.N == 8) return launch_tma<T, NV, 3, 4, 4, 8, true>(prm, B, st);
            if (prm.N == 10) return launch_tma<T, NV, 2, 5, 3, 10, true>(prm, B, st);
            if (prm.N == 5) return launch_tma<T, NV, 3, 5, 3, 5, true>(prm, B, st);
        }
    }
    // the arg-max of the mean alone (no maps, no scores): its segmented sweep
    if (!prm.need_ent && prm.amax && !prm.samax && prm.variant != K1_SAMPLE_OUTER)
        return aligned ? launch_argmax<T, NV, true>(prm, B, st) : launch_argmax<T, 1, true>(prm, B, st);
    // per-sample arg-max next to the maps, unaligned stacks: sample-outer kernel with class sums in shared memory
    int rc = 1;
    if (aligned) {
        // widest vector whose class sums fit in shared memory
        constexpr int V4 = NV > 4 ? 4 : NV;
        prm.blocks_per_vol = ceil_div(ceil_div(V, V4), kThreads);
        if (prm.blocks_per_vol * B > 0x7fffffffLL) return set_error(VALUES_ERR_UNSUPPORTED, "grid too large");
        rc = launch_smem<T, V4>(prm, prm.blocks_per_vol * B, st);
    }
    if (rc == 1) {
        prm.blocks_per_vol = ceil_div(V, kThreads);
        if (prm.blocks_per_vol * B > 0x7fffffffLL) return set_error(VALUES_ERR_UNSUPPORTED, "grid too large");
        rc = launch_smem<T, 1>(prm, prm.blocks_per_vol * B, st);
    }
    if (rc == 1)
        return set_error(VALUES_ERR_UNSUPPORTED, "C=%lld class sums do not fit in shared memory",
                         (long long)prm.C);
    return rc;
}

static int64_t k1_blocks_per_vol_upper(int64_t V) { return ceil_div(V, kThreads); }

}  // namespace vb

using namespace vb;

extern "C" size_t values_uncertainty_workspace_bytes(int64_t B, int64_t V, int dtype) {
    (void)dtype;
    if (B <= 0 || V <= 0) return 0;
    // [B, blocks, 9] fp64 block partials, then [B] arrival counters
    return (size_t)(B * k1_blocks_per_vol_upper(V) * 9) * sizeof(double) + (size_t)B * sizeof(unsigned int);
}

extern "C" int values_uncertainty_fused(const void* probs, int dtype, int64_t B, int64_t N,
                                        int64_t C, int64_t V, int64_t stride_b, int64_t stride_n,
                                        int64_t stride_c, float* pe, float* ee, float* mi,
                                        int64_t map_stride_b, uint8_t* mean_argmax,
                                        uint8_t* sample_argmax, double* scores, int64_t score_stride,
                                        const double* thresholds_host, void* workspace,
                                        size_t workspace_bytes, int variant, int tiles_per_cta,
                                        void* stream) {
    if (B < 0 || N <= 0 || C <= 0 || V < 0)
        return set_error(VALUES_ERR_INVALID_ARG, "bad sizes B=%lld N=%lld C=%lld V=%lld",
                         (long long)B, (long long)N, (long long)C, (long long)V);
    if (B == 0 || V == 0) return VALUES_OK;  // empty stack: nothing to do (pointers may be NULL)
    if (!probs) return set_error(VALUES_ERR_INVALID_ARG, "probs is NULL");
    if ((mean_argmax || sample_argmax) && C > 256)
        return set_error(VALUES_ERR_UNSUPPORTED, "uint8 arg-max needs C <= 256");
    if (stride_b < 0 || stride_n < 0 || stride_c < 0)
        return set_error(VALUES_ERR_INVALID_ARG, "negative strides");
    if (variant < 0 || variant >= K1_VARIANTS || tiles_per_cta < 0 || tiles_per_cta > 64)
        return set_error(VALUES_ERR_INVALID_ARG, "unknown variant %d / tiles_per_cta %d", variant, tiles_per_cta);
    if (B == 0 || V == 0) return VALUES_OK;
    cudaStream_t st = (cudaStream_t)stream;
    K1Params prm{};
    prm.probs = probs;
    prm.variant = variant; prm.tiles_per_cta = tiles_per_cta;
    prm.N = N; prm.C = C; prm.V = V; prm.sb = stride_b; prm.sn = stride_n; prm.sc = stride_c;
    prm.pe = pe; prm.ee = ee; prm.mi = mi; prm.amax = mean_argmax; prm.samax = sample_argmax;
    prm.need_ent = (pe || ee || mi || scores) ? 1 : 0;
    prm.inv_n = (float)(1.0 / (double)N);
    prm.so = map_stride_b > 0 ? map_stride_b : V;
    if (prm.so < V) return set_error(VALUES_ERR_INVALID_ARG, "map_stride_b < V");
    prm.has_thr = thresholds_host ? 1 : 0;
    for (int k = 0; k < 3; ++k) prm.thr_f[k] = thresholds_host ? (float)thresholds_host[k] : 0.f;
    if (N * C > 0x7fffffffLL) return set_error(VALUES_ERR_UNSUPPORTED, "N*C too large");
    if (scores) {
        const size_t need = values_uncertainty_workspace_bytes(B, V, dtype);
        if (!workspace || workspace_bytes < need)
            return set_error(VALUES_ERR_WORKSPACE, "workspace %zu < %zu bytes", workspace_bytes, need);
        prm.partials = reinterpret_cast<double*>(workspace);
        prm.counters = reinterpret_cast<unsigned int*>(prm.partials + B * k1_blocks_per_vol_upper(V) * 9);
        if (score_stride < 3) return set_error(VALUES_ERR_INVALID_ARG, "score_stride < 3");
        prm.scores = scores; prm.score_stride = score_stride;
        if (cudaMemsetAsync(prm.counters, 0, (size_t)B * sizeof(unsigned int), st) != cudaSuccess)
            return set_error(VALUES_ERR_CUDA, "cudaMemsetAsync(counters) failed");
    }
    const size_t es = dtype == VALUES_F64 ? 8 : dtype == VALUES_F32 ? 4 : 2;
    const int nv = 16 / (int)es;
    auto al = [](const void* p, size_t a) { return p == nullptr || ((uintptr_t)p % a) == 0; };
    const bool aligned = V % nv == 0 && stride_b % nv == 0 && stride_n % nv == 0 &&
                         stride_c % nv == 0 && prm.so % 4 == 0 && al(probs, 16) && al(pe, 16) && al(ee, 16) &&
                         al(mi, 16) && al(mean_argmax, nv) && al(sample_argmax, nv);
    bool shiftable = false;
    if (!aligned && ((uintptr_t)probs % es) == 0) {
        uintptr_t base = 0;
        size_t size = 0;
        if (allocation_range(probs, &base, &size)) {
            const uintptr_t first = (uintptr_t)probs;
            const uintptr_t last = first + (uintptr_t)((B - 1) * stride_b + (N - 1) * stride_n + (C - 1) * stride_c + V) * es;
            shiftable = (first & ~(uintptr_t)15) >= base && ((last + 15) & ~(uintptr_t)15) <= base + size;
        }
    }
    int rc;
    switch (dtype) {
        case VALUES_F32: rc = dispatch_k1<float>(prm, B, aligned, shiftable, st); break;
        case VALUES_F64: rc = dispatch_k1<double>(prm, B, aligned, shiftable, st); break;
        case VALUES_BF16: rc = dispatch_k1<__nv_bfloat16>(prm, B, aligned, shiftable, st); break;
        default: return set_error(VALUES_ERR_INVALID_ARG, "unknown dtype %d", dtype);
    }
    return rc;
}

// Test/benchmark hook: force the number of voxel tiles per CTA (0 = automatic).

extern "C" int values_one_minus_msr(const void* probs, int dtype, int64_t B, int64_t C, int64_t V,
                                    int64_t stride_b, int64_t stride_c, void* out, void* stream) {
    if (B < 0 || C <= 0 || V < 0) return set_error(VALUES_ERR_INVALID_ARG, "bad sizes");
    if (B == 0 || V == 0) return VALUES_OK;
    if (!probs || !out) return set_error(VALUES_ERR_INVALID_ARG, "NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case VALUES_F32: return launch_msr<float>(probs, B, C, V, stride_b, stride_c, out, st);
        case VALUES_F64: return launch_msr<double>(probs, B, C, V, stride_b, stride_c, out, st);
        case VALUES_BF16: return launch_msr<__nv_bfloat16>(probs, B, C, V, stride_b, stride_c, out, st);
        default: return set_error(VALUES_ERR_INVALID_ARG, "unknown dtype %d", dtype);
    }
}
