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
// HBM-bound: algorithmic bytes per voxel = N*C*sizeof(T) + 3*4 (+1 arg-max byte).
#include "common.cuh"

namespace vb {

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
    double thr[3];
    int need_ent;
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

// acc += p*log(p), NaN terms skipped (test_3D.py:490-494, 500-504).  A term is NaN exactly when
// p is 0 (0 * -inf), negative (log -> NaN) or NaN, i.e. when !(p > 0); +inf is NOT skipped.
__device__ __forceinline__ void accum_term(float& acc, float p) {
    const float t = fmaf(p, fast_logf(p), acc);
    acc = (p > 0.f) ? t : acc;  // select, not a branch
}
__device__ __forceinline__ void accum_term(float& acc, double p) {
    const double t = p * log(p);
    acc = (t == t) ? (float)((double)acc + t) : acc;  // add in fp64, round into fp32
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
    const int64_t o = b * prm.V + v0;
    if (prm.pe) store_f32<VEC>(prm.pe + o, pe);
    if (prm.ee) store_f32<VEC>(prm.ee + o, ee);
    if (prm.mi) store_f32<VEC>(prm.mi + o, mi);
    if (prm.amax) store_u8<VEC>(prm.amax + o, idx);
    if (prm.partials) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            const double m3[3] = {(double)pe[j], (double)ee[j], (double)mi[j]};
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                part[3 * k] += m3[k];
                if (m3[k] >= prm.thr[k]) { part[3 * k + 1] += m3[k]; part[3 * k + 2] += 1.0; }
            }
        }
    }
}

__device__ __forceinline__ void k1_write_partials(const K1Params& prm, double (&part)[9]) {
    __shared__ double red[9 * 8];
    block_sum<9>(part, red);
    if (threadIdx.x == 0) {
        double* dst = prm.partials + (int64_t)blockIdx.x * 9;
#pragma unroll
        for (int k = 0; k < 9; ++k) dst[k] = part[k];
    }
}

// ---- class sums in registers (compile-time C): the fast path.  Always computes the three
// maps; per-sample arg-max and the arg-max-only mode go through the shared-memory kernel.
template <typename T, int CT, int VEC, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) k1_reg_kernel(const K1Params prm) {
    using A = typename In<T>::acc_t;
    const int64_t b = blockIdx.x / prm.blocks_per_vol;
    const int64_t blk = blockIdx.x - b * prm.blocks_per_vol;
    const int64_t v0 = (blk * kThreads + threadIdx.x) * VEC;
    double part[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) part[k] = 0.0;
    if (v0 < prm.V) {
        const T* base = reinterpret_cast<const T*>(prm.probs) + b * prm.sb + v0;
        A S[CT][VEC];
        float E[VEC];
        {   // sample 0 initialises the class sums (torch.mean / np.mean start from the first row)
            A p[CT][VEC];
#pragma unroll
            for (int c = 0; c < CT; ++c) load_elems<T, VEC>(base + c * prm.sc, p[c]);
#pragma unroll
            for (int j = 0; j < VEC; ++j) E[j] = 0.f;
#pragma unroll
            for (int c = 0; c < CT; ++c) {
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    S[c][j] = p[c][j];
                    accum_term(E[j], p[c][j]);
                }
            }
        }
#pragma unroll 1
        for (int64_t n = 1; n < prm.N; ++n) {
            A p[CT][VEC];
            const T* pn = base + n * prm.sn;
#pragma unroll
            for (int c = 0; c < CT; ++c) load_elems<T, VEC>(pn + c * prm.sc, p[c]);
            float H[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j) H[j] = 0.f;
#pragma unroll
            for (int c = 0; c < CT; ++c) {
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    S[c][j] += p[c][j];
                    accum_term(H[j], p[c][j]);
                }
            }
#pragma unroll
            for (int j = 0; j < VEC; ++j) E[j] += H[j];
        }
        k1_epilogue<A, VEC, CT>(prm, b, v0, CT, [&](int c, int j) -> A { return S[c][j]; }, E, part);
    }
    if (prm.partials) k1_write_partials(prm, part);
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

template <typename T>
static int dispatch_k1(K1Params& prm, int64_t B, bool aligned, cudaStream_t st) {
    constexpr int NV = In<T>::VEC;
    const int64_t V = prm.V;
    // register path: class sums of one thread must fit comfortably in registers
    constexpr int kMaxRegC = NV >= 8 ? 4 : 8;
    if (aligned && prm.C <= kMaxRegC && prm.need_ent && !prm.samax) {
        prm.blocks_per_vol = ceil_div(ceil_div(V, NV), kThreads);
        const int64_t grid = prm.blocks_per_vol * B;
        if (grid > 0x7fffffffLL) return set_error(VALUES_ERR_UNSUPPORTED, "grid too large");
#define VB_CASE(CT, MINB) \
    case CT: k1_reg_kernel<T, CT, NV, MINB><<<(unsigned)grid, kThreads, 0, st>>>(prm); break;
        switch (prm.C) {
            VB_CASE(1, 4) VB_CASE(2, 4) VB_CASE(3, 3) VB_CASE(4, 3)
            default:
                if constexpr (kMaxRegC >= 8) {
                    switch (prm.C) { VB_CASE(5, 2) VB_CASE(6, 2) VB_CASE(7, 2) VB_CASE(8, 2) }
                }
        }
#undef VB_CASE
        return check_launch("k1_reg_kernel");
    }
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
    return (size_t)(B * k1_blocks_per_vol_upper(V) * 9) * sizeof(double);
}

extern "C" int values_uncertainty_fused(const void* probs, int dtype, int64_t B, int64_t N,
                                        int64_t C, int64_t V, int64_t stride_b, int64_t stride_n,
                                        int64_t stride_c, float* pe, float* ee, float* mi,
                                        uint8_t* mean_argmax, uint8_t* sample_argmax,
                                        double* scores, const double* thresholds_host,
                                        void* workspace, size_t workspace_bytes, void* stream) {
    if (B < 0 || N <= 0 || C <= 0 || V < 0)
        return set_error(VALUES_ERR_INVALID_ARG, "bad sizes B=%lld N=%lld C=%lld V=%lld",
                         (long long)B, (long long)N, (long long)C, (long long)V);
    if (B == 0 || V == 0) return VALUES_OK;  // empty stack: nothing to do (pointers may be NULL)
    if (!probs) return set_error(VALUES_ERR_INVALID_ARG, "probs is NULL");
    if ((mean_argmax || sample_argmax) && C > 256)
        return set_error(VALUES_ERR_UNSUPPORTED, "uint8 arg-max needs C <= 256");
    if (stride_b < 0 || stride_n < 0 || stride_c < 0)
        return set_error(VALUES_ERR_INVALID_ARG, "negative strides");
    if (B == 0 || V == 0) return VALUES_OK;
    cudaStream_t st = (cudaStream_t)stream;
    K1Params prm{};
    prm.probs = probs;
    prm.N = N; prm.C = C; prm.V = V; prm.sb = stride_b; prm.sn = stride_n; prm.sc = stride_c;
    prm.pe = pe; prm.ee = ee; prm.mi = mi; prm.amax = mean_argmax; prm.samax = sample_argmax;
    prm.need_ent = (pe || ee || mi || scores) ? 1 : 0;
    for (int k = 0; k < 3; ++k)
        prm.thr[k] = thresholds_host ? thresholds_host[k] : __builtin_inf();
    if (scores) {
        const size_t need = values_uncertainty_workspace_bytes(B, V, dtype);
        if (!workspace || workspace_bytes < need)
            return set_error(VALUES_ERR_WORKSPACE, "workspace %zu < %zu bytes", workspace_bytes, need);
        prm.partials = reinterpret_cast<double*>(workspace);
    }
    const size_t es = dtype == VALUES_F64 ? 8 : dtype == VALUES_F32 ? 4 : 2;
    const int nv = 16 / (int)es;
    auto al = [](const void* p, size_t a) { return p == nullptr || ((uintptr_t)p % a) == 0; };
    const bool aligned = V % nv == 0 && stride_b % nv == 0 && stride_n % nv == 0 &&
                         stride_c % nv == 0 && al(probs, 16) && al(pe, 16) && al(ee, 16) &&
                         al(mi, 16) && al(mean_argmax, nv) && al(sample_argmax, nv);
    int rc;
    switch (dtype) {
        case VALUES_F32: rc = dispatch_k1<float>(prm, B, aligned, st); break;
        case VALUES_F64: rc = dispatch_k1<double>(prm, B, aligned, st); break;
        case VALUES_BF16: rc = dispatch_k1<__nv_bfloat16>(prm, B, aligned, st); break;
        default: return set_error(VALUES_ERR_INVALID_ARG, "unknown dtype %d", dtype);
    }
    if (rc != VALUES_OK) return rc;
    if (scores) return launch_reduce_partials(prm.partials, B, prm.blocks_per_vol, 9, scores, st);
    return VALUES_OK;
}

extern "C" int values_one_minus_msr(const void* probs, int dtype, int64_t B, int64_t C, int64_t V,
                                    int64_t stride_b, int64_t stride_c, void* out, void* stream) {
    if (B < 0 || C <= 0 || V < 0) return set_error(VALUES_ERR_INVALID_ARG, "bad sizes");
    if (B == 0 || V == 0) return VALUES_OK;
    if (!probs || !out) return set_error(VALUES_ERR_INVALID_ARG, "NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t bpv = ceil_div(V, kThreads);
    if (bpv * B > 0x7fffffffLL) return set_error(VALUES_ERR_UNSUPPORTED, "grid too large");
    const unsigned grid = (unsigned)(bpv * B);
    switch (dtype) {
        case VALUES_F32:
            msr_kernel<float><<<grid, kThreads, 0, st>>>((const float*)probs, C, V, stride_b,
                                                         stride_c, (float*)out, bpv);
            break;
        case VALUES_F64:
            msr_kernel<double><<<grid, kThreads, 0, st>>>((const double*)probs, C, V, stride_b,
                                                          stride_c, (double*)out, bpv);
            break;
        case VALUES_BF16:
            msr_kernel<__nv_bfloat16><<<grid, kThreads, 0, st>>>(
                (const __nv_bfloat16*)probs, C, V, stride_b, stride_c, (__nv_bfloat16*)out, bpv);
            break;
        default: return set_error(VALUES_ERR_INVALID_ARG, "unknown dtype %d", dtype);
    }
    return check_launch("msr_kernel");
}
