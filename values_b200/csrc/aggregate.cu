// K2: the evaluation/uncertainty_aggregation strategies on stored uncertainty maps.
//   K2a map_reduce : image-level sum + threshold sum/count   (aggregate_uncertainties.py:34-37, 40-67)
//   K2b patch_max  : fp64 sliding box-sum, max, first-isclose index      (:13-31)
//   normalize_maps : map / clip(count, 1) in fp64                        (data_carrier_3D.py:326-329)
// All HBM/L2-bound; algorithmic bytes = sizeof(T) per map voxel, outputs O(1).
#include <algorithm>
#include <type_traits>

#include "common.cuh"
#include "tma_host.cuh"

namespace vb {

// K2b implementations, selected per call (values_patch_max `path`): 0 = automatic (10x10 in-plane
// patches: fp32 strip filter by TMA in front of the exact fp64 march kernel; else the fused tile
// kernel when the patch fits its shared memory; else the generic tiled path), 5 = march kernel
// without the filter, 4 = fused tile kernel even where the march applies, 2 = generic tiled path.

// Where K2b puts a map's result: max_score[m * score_stride]; the box corner (z, y, x) at
// bbox[m * bbox_stride + 0..2] as int64 or -- for the fp64 score table [B, 3, 7] the pipeline hands
// to its callers -- as doubles.  (-1, -1, -1) = no window is close to the maximum (NaN map: the
// reference raises IndexError).
struct PatchOut {
    double* max_score; int64_t score_stride;
    int64_t* bbox_i64; double* bbox_f64; int64_t bbox_stride;
    __device__ __forceinline__ void put_score(int64_t m, double g) const { max_score[m * score_stride] = g; }
    __device__ __forceinline__ void put_corner(int64_t m, unsigned long long best, int64_t O1, int64_t O2) const {
        int64_t c[3] = {-1, -1, -1};
        if (best != ~0ull) {
            const int64_t lin = (int64_t)best;
            c[2] = lin % O2; c[1] = (lin / O2) % O1; c[0] = lin / (O2 * O1);
        }
        for (int d = 0; d < 3; ++d) {
            if (bbox_f64) bbox_f64[m * bbox_stride + d] = (double)c[d];
            else bbox_i64[m * bbox_stride + d] = c[d];
        }
    }
    PatchOut offset(int64_t m0) const {
        PatchOut o = *this;
        o.max_score += m0 * score_stride;
        if (o.bbox_f64) o.bbox_f64 += m0 * bbox_stride; else o.bbox_i64 += m0 * bbox_stride;
        return o;
    }
};

// =============================================================================== K2a
struct ThrTable { double v[16]; int n; };

constexpr int kReduceEPT = 16;  // elements per thread per block

template <typename T>
__global__ void __launch_bounds__(kThreads) map_reduce_kernel(const T* __restrict__ maps, int64_t V,
                                                              int64_t stride_m, int64_t bpm,
                                                              ThrTable thr,
                                                              double* __restrict__ partials) {
    __shared__ double red[3 * 8];
    const int64_t m = blockIdx.x / bpm;
    const int64_t blk = blockIdx.x - m * bpm;
    const T* src = maps + m * stride_m;
    const double t = thr.n > 0 ? thr.v[m % thr.n] : __longlong_as_double(0x7ff0000000000000LL);
    double acc[3] = {0.0, 0.0, 0.0};
    const int64_t base = blk * (int64_t)(kThreads * kReduceEPT) + threadIdx.x;
#pragma unroll 4
    for (int i = 0; i < kReduceEPT; ++i) {
        const int64_t v = base + (int64_t)i * kThreads;
        if (v < V) {
            const double x = (double)In<T>::load_one(src + v);
            acc[0] += x;
            if (x >= t) { acc[1] += x; acc[2] += 1.0; }
        }
    }
    block_sum<3>(acc, red);
    if (threadIdx.x == 0) {
        double* dst = partials + (int64_t)blockIdx.x * 3;
        dst[0] = acc[0]; dst[1] = acc[1]; dst[2] = acc[2];
    }
}

// =============================================================================== normalize
template <typename T>
__global__ void __launch_bounds__(kThreads) normalize_kernel(const T* __restrict__ maps, int64_t V,
                                                             int64_t stride_m, int64_t bpm,
                                                             const double* __restrict__ count,
                                                             double clip_min,
                                                             double* __restrict__ out) {
    const int64_t m = blockIdx.x / bpm;
    const int64_t v = (blockIdx.x - m * bpm) * kThreads + threadIdx.x;
    if (v >= V) return;
    // np.clip(count, 1, None) for clip_min = 1; clip_min = 0 (weighted stitching): uncovered -> 1
    const double c = clip_min > 0.0 ? fmax(count[v], clip_min) : (count[v] > 0.0 ? count[v] : 1.0);
    out[m * V + v] = (double)In<T>::load_one(maps + m * stride_m + v) / c;
}

// Vector form (V % 4 == 0, 16 / 32-byte aligned rows): a thread owns four consecutive voxels of EVERY map,
// so the count is read once, not once per map, and every access is 16 bytes wide.  Same arithmetic (one
// IEEE division per element).  cfg3 (3 fp32 maps of 256^3): 0.28 ms -> see DESIGN.md.
template <typename T>
__global__ void __launch_bounds__(kThreads) normalize_vec_kernel(const T* __restrict__ maps, int64_t V, int64_t M,
                                                                 int64_t stride_m, const double* __restrict__ count,
                                                                 double clip_min, double* __restrict__ out) {
    const int64_t v = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * 4;
    if (v >= V) return;
    const double2 c01 = *reinterpret_cast<const double2*>(count + v), c23 = *reinterpret_cast<const double2*>(count + v + 2);
    double c[4] = {c01.x, c01.y, c23.x, c23.y};
#pragma unroll
    for (int q = 0; q < 4; ++q) c[q] = clip_min > 0.0 ? fmax(c[q], clip_min) : (c[q] > 0.0 ? c[q] : 1.0);
    for (int64_t m = 0; m < M; ++m) {
        double x[4];
        if constexpr (sizeof(T) == 4) {
            const float4 r = *reinterpret_cast<const float4*>(maps + m * stride_m + v);
            x[0] = (double)r.x; x[1] = (double)r.y; x[2] = (double)r.z; x[3] = (double)r.w;
        } else {
            const double2 a = *reinterpret_cast<const double2*>(maps + m * stride_m + v);
            const double2 b = *reinterpret_cast<const double2*>(maps + m * stride_m + v + 2);
            x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y;
        }
        double* o = out + m * V + v;
        *reinterpret_cast<double2*>(o) = make_double2(x[0] / c[0], x[1] / c[1]);
        *reinterpret_cast<double2*>(o + 2) = make_double2(x[2] / c[2], x[3] / c[3]);
    }
}

// =============================================================================== K2b
constexpr int kTX = 32;   // outputs per tile along the contiguous axis
constexpr int kTY = 16;   // outputs per tile along axis 1
constexpr int kSeg = 8;   // outputs per sliding x-segment
constexpr int kOwn = kTX * kTY / kThreads;  // outputs owned per thread per plane (2)

struct PatchParams {
    const void* maps;
    int64_t stride_m;
    int64_t D0, D1, D2;   // map shape (unused leading axes = 1)
    int64_t O0, O1, O2;   // number of windows per axis
    int p0, p1, p2;
    int tiles_x, tiles_y, chunks_z, zc;
    int64_t ntiles;
    double denom;         // prod(patch) when mean, else 1
    int mean_flag;
    double rtol, atol;
    double* tile_max;               // [M, ntiles]
    const double* gmax;             // [M]       (pass 2)
    unsigned long long* best;       // [M]       (pass 2)
};

__device__ __forceinline__ bool np_isclose(double a, double b, double rtol, double atol) {
    // numpy.isclose(a, b): |a-b| <= atol + rtol*|b|; equal infinities are close; NaN never is
    if (a == b) return true;
    if (isinf(a) || isinf(b)) return false;
    return fabs(a - b) <= atol + rtol * fabs(b);
}
__device__ __forceinline__ double nanmax(double m, double v) {  // np.max: NaN propagates
    return (v > m || v != v) ? v : m;
}

template <typename T, int PASS>
__global__ void __launch_bounds__(kThreads) patch_kernel(const PatchParams prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int R = kTY + prm.p1 - 1;            // input rows per plane tile
    const int W = kTX + prm.p2 - 1;            // input cols per plane tile
    const int pitch = W | 1;                   // odd pitch: rows land in distinct 64-bit banks
    double* in_tile = reinterpret_cast<double*>(smem_raw);          // [R][pitch]
    double* rowsum = in_tile + (size_t)R * pitch;                   // [R][kTX]
    double* ring = rowsum + (size_t)R * kTX;                        // [p0][kTX*kTY]
    __shared__ double red[8];

    const int64_t m = blockIdx.y;
    int tile = blockIdx.x;
    const int tx_i = tile % prm.tiles_x; tile /= prm.tiles_x;
    const int ty_i = tile % prm.tiles_y; tile /= prm.tiles_y;
    const int zc_i = tile;
    if (PASS == 2) {
        const double tm = prm.tile_max[m * prm.ntiles + blockIdx.x];
        if (!np_isclose(tm, prm.gmax[m], prm.rtol, prm.atol)) return;  // exact cull (monotone)
    }
    const int64_t x0 = (int64_t)tx_i * kTX, y0 = (int64_t)ty_i * kTY;
    const int64_t zo0 = (int64_t)zc_i * prm.zc;                       // first output plane
    const int64_t zo1 = min(zo0 + prm.zc, prm.O0);                    // one past last output plane
    const T* src = reinterpret_cast<const T*>(prm.maps) + m * prm.stride_m;
    const int tid = threadIdx.x;
    const int lx = tid % kTX, ly = tid / kTX;  // owns outputs (ly + 8k, lx)

    double run[kOwn];
#pragma unroll
    for (int k = 0; k < kOwn; ++k) run[k] = 0.0;
    double tmax = -__longlong_as_double(0x7ff0000000000000LL);
    unsigned long long tbest = ~0ull;
    const double gmax = PASS == 2 ? prm.gmax[m] : 0.0;

    const int64_t nplanes = (zo1 - zo0) + prm.p0 - 1;
    for (int64_t zi = 0; zi < nplanes; ++zi) {
        const int64_t z = zo0 + zi;
        // 1. stage the input plane tile (coalesced along x), widened to fp64
        const T* plane = src + z * prm.D1 * prm.D2;
        for (int idx = tid; idx < R * W; idx += kThreads) {
            const int r = idx / W, cx = idx - r * W;
            const int64_t y = y0 + r, x = x0 + cx;
            double v = 0.0;
            if (y < prm.D1 && x < prm.D2) v = (double)In<T>::load_one(plane + y * prm.D2 + x);
            in_tile[r * pitch + cx] = v;
        }
        __syncthreads();
        // 2. x-pass: sliding sums in registers, kSeg outputs per task
        for (int task = tid; task < R * (kTX / kSeg); task += kThreads) {
            const int r = task % R, seg = task / R;
            const double* row = in_tile + r * pitch + seg * kSeg;
            double s = 0.0;
            for (int k = 0; k < prm.p2; ++k) s += row[k];
            double* dst = rowsum + r * kTX + seg * kSeg;
            dst[0] = s;
#pragma unroll
            for (int i = 1; i < kSeg; ++i) {
                s += row[i + prm.p2 - 1] - row[i - 1];
                dst[i] = s;
            }
        }
        __syncthreads();
        // 3. y-pass (direct) + z-pass (sliding over a private ring column)
        const int slot = (int)(zi % prm.p0);
#pragma unroll
        for (int k = 0; k < kOwn; ++k) {
            const int oy = ly + k * (kThreads / kTX);
            double s2 = 0.0;
            for (int j = 0; j < prm.p1; ++j) s2 += rowsum[(oy + j) * kTX + lx];
            double* rs = ring + (size_t)slot * (kTX * kTY) + oy * kTX + lx;
            const double old = zi >= prm.p0 ? *rs : 0.0;
            *rs = s2;
            run[k] += s2 - old;
            if (zi >= prm.p0 - 1) {
                const int64_t oz = z - (prm.p0 - 1);
                const int64_t y = y0 + oy, x = x0 + lx;
                if (y < prm.O1 && x < prm.O2) {
                    const double v = prm.mean_flag ? run[k] / prm.denom : run[k];
                    if (PASS == 1) {
                        tmax = nanmax(tmax, v);
                    } else if (np_isclose(v, gmax, prm.rtol, prm.atol)) {
                        const unsigned long long lin =
                            (unsigned long long)((oz * prm.O1 + y) * prm.O2 + x);
                        tbest = lin < tbest ? lin : tbest;
                    }
                }
            }
        }
        __syncthreads();
    }
    if (PASS == 1) {
        // block max with NaN propagation
        const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tmax = nanmax(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
        if (lane == 0) red[warp] = tmax;
        __syncthreads();
        if (tid == 0) {
            double mm = red[0];
            for (int w = 1; w < kThreads / 32; ++w) mm = nanmax(mm, red[w]);
            prm.tile_max[m * prm.ntiles + blockIdx.x] = mm;
        }
    } else {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, tbest, o);
            tbest = other < tbest ? other : tbest;
        }
        if ((tid & 31) == 0 && tbest != ~0ull) atomicMin(prm.best + m, tbest);  // min is order-free
    }
}

// one block per map: global max over tile maxima; reset the first-index slot
__global__ void __launch_bounds__(kThreads) patch_select_kernel(const double* __restrict__ tile_max,
                                                                int64_t ntiles,
                                                                double* __restrict__ gmax,
                                                                PatchOut out,
                                                                unsigned long long* __restrict__ best) {
    __shared__ double red[8];
    const int64_t m = blockIdx.x;
    double mm = -__longlong_as_double(0x7ff0000000000000LL);
    for (int64_t i = threadIdx.x; i < ntiles; i += kThreads) mm = nanmax(mm, tile_max[m * ntiles + i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mm = nanmax(mm, __shfl_xor_sync(0xffffffffu, mm, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mm;
    __syncthreads();
    if (threadIdx.x == 0) {
        double g = red[0];
        for (int w = 1; w < kThreads / 32; ++w) g = nanmax(g, red[w]);
        gmax[m] = g;
        out.put_score(m, g);
        best[m] = ~0ull;
    }
}

__global__ void patch_finish_kernel(const unsigned long long* __restrict__ best, int64_t M,
                                    int64_t O1, int64_t O2, PatchOut out) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    out.put_corner(m, best[m], O1, O2);
}

// ------------------------------------------------------------------ K2b fused tile kernel
// One kernel per pass does all three box passes: a CTA owns TY x TX windows in-plane and a
// chunk of ZC output planes, marching over z.  Per input plane: (1) the plane tile (+halo)
// goes from L2 to shared memory through registers, one plane AHEAD of the arithmetic;
// (2) x-pass: sliding fp64 sums in registers, 8 outputs per task, rows striped over lanes
// (odd pitch -> conflict-free); (3) y-pass: a thread owns OWN consecutive rows of one column
// (sliding), z-pass: sliding over a private shared-memory ring of the last p0 plane sums.
// Two block barriers per plane.  Pass 1 keeps the tile maximum; pass 2 recomputes only the
// tiles whose maximum is np.isclose to the global one and takes the minimum C-order index;
// the last CTA of a map (ticket) converts it to the bounding box, so there is no select /
// finish launch.  fp64 throughout: the result is independent of tiling and launch shape.
constexpr int kFusedThreads = 256;   // measured: the march is shared-memory-bandwidth bound, so fewer
                                     // threads with more outputs each (register reuse) beat 512
constexpr int kMaxActive = 64;       // pass-2 work list entries per map (more -> full re-walk)

template <int TY, int TX> struct FusedTile {
    static constexpr int kSegF = 8;                          // outputs per x-pass task
    static constexpr int kOwnF = TY * TX / kFusedThreads;    // consecutive y outputs per thread
    static constexpr int kStage = 8;                         // staged input elements per thread
    static_assert(kOwnF >= 1 && kOwnF * (kFusedThreads / TX) == TY, "tile shape");
};

struct FusedParams {
    const void* maps;
    int64_t stride_m;
    int64_t D0, D1, D2, O0, O1, O2;
    int64_t pitch;                // elements between rows (== D2, or D2 rounded up to 4 for the pitched scratch copy)
    int p0, p1, p2;
    int tiles_x, tiles_y, chunks_z, zc;
    int zsub, zc_fine;            // pass 2 splits a pass-1 z-chunk into zsub pieces of zc_fine planes
    int64_t ntiles;
    int64_t nent;                 // tile_max entries per map: ntiles (fused) or ntiles * zsub (march)
    int tx_per, xs_stride;        // march tiles: x origin of tile column tx = (tx / tx_per) * xs_stride + (tx % tx_per) * 64
    double denom;
    int mean_flag;
    double rtol, atol;
    double* tile_max;             // [M, nent]
    double* gmax;                 // [M]   written by the last pass-1 CTA of a map
    int* active;                  // [M, 1 + kMaxActive]: count (or -1 = walk every tile), tiles
    unsigned long long* best;     // [M]
    unsigned int* tickets;        // [M, 4]  filter / pass-1 / pass-2 arrival counters, max |input| bits
    int use_list;                 // pass 1 walks only the (tile, z sub-chunk) entries the fp32 filter listed
    double err_coef;              // fp32 filter: |fp32 box sum - true box sum| <= err_coef * max |input|
    PatchOut out;                 // max_score / box corner of every map
};

__device__ __noinline__ double box_mean_div(double s, double denom) { return s / denom; }  // mean=True only

// sum of N consecutive doubles as a balanced tree (short dependency chains)
template <int N> __device__ __forceinline__ double tree_sum(const double* v) {
    if constexpr (N == 1) return v[0];
    else return tree_sum<N / 2>(v) + tree_sum<N - N / 2>(v + N / 2);
}

// Shared tail of the fused / march kernels.  PASS 1: tile maximum -> tile_max; the last CTA of a
// map reduces them to the map maximum and lists the tiles np.isclose to it.  PASS 2: minimum
// C-order index -> best (atomicMin, order-free); the last CTA publishes the bounding-box corner.
template <int PASS, int NT = kFusedThreads, bool FINE = false>
__device__ __forceinline__ void box_pass_finish(const FusedParams& prm, int64_t m, double tmax,
                                                unsigned long long tbest, double* red, int& s_flag,
                                                int& s_count, float amax = 0.f,
                                                unsigned int n_ctas = gridDim.x) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double ninf = -__longlong_as_double(0x7ff0000000000000LL);
    if constexpr (PASS == 0) {
        // fp32 filter pass (march kernel only): tile_max holds fp32 sub-chunk maxima of box SUMS.
        // The last CTA of a map turns them into the pass-1 work list: every (tile, z sub-chunk)
        // whose true maximum could still be np.isclose to the true map maximum, given
        // |fp32 sum - true sum| <= E = err_coef * max |input|.  With g the fp32 map maximum the
        // true maximum G lies in [g - E, g + E]; a sub-chunk with fp32 maximum t holds a window
        // close to G only if t + E >= G - atol - rtol |G| >= g - E - atol - rtol (|g| + E).
        // Entries not listed are reset to -inf, so the exact pass-1 finish ignores them.  Anything
        // non-finite (NaN / inf inputs, fp32 overflow) or a list that overflows sends the whole
        // map through the exact pass.
        const unsigned int abits = __reduce_max_sync(0xffffffffu, __float_as_uint(amax));  // amax >= +0
        if (lane == 0) atomicMax(prm.tickets + 4 * m + 3, abits);
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            s_flag = atomicAdd(prm.tickets + 4 * m, 1u) == n_ctas - 1;
            s_count = 0;
        }
        __syncthreads();
        if (!s_flag) return;
        __threadfence();
        double* tm = prm.tile_max + m * prm.nent;
        double mm = ninf;
        for (int64_t i = tid; i < prm.nent; i += NT) mm = nanmax(mm, __ldcg(tm + i));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mm = nanmax(mm, __shfl_xor_sync(0xffffffffu, mm, o));
        if (lane == 0) red[warp] = mm;
        __syncthreads();
        double g = red[0];
        for (int w = 1; w < NT / 32; ++w) g = nanmax(g, red[w]);
        const double a = (double)__uint_as_float(atomicMax(prm.tickets + 4 * m + 3, 0u));  // atomic read
        const double E = prm.err_coef * a;
        const double atol_s = prm.mean_flag ? prm.atol * prm.denom : prm.atol;   // the filter compares sums
        const double cut = g - 2.0 * E - atol_s - prm.rtol * (fabs(g) + E);
        // false for NaN / inf inputs, and when an fp32 box sum could overflow
        const bool finite = fabs(g) < 1e300 && a * ((double)prm.p0 * prm.p1 * prm.p2) < 1e37 && E < 1e300;
        int* lst = prm.active + m * (1 + kMaxActive);
        if (finite) {
            for (int64_t i = tid; i < prm.nent; i += NT) {
                if (__ldcg(tm + i) >= cut) {
                    const int n = atomicAdd(&s_count, 1);
                    if (n < kMaxActive) lst[1 + n] = (int)i;
                } else {
                    tm[i] = ninf;
                }
            }
        }
        __syncthreads();
        if (tid == 0) lst[0] = (!finite || s_count > kMaxActive) ? -1 : s_count;
    } else if constexpr (PASS == 1) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tmax = nanmax(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
        if (lane == 0) red[warp] = tmax;
        __syncthreads();
        if (tid == 0) {
            if (!FINE) {   // FINE: the march already stored one maximum per z sub-chunk
                double mm = red[0];
                for (int w = 1; w < NT / 32; ++w) mm = nanmax(mm, red[w]);
                prm.tile_max[m * prm.nent + blockIdx.x] = mm;
            }
            __threadfence();
            s_flag = atomicAdd(prm.tickets + 4 * m + 1, 1u) == n_ctas - 1;
            s_count = 0;
        }
        __syncthreads();
        if (!s_flag) return;
        // last CTA of this map: map maximum (NaN propagates, as np.max) and the pass-2 work list
        __threadfence();
        const double* tm = prm.tile_max + m * prm.nent;
        double mm = ninf;
        for (int64_t i = tid; i < prm.nent; i += NT) mm = nanmax(mm, __ldcg(tm + i));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mm = nanmax(mm, __shfl_xor_sync(0xffffffffu, mm, o));
        if (lane == 0) red[warp] = mm;
        __syncthreads();
        double g = red[0];
        for (int w = 1; w < NT / 32; ++w) g = nanmax(g, red[w]);
        int* lst = prm.active + m * (1 + kMaxActive);
        for (int64_t i = tid; i < prm.nent; i += NT) {   // list order is irrelevant (min index wins)
            if (np_isclose(__ldcg(tm + i), g, prm.rtol, prm.atol)) {
                const int n = atomicAdd(&s_count, 1);
                if (n < kMaxActive) lst[1 + n] = (int)i;
            }
        }
        __syncthreads();
        if (tid == 0) {
            if (atomicOr(prm.tickets + 4 * m + 2, 0u) & 0x80000000u) {   // score and corner already written by the
                lst[0] = 0;                                              // CTA that walked the single candidate
            } else {
                lst[0] = s_count > kMaxActive ? -1 : s_count;
                prm.gmax[m] = g;
                prm.out.put_score(m, g);
                prm.best[m] = ~0ull;
            }
        }
    } else {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, tbest, o);
            tbest = other < tbest ? other : tbest;
        }
        if (lane == 0 && tbest != ~0ull) atomicMin(prm.best + m, tbest);  // min is order-free
        // last CTA of this map: publish the bounding-box corner
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            s_flag = atomicAdd(prm.tickets + 4 * m + 2, 1u) == n_ctas - 1;
        }
        __syncthreads();
        if (s_flag && tid == 0) {
            __threadfence();
            const unsigned long long b = atomicMin(prm.best + m, ~0ull);  // atomic read
            prm.out.put_corner(m, b, prm.O1, prm.O2);
        }
    }
}

// One kernel does all three box passes: a CTA owns TY x TX windows in-plane and a chunk of
// output planes, marching over z.  Per input plane: (1) the plane tile (+halo) goes from L2
// to shared memory through registers, one plane AHEAD of the arithmetic, widened to fp64;
// (2) x-pass: tasks of 8 outputs, rows striped over lanes (odd pitch -> conflict-free);
// (3) y-pass: a thread owns kOwnF consecutive rows of one column, z-pass: sliding over a
// shared-memory ring of the last p0 plane sums.  Two block barriers per plane.
// PASS 1 keeps the tile maximum; the last CTA of a map (ticket) reduces them to the map
// maximum and lists the tiles whose maximum is np.isclose to it.  PASS 2 re-walks only the
// listed tiles (split finer in z) for the minimum C-order index; its last CTA converts it to
// the bounding-box corner.  fp64 throughout: results do not depend on tiling or launch shape.
// PC > 0: in-plane patch extents p1 == p2 == PC known at compile time (the reference's configs
// all use 10: evaluation/configs/tasks/aggregation_patch_*.yaml), so both box loops unroll.
template <typename T, int TY, int TX, int PC, int PASS>
__global__ void __launch_bounds__(kFusedThreads, 2) box_fused_kernel(const FusedParams prm) {
    using FT = FusedTile<TY, TX>;
    constexpr int NT = kFusedThreads;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[NT / 32];
    __shared__ int s_flag;
    __shared__ int s_count;
    const int p1 = PC > 0 ? PC : prm.p1, p2 = PC > 0 ? PC : prm.p2, p0 = prm.p0;
    const int R = TY + p1 - 1, W = TX + p2 - 1;
    const int pitch_in = W | 1, pitch_rs = TX + 1;
    double* rowsum = reinterpret_cast<double*>(smem_raw);     // [R][pitch_rs]
    double* ring = rowsum + (size_t)R * pitch_rs;             // [p0][TY*TX]
    double* in_tile = ring + (size_t)p0 * TY * TX;            // [R][pitch_in], widened to fp64

    const int64_t m = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double ninf = -__longlong_as_double(0x7ff0000000000000LL);
    const int tiles_xy = prm.tiles_x * prm.tiles_y;
    const int zsub = PASS == 2 ? prm.zsub : 1;
    const int zc_fine = PASS == 2 ? prm.zc_fine : prm.zc;
    double gmax = 0.0;
    unsigned long long tbest = ~0ull;
    double tmax = ninf;
    // ---- work items of this CTA: pass 1 = its own tile; pass 2 = (listed tile, z sub-chunk)
    int n_work = 1, work = 0, work_step = 1;
    const int* list = nullptr;
    if (PASS == 2) {
        gmax = prm.gmax[m];
        list = prm.active + m * (1 + kMaxActive);
        const int n_act = list[0];
        n_work = (n_act < 0 ? (int)prm.ntiles : n_act) * zsub;
        work = blockIdx.x; work_step = gridDim.x;
    }
    for (; work < n_work; work += work_step) {
        int tile, fine;
        if (PASS == 1) { tile = blockIdx.x; fine = 0; }
        else { const int a = work / zsub; fine = work - a * zsub; tile = list[0] < 0 ? a : list[1 + a]; }
        const int tx_i = tile % prm.tiles_x;
        const int ty_i = (tile / prm.tiles_x) % prm.tiles_y;
        const int zc_i = tile / tiles_xy;
        const int64_t zo0 = (int64_t)zc_i * prm.zc + (int64_t)fine * zc_fine;
        const int64_t zo1 = min(zo0 + zc_fine, min((int64_t)(zc_i + 1) * prm.zc, prm.O0));
        if (zo0 >= zo1) continue;
        const int64_t x0 = (int64_t)tx_i * TX, y0 = (int64_t)ty_i * TY;
        const int nplanes = (int)(zo1 - zo0) + p0 - 1;
        const T* src = reinterpret_cast<const T*>(prm.maps) + m * prm.stride_m;
        const int64_t plane_elems = prm.D1 * prm.D2;

        // staging map of this thread: global offset inside a plane (-1 = zero fill) and the
        // shared-memory slot, fixed for the whole march
        const int n_in = R * W;
        int goff[FT::kStage], soff[FT::kStage];
#pragma unroll
        for (int i = 0; i < FT::kStage; ++i) {
            const int idx = tid + i * NT;
            goff[i] = -1; soff[i] = -1;
            if (idx < n_in) {
                const int r = idx / W, cx = idx - r * W;
                const int64_t y = y0 + r, x = x0 + cx;
                soff[i] = r * pitch_in + cx;
                if (y < prm.D1 && x < prm.D2) goff[i] = (int)(y * prm.D2 + x);
            }
        }
        const bool staged = n_in <= FT::kStage * NT && plane_elems < 0x7fffffffLL;
        T st[FT::kStage];
        auto fetch = [&](int64_t z) {
            const T* pl = src + z * plane_elems;
#pragma unroll
            for (int i = 0; i < FT::kStage; ++i) st[i] = goff[i] >= 0 ? In<T>::load_one(pl + goff[i]) : (T)0;
        };
        auto commit = [&]() {
#pragma unroll
            for (int i = 0; i < FT::kStage; ++i) if (soff[i] >= 0) in_tile[soff[i]] = (double)st[i];
        };
        auto load_direct = [&](int64_t z) {   // halo too large for the register stage
            const T* pl = src + z * plane_elems;
            for (int idx = tid; idx < n_in; idx += NT) {
                const int r = idx / W, cx = idx - r * W;
                const int64_t y = y0 + r, x = x0 + cx;
                in_tile[r * pitch_in + cx] =
                    (y < prm.D1 && x < prm.D2) ? (double)In<T>::load_one(pl + y * prm.D2 + x) : 0.0;
            }
        };
        // x-pass tasks: (row r, segment of kSegF outputs), rows striped over lanes
        const int ntask = R * (TX / FT::kSegF);
        const int task_r = tid % R, task_seg = tid / R;
        const int ox = tid % TX, oy0 = (tid / TX) * FT::kOwnF;   // y/z-pass ownership
        double run[FT::kOwnF];
        unsigned int valid = 0;   // bit k: output (oy0 + k, ox) of this thread lies inside the map
#pragma unroll
        for (int k = 0; k < FT::kOwnF; ++k) {
            run[k] = 0.0;
            if (y0 + oy0 + k < prm.O1 && x0 + ox < prm.O2) valid |= 1u << k;
        }
        auto x_task = [&](int r, int seg) {
            const double* row = in_tile + r * pitch_in + seg * FT::kSegF;
            double* dst = rowsum + r * pitch_rs + seg * FT::kSegF;
            if constexpr (PC > 0) {
                double v[PC + FT::kSegF - 1];
#pragma unroll
                for (int k = 0; k < PC + FT::kSegF - 1; ++k) v[k] = row[k];
                double s = tree_sum<PC>(v);
                dst[0] = s;
#pragma unroll
                for (int i = 1; i < FT::kSegF; ++i) { s += v[i + PC - 1] - v[i - 1]; dst[i] = s; }
            } else {
                double s = 0.0;
                for (int k = 0; k < p2; ++k) s += row[k];
                dst[0] = s;
#pragma unroll
                for (int i = 1; i < FT::kSegF; ++i) { s += row[i + p2 - 1] - row[i - 1]; dst[i] = s; }
            }
        };

        __syncthreads();   // previous work item of this CTA is done with the shared tiles
        if (staged) fetch(zo0);
        int slot = 0;
        for (int zi = 0; zi < nplanes; ++zi) {
            if (staged) commit(); else load_direct(zo0 + zi);
            __syncthreads();
            if (staged && zi + 1 < nplanes) fetch(zo0 + zi + 1);   // next plane, in flight during the math
            // ---- x-pass
            if (tid < ntask) x_task(task_r, task_seg);
            for (int task = tid + NT; task < ntask; task += NT) x_task(task % R, task / R);
            __syncthreads();
            // ---- y-pass (sliding down the strip) + z-pass (sliding over the ring)
            const double* col = rowsum + oy0 * pitch_rs + ox;
            double s2;
            if constexpr (PC > 0) {
                double c[PC];
#pragma unroll
                for (int j = 0; j < PC; ++j) c[j] = col[j * pitch_rs];
                s2 = tree_sum<PC>(c);
            } else {
                s2 = 0.0;
                for (int j = 0; j < p1; ++j) s2 += col[j * pitch_rs];
            }
            const bool warm = zi >= p0 - 1;     // the z-window is complete: an output plane
            const int64_t oz = zo0 + zi - (p0 - 1);
#pragma unroll
            for (int k = 0; k < FT::kOwnF; ++k) {
                if (k > 0) s2 += col[(k + p1 - 1) * pitch_rs] - col[(k - 1) * pitch_rs];
                double* rs = ring + (size_t)slot * (TY * TX) + (oy0 + k) * TX + ox;
                const double old = zi >= p0 ? *rs : 0.0;
                *rs = s2;
                run[k] += s2 - old;
                if (warm && ((valid >> k) & 1u)) {
                    double v = run[k];
                    if (prm.mean_flag) v = box_mean_div(v, prm.denom);
                    if (PASS == 1) {
                        tmax = nanmax(tmax, v);
                    } else if (np_isclose(v, gmax, prm.rtol, prm.atol)) {
                        const unsigned long long lin =
                            (unsigned long long)((oz * prm.O1 + (y0 + oy0 + k)) * prm.O2 + (x0 + ox));
                        tbest = lin < tbest ? lin : tbest;
                    }
                }
            }
            slot = slot + 1 == p0 ? 0 : slot + 1;
        }
    }
    box_pass_finish<PASS>(prm, m, tmax, tbest, red, s_flag, s_count);
}

// ------------------------------------------------------------------ K2b march kernel (z first)
// Same tiling, passes and finish protocol as box_fused_kernel, but the three box passes run
// z -> x -> y: a thread keeps the z-window sums of its input positions (rows g, g+G, ... of one
// tile column) in REGISTERS, sliding them with the entering plane and the leaving plane (both
// re-read through L2, one plane ahead of the arithmetic), so there is no shared-memory ring of
// plane sums.  That frees shared memory for 32 x 64 tiles (halo 1.28 x 1.14 instead of
// 1.56 x 1.14), lets the x-pass produce 16 and the y-pass 8 outputs per task (25 resp. 17
// shared loads), and makes the p0-1 warm-up planes of a z-chunk cost a register add instead of
// a full plane of x/y passes.  The three stages are software-pipelined over planes with
// double-buffered shared memory: ONE block barrier per output plane.
// Input rows / columns past the map edge are CLAMPED, not zero-filled: they only ever feed
// windows that are not fully inside the map, and those outputs are masked -- so every load is
// unconditional (uniform plane pointer + 32-bit per-thread offset).  Requires the in-plane patch
// extents to equal the compile-time PC (10 in all reference configs) and D1*D2 < 2^31.
template <int TY, int TX, int PC> struct MarchTile {
    static constexpr int NT = kFusedThreads;
    static constexpr int R = TY + PC - 1, W = TX + PC - 1;
    static constexpr int CW = (W + 15) / 16 * 16;              // threads per row group
    static constexpr int G = NT / CW;                          // row groups marching in lock step
    static constexpr int NSLOT = (R + G - 1) / G;              // input rows per thread
    static constexpr int pitchA = W | 1, pitchB = TX + 1;      // odd pitches: rows striped over lanes
    static constexpr int SEG = 16, NSEG = TX / SEG;            // x-pass task = 16 outputs of one row
    static constexpr int RUN = TY * TX / NT;                   // y-pass: consecutive rows per thread
    template <typename ACC> static constexpr size_t smem_bytes() {
        return (size_t)2 * (R * pitchA + R * pitchB) * sizeof(ACC);   // double-buffered
    }
    static_assert(G >= 1 && TX % SEG == 0 && RUN * (NT / TX) == TY && R * NSEG <= NT, "tile shape");
    static_assert(SEG % 2 == 0 && RUN % 2 == 0 && RUN / 2 <= PC && SEG / 2 <= PC, "two chains per task");
    static_assert((NSLOT - 1) * G <= R, "only the last slot of a thread can fall outside the tile");
};

template <int N, typename A> __device__ __forceinline__ A tree_sum_t(const A* v) {
    if constexpr (N == 1) return v[0];
    else return tree_sum_t<N / 2, A>(v) + tree_sum_t<N - N / 2, A>(v + N / 2);
}
__device__ __forceinline__ double acc_max(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float acc_max(float a, float b) { return fmaxf(a, b); }
constexpr int tree_depth(int n) { return n <= 1 ? 0 : 1 + tree_depth(n - n / 2); }

// PASS 1 after the fp32 strip filter (prm.use_list): only the (tile, z sub-chunk) entries the filter
// listed are walked; results are bit-identical to running PASS 1 over everything.
// FUSE (PASS 1 behind the filter): a map whose filter list holds ONE entry needs no second walk -- the
// CTA that walks the entry keeps its window sums (at most kFusePlanes planes of TY x TX doubles) in
// shared memory, takes their maximum and the first C-order index np.isclose to it, writes score and
// corner and marks the map done (bit 31 of its pass-2 ticket): pass 2 then finds an empty list.
constexpr int kFusePlanes = 2;
constexpr unsigned int kFusedDone = 0x80000000u;
template <typename T, typename ACC, int TY, int TX, int PC, int PASS, int MINB, bool FUSE = false>
__global__ void __launch_bounds__(kFusedThreads, MINB) box_march_kernel(const FusedParams prm) {
    using MT = MarchTile<TY, TX, PC>;
    constexpr int NT = kFusedThreads;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[NT / 32];
    __shared__ unsigned long long red_idx[NT / 32];
    __shared__ int s_flag;
    __shared__ int s_count;
    ACC* A = reinterpret_cast<ACC*>(smem_raw);                // [2][R][pitchA] z-window sums
    ACC* Bs = A + 2 * MT::R * MT::pitchA;                     // [2][R][pitchB] z-x sums
    double* saved = reinterpret_cast<double*>(Bs + 2 * MT::R * MT::pitchB);   // FUSE: [kFusePlanes][TY][TX]
    const int p0 = prm.p0;
    const int64_t m = blockIdx.y;
    const int tid = threadIdx.x;
    const double ninf = -__longlong_as_double(0x7ff0000000000000LL);
    const ACC ninf_a = (ACC)ninf;
    const int tiles_xy = prm.tiles_x * prm.tiles_y;
    const int zsub = prm.zsub;
    double gmax = 0.0;
    unsigned long long tbest = ~0ull;
    ACC tmax = ninf_a;
    bool saw_nan = false;
    int n_work = 1, work = 0, work_step = 1;
    const int* list = prm.active + m * (1 + kMaxActive);
    // listed: the work items are (tile, z sub-chunk) entries of the list (all entries if it
    // overflowed) instead of this CTA's own whole tile
    bool listed = false;
    if (PASS == 2) {
        gmax = prm.gmax[m];
        const int n_act = list[0];
        n_work = n_act < 0 ? (int)prm.nent : n_act;
        work = blockIdx.x; work_step = gridDim.x;
        listed = true;
    } else if (PASS == 1 && prm.use_list) {
        // behind the filter the grid is a few CTAs per map: they share the listed entries, or -- when
        // the list overflowed (constant maps) -- stride over every tile
        work = blockIdx.x; work_step = gridDim.x;
        if (list[0] >= 0) { n_work = list[0]; listed = true; }
        else n_work = (int)prm.ntiles;
    }
    const bool fuse = FUSE && PASS == 1 && listed && n_work == 1 && prm.zc_fine <= kFusePlanes;
    const int zc_fine = listed ? prm.zc_fine : prm.zc;
    // z-slide ownership: column cx of the input tile, rows g, g + G, ...
    const int cx = min(tid % MT::CW, MT::W - 1), g = min(tid / MT::CW, MT::G - 1);
    const bool zstore = tid % MT::CW < MT::W && tid / MT::CW < MT::G;
    const bool last_in = g + (MT::NSLOT - 1) * MT::G < MT::R;   // last slot inside the tile?
    // x-pass task and y-pass ownership
    const int task_r = tid % MT::R, task_seg = tid / MT::R;
    const int ox = tid % TX, oy0 = (tid / TX) * MT::RUN;
    const unsigned int D1 = (unsigned int)prm.D1, D2 = (unsigned int)prm.D2, PT = (unsigned int)prm.pitch;
    for (; work < n_work; work += work_step) {
        int tile, fine;
        if (!listed) { tile = prm.use_list ? work : blockIdx.x; fine = 0; }
        else { const int id = list[0] < 0 ? work : list[1 + work]; tile = id / zsub; fine = id - tile * zsub; }
        const int tx_i = tile % prm.tiles_x;
        const int ty_i = (tile / prm.tiles_x) % prm.tiles_y;
        const int zc_i = tile / tiles_xy;
        const int64_t zo0 = (int64_t)zc_i * prm.zc + (int64_t)fine * zc_fine;
        const int64_t zo1 = min(zo0 + zc_fine, min((int64_t)(zc_i + 1) * prm.zc, prm.O0));
        if (zo0 >= zo1) continue;
        // tile columns follow the filter's x tiling when it runs in front (FusedParams::tx_per)
        const unsigned int x0 = (unsigned int)((tx_i / prm.tx_per) * prm.xs_stride + (tx_i % prm.tx_per) * TX);
        const unsigned int y0 = (unsigned int)ty_i * TY;
        const int nplanes = (int)(zo1 - zo0) + p0 - 1;
        const unsigned int plane_elems = D1 * PT;
        const T* src = reinterpret_cast<const T*>(prm.maps) + m * prm.stride_m + zo0 * (int64_t)plane_elems;
        unsigned int off[MT::NSLOT];     // clamped in-plane offsets of this thread's input rows
#pragma unroll
        for (int i = 0; i < MT::NSLOT; ++i)
            off[i] = min(y0 + g + i * MT::G, D1 - 1) * PT + min(x0 + cx, D2 - 1);
        ACC zs[MT::NSLOT];
        T nw[MT::NSLOT], od[MT::NSLOT];
#pragma unroll
        for (int i = 0; i < MT::NSLOT; ++i) { zs[i] = (ACC)0; od[i] = (T)0; }
        unsigned int valid = 0;   // bit k: output (oy0 + k, ox) lies inside the map
#pragma unroll
        for (int k = 0; k < MT::RUN; ++k)
            if (y0 + oy0 + k < prm.O1 && x0 + ox < prm.O2) valid |= 1u << k;
        const bool all_valid = valid == (1u << MT::RUN) - 1;

        // A listed entry is serial latency (one CTA, zc_fine + p0 - 1 planes, one plane of look-ahead): the
        // p0 - 1 planes that only fill the z-window are fetched kWarm planes at a time instead (same adds,
        // same order).
        int zi_first = 0;
        if (listed) {
            constexpr int kWarm = MINB == 1 ? 3 : 2;
            const int warm = p0 - 1;
            for (; zi_first + kWarm <= warm; zi_first += kWarm) {
                T w[kWarm][MT::NSLOT];
#pragma unroll
                for (int k = 0; k < kWarm; ++k) {
                    const T* pk = src + (int64_t)(zi_first + k) * plane_elems;
#pragma unroll
                    for (int i = 0; i < MT::NSLOT; ++i) w[k][i] = In<T>::load_one(pk + off[i]);
                }
#pragma unroll
                for (int k = 0; k < kWarm; ++k) {
#pragma unroll
                    for (int i = 0; i < MT::NSLOT; ++i) zs[i] += (ACC)w[k][i];
                }
            }
        }
        {   // the first plane of the pipelined march enters
            const T* pn = src + (int64_t)zi_first * plane_elems;
#pragma unroll
            for (int i = 0; i < MT::NSLOT; ++i) nw[i] = In<T>::load_one(pn + off[i]);
        }
        // Software pipeline over planes, ONE block barrier per plane: in the same phase the CTA
        // stores the z-window sums of output plane t (A[t & 1]), x-passes plane t-1
        // (A[(t-1) & 1] -> Bs[(t-1) & 1]) and y-passes plane t-2 (Bs[t & 1]); two drain steps.
        const int nout = nplanes - (p0 - 1);
        int sub_planes = 0, sub_idx = 0;      // planes finished in the current z sub-chunk, its index
        for (int zi = zi_first; zi < nplanes + 2; ++zi) {
            const int t = zi - (p0 - 1);
            if (zi < nplanes) {
#pragma unroll
                for (int i = 0; i < MT::NSLOT; ++i) {
                    zs[i] += (ACC)nw[i] - (ACC)od[i];
                }
                if (zi + 1 < nplanes) {    // entering plane zi+1 and leaving plane zi+1-p0: in flight
                    const T* pn = src + (int64_t)(zi + 1) * plane_elems;   // during the x / y passes
#pragma unroll
                    for (int i = 0; i < MT::NSLOT; ++i) nw[i] = In<T>::load_one(pn + off[i]);
                    if (zi + 1 >= p0) {
                        const T* po = pn - (int64_t)p0 * plane_elems;
#pragma unroll
                        for (int i = 0; i < MT::NSLOT; ++i) od[i] = In<T>::load_one(po + off[i]);
                    }
                }
                if (t < 0) continue;                  // z-window still filling: registers only
                if (zstore) {
                    ACC* a = A + (t & 1) * MT::R * MT::pitchA + g * MT::pitchA + cx;
#pragma unroll
                    for (int i = 0; i < MT::NSLOT - 1; ++i) a[i * MT::G * MT::pitchA] = zs[i];
                    if (last_in) a[(MT::NSLOT - 1) * MT::G * MT::pitchA] = zs[MT::NSLOT - 1];
                }
            }
            // ---- x-pass of plane t-1: 16 outputs per task as two independent sliding chains of 8
            // (the sliding sum is a dependent chain; halving it is worth the second tree sum)
            if (t >= 1 && t - 1 < nout && tid < MT::R * MT::NSEG) {
                const ACC* row = A + ((t - 1) & 1) * MT::R * MT::pitchA + task_r * MT::pitchA + task_seg * MT::SEG;
                ACC* dst = Bs + ((t - 1) & 1) * MT::R * MT::pitchB + task_r * MT::pitchB + task_seg * MT::SEG;
                constexpr int H = MT::SEG / 2;
                ACC va[PC], vb2[PC];
#pragma unroll
                for (int k = 0; k < PC; ++k) { va[k] = row[k]; vb2[k] = row[H + k]; }
                ACC sa = tree_sum_t<PC, ACC>(va), sb = tree_sum_t<PC, ACC>(vb2);
                dst[0] = sa; dst[H] = sb;
#pragma unroll
                for (int i = 1; i < H; ++i) {
                    const ACC ia = row[i + PC - 1], ib = row[H + i + PC - 1];
                    sa += ia - va[(i - 1) % PC];
                    sb += ib - vb2[(i - 1) % PC];
                    dst[i] = sa; dst[H + i] = sb;
                }
            }
            // ---- y-pass of plane t-2: RUN outputs down one column, again as two independent chains
            if (t >= 2 && t - 2 < nout) {
                const ACC* cb = Bs + (t & 1) * MT::R * MT::pitchB + oy0 * MT::pitchB + ox;
                ACC o[MT::RUN];
                {
                    constexpr int H = MT::RUN / 2;
                    ACC ca[PC], cb2[PC];
#pragma unroll
                    for (int j = 0; j < PC; ++j) { ca[j] = cb[j * MT::pitchB]; cb2[j] = cb[(H + j) * MT::pitchB]; }
                    o[0] = tree_sum_t<PC, ACC>(ca); o[H] = tree_sum_t<PC, ACC>(cb2);
#pragma unroll
                    for (int k = 1; k < H; ++k) {
                        o[k] = o[k - 1] + (cb[(k + PC - 1) * MT::pitchB] - ca[k - 1]);
                        o[H + k] = o[H + k - 1] + (cb[(H + k + PC - 1) * MT::pitchB] - cb2[k - 1]);
                    }
                }
                if (prm.mean_flag) {
#pragma unroll
                    for (int k = 0; k < MT::RUN; ++k) o[k] = box_mean_div(o[k], prm.denom);
                }
                if constexpr (PASS <= 1) {
                    if (FUSE && fuse) {     // window sums of this plane, -inf where the window is outside the map
#pragma unroll
                        for (int k = 0; k < MT::RUN; ++k)
                            saved[((t - 2) * TY + oy0 + k) * TX + ox] = ((valid >> k) & 1u) ? (double)o[k] : ninf;
                    }
                    if (all_valid) {
#pragma unroll
                        for (int k = 0; k < MT::RUN; ++k) { tmax = acc_max(tmax, o[k]); saw_nan |= o[k] != o[k]; }
                    } else {
#pragma unroll
                        for (int k = 0; k < MT::RUN; ++k)
                            if ((valid >> k) & 1u) { tmax = acc_max(tmax, o[k]); saw_nan |= o[k] != o[k]; }
                    }
                    // one maximum per z sub-chunk of prm.zc_fine output planes: pass 2 re-walks only the
                    // sub-chunks np.isclose to the map maximum, not the whole z-chunk of the tile
                    // (counted: a runtime modulo per plane costs ~25 instructions)
                    if (++sub_planes == prm.zc_fine || t - 1 == nout) {
                        double tm = saw_nan ? __longlong_as_double(0x7ff8000000000000LL) : (double)tmax;
#pragma unroll
                        for (int off = 16; off > 0; off >>= 1) tm = nanmax(tm, __shfl_xor_sync(0xffffffffu, tm, off));
                        if ((tid & 31) == 0) red[tid >> 5] = tm;   // red[] is free: the last read precedes a barrier
                        __syncthreads();
                        if (tid == 0) {
                            double mm = red[0];
                            for (int w = 1; w < NT / 32; ++w) mm = nanmax(mm, red[w]);
                            prm.tile_max[m * prm.nent + (int64_t)tile * zsub + fine + sub_idx] = mm;
                        }
                        tmax = ninf_a; saw_nan = false; sub_planes = 0; ++sub_idx;
                    }
                } else {
                    const int64_t oz = zo0 + (t - 2);
#pragma unroll
                    for (int k = 0; k < MT::RUN; ++k) {
                        if (((valid >> k) & 1u) && np_isclose((double)o[k], gmax, prm.rtol, prm.atol)) {
                            const unsigned long long lin =
                                (unsigned long long)((oz * prm.O1 + (y0 + oy0 + k)) * prm.O2 + (x0 + ox));
                            tbest = lin < tbest ? lin : tbest;
                        }
                    }
                }
            }
            __syncthreads();
        }
        if (PASS <= 1 && !listed && tid == 0) {   // sub-chunks past the end of a short last z-chunk hold nothing
            const int written = (int)((zo1 - zo0 + prm.zc_fine - 1) / prm.zc_fine);
            for (int f = written; f < zsub; ++f) prm.tile_max[m * prm.nent + (int64_t)tile * zsub + f] = ninf;
        }
        if (FUSE && fuse) {
            // the only candidate of this map was just walked: maximum (NaN propagates, as np.max), then
            // the first C-order window np.isclose to it, from the sums kept in shared memory
            const int nsaved = (int)(zo1 - zo0) * TY * TX;
            double g = ninf;
            for (int i = tid; i < nsaved; i += NT) g = nanmax(g, saved[i]);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) g = nanmax(g, __shfl_xor_sync(0xffffffffu, g, off));
            if ((tid & 31) == 0) red[tid >> 5] = g;
            __syncthreads();
            g = red[0];
            for (int w = 1; w < NT / 32; ++w) g = nanmax(g, red[w]);
            unsigned long long best = ~0ull;
            for (int i = tid; i < nsaved; i += NT) {
                if (np_isclose(saved[i], g, prm.rtol, prm.atol)) {
                    const int pz = i / (TY * TX), py = (i / TX) % TY, px = i % TX;
                    const unsigned long long lin =
                        (unsigned long long)(((zo0 + pz) * prm.O1 + (y0 + py)) * prm.O2 + (x0 + px));
                    best = lin < best ? lin : best;
                }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, off);
                best = other < best ? other : best;
            }
            if ((tid & 31) == 0) red_idx[tid >> 5] = best;
            __syncthreads();
            if (tid == 0) {
                for (int w = 1; w < NT / 32; ++w) best = red_idx[w] < best ? red_idx[w] : best;
                prm.gmax[m] = g;
                prm.out.put_score(m, g);
                prm.out.put_corner(m, best, prm.O1, prm.O2);
                atomicExch(prm.tickets + 4 * m + 2, kFusedDone);     // ordered before this CTA's pass-1 ticket
            }
        }
    }
    box_pass_finish<PASS, kFusedThreads, true>(prm, m, 0.0, tbest, red, s_flag, s_count);
}

// ------------------------------------------------------------------ K2b filter: strip kernel (TMA)
// The fp32 FILTER in front of the exact march: fp32 box sums over every window, one maximum per
// (march tile, z sub-chunk) entry plus a bound on max |input|; box_pass_finish<0> turns them into
// the work list of the exact fp64 passes (rigorous error bound, filter_err_coef).  fp32 sliding
// sums are not the answer -- np.isclose decisions need fp64 -- but they say where the answer can be.
//
// Built around the L1 / shared-memory data pipe, which bounded the previous filter (ncu, r01g: 64 %
// of its wavefronts, one block barrier per plane): here a WARP owns a strip of K output rows over
// the whole staged width and keeps everything between the plane tile and the maximum in registers.
//  * planes arrive by TMA (cp.async.bulk.tensor, one 4-D box [1 map, 1 plane, RIN rows, WF columns]
//    for the entering plane and one for the leaving plane, out-of-range rows / columns zero-filled)
//    into a two-stage shared-memory ring; one producer warp, full / empty mbarriers, no block barrier
//    anywhere in the march;
//  * z-stage: lane = one float4 column of the strip's K + 9 input rows: LDS.128 of the entering and
//    the leaving plane, packed f32x2 adds into K + 9 float4 window sums (registers);
//  * y-stage: tree sum of 10 rows, then K - 1 slides, all in registers (no exchange);
//  * x-stage: the 10-wide window of output x = 4 lane + j spans lanes l .. l+3: five SHFL per float4
//    (lane sums Q, pair sum A, three single elements), one tree-shaped start S0 = (Q_l + Q_{l+1}) + A_{l+2}
//    and three slides;
//  * the strip's maximum per z sub-chunk goes through one shared-memory atomicMax per half-warp; the
//    last warp of the CTA to close a sub-chunk writes the entries.
// Per output: 11 thread instructions and 0.19 data-pipe wavefronts (previous filter: 24 and 0.39).
// LW = 32: staged rows of 128 floats, 119 outputs per x tile (two march tile columns: lanes 0-15 and
// 16-31); LW = 16 (maps up to 64 wide): rows of 64 floats, two strips side by side in a warp.
template <int K_, int NW_, int LW_, int NSTAGE_ = 2, int MINB_ = 2> struct StripTile {
    static constexpr int K = K_, NW = NW_, LW = LW_, PC = 10;
    static constexpr int SUBS = 32 / LW;                 // strips side by side in one warp
    static constexpr int ROWS = NW * SUBS * K;           // output rows per CTA
    static constexpr int RIN = ROWS + PC - 1;            // input rows per staged plane
    static constexpr int RW = K + PC - 1;                // input rows per strip
    static constexpr int WF = LW * 4;                    // input columns per staged plane
    static constexpr int XS = WF - PC + 1;               // outputs per x tile (119 / 55)
    static constexpr int XSTEP = XS & ~3;                // x tile pitch: the innermost TMA coordinate must be a
                                                         // multiple of 16 bytes (an odd origin traps as an illegal
                                                         // instruction), so neighbouring tiles share 3 outputs
    static constexpr int NSTAGE = NSTAGE_, MINB = MINB_;
    static constexpr int PLANE = RIN * WF;               // floats per staged plane
    static constexpr int NT = (NW + 1) * 32;             // consumer warps + the producer warp
    static constexpr int TYC = ROWS / 32;                // march tile rows per CTA
    static constexpr int TXC = (XS + 63) / 64;           // march tile columns per x tile (2 / 1)
    static constexpr int NSLOT = TYC * TXC;
    static constexpr size_t smem = (size_t)NSTAGE * 2 * PLANE * sizeof(float) + 128;   // + alignment slack
    static_assert(ROWS % 32 == 0 && 32 % K == 0, "strips must not straddle march tiles");
    static_assert((PLANE * sizeof(float)) % 128 == 0, "TMA destinations are 128-byte aligned");
    static_assert(RIN <= 256 && WF <= 256, "TMA box limits");
};

__device__ __forceinline__ int strip_fkey(float f) {      // order-preserving float -> int
    const int b = __float_as_int(f);
    return b >= 0 ? b : b ^ 0x7fffffff;
}
__device__ __forceinline__ float strip_funkey(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }
constexpr int kStripKeyNinf = (int)0x807fffff;            // strip_fkey(-inf)

__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.f, -1.f), a); }
__device__ __forceinline__ float4 add4(float4 a, float4 b) {
    const float2 lo = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
    const float2 hi = __fadd2_rn(make_float2(a.z, a.w), make_float2(b.z, b.w));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ float4 sub4(float4 a, float4 b) {
    const float2 lo = sub2(make_float2(a.x, a.y), make_float2(b.x, b.y));
    const float2 hi = sub2(make_float2(a.z, a.w), make_float2(b.z, b.w));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}
// running max |v| with NaN propagation: one FMNMX.NAN.XORSIGN per element (the sign of the result is
// meaningless; the caller takes the magnitude).  An OR of the magnitude bits would be cheaper but is not a
// bound: the exponent fields of 1.x and 2.x OR to all ones.
__device__ __forceinline__ float absmax4(float acc, float4 v) {
    asm("max.NaN.xorsign.abs.f32 %0, %0, %1;" : "+f"(acc) : "f"(v.x));
    asm("max.NaN.xorsign.abs.f32 %0, %0, %1;" : "+f"(acc) : "f"(v.y));
    asm("max.NaN.xorsign.abs.f32 %0, %0, %1;" : "+f"(acc) : "f"(v.z));
    asm("max.NaN.xorsign.abs.f32 %0, %0, %1;" : "+f"(acc) : "f"(v.w));
    return acc;
}
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
    float4 q;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "r"(a));
    return q;
}
template <int N> __device__ __forceinline__ float4 tree_sum4(const float4* v) {
    if constexpr (N == 1) return v[0];
    else return add4(tree_sum4<N / 2>(v), tree_sum4<N - N / 2>(v + N / 2));
}

template <typename ST>
__global__ void __launch_bounds__(ST::NT, ST::MINB)
box_strip_filter_kernel(const __grid_constant__ CUtensorMap tmap, const FusedParams prm, int cta_y, int nxs, int ysteps,
                        unsigned int ctas_per_map) {
    constexpr int K = ST::K, NW = ST::NW, LW = ST::LW;
    constexpr int NT = ST::NT, RW = ST::RW, PC = ST::PC;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[ST::NSTAGE];
    __shared__ __align__(8) uint64_t empty_bar[ST::NSTAGE];
    __shared__ int s_key[4][ST::NSLOT];     // sub-chunk maxima in flight (order-preserving keys), 4 deep
    __shared__ int s_cnt[4];
    __shared__ double red[NT / 32];
    __shared__ int s_flag;
    __shared__ int s_count;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int p0 = prm.p0;
    // Two march directions.  z-march (ysteps == 0): work item = (map, z-chunk, CTA row of march tiles,
    // x tile), x fastest so that neighbours share halo rows in L2; the ring streams the planes of the
    // z-chunk.  y-march (single-plane maps, i.e. 2-D images; ysteps > 0): work item = (map, chunk of
    // ysteps CTA rows, x tile) and the ring streams the CTA rows of the chunk top to bottom -- a
    // one-plane tile would otherwise be load, compute, exit with nothing in flight.
    const bool ym = ysteps > 0;
    int item = blockIdx.x;
    const int xt = item % nxs; item /= nxs;
    int cy, zc_i, nsteps, nout;
    int64_t m, zo0;
    if (ym) {
        const int ychunks = (cta_y + ysteps - 1) / ysteps;
        cy = (item % ychunks) * ysteps; m = item / ychunks;
        zc_i = 0; zo0 = 0; nout = 1;
        nsteps = min(ysteps, cta_y - cy);
    } else {
        cy = item % cta_y; item /= cta_y;
        zc_i = item % prm.chunks_z; m = item / prm.chunks_z;
        zo0 = (int64_t)zc_i * prm.zc;
        nout = (int)(min(zo0 + prm.zc, prm.O0) - zo0);
        nsteps = nout + p0 - 1;
    }
    const int x0 = xt * ST::XSTEP, y0 = cy * ST::ROWS;
    const uint32_t stage_base = (smem_u32(smem_raw) + 127u) & ~127u;
    constexpr uint32_t kPlaneBytes = ST::PLANE * sizeof(float);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < ST::NSTAGE; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, NW * kArriveLanes); }
        for (int q = 0; q < 4; ++q) {
            s_cnt[q] = 0;
            for (int i = 0; i < ST::NSLOT; ++i) s_key[q][i] = kStripKeyNinf;
        }
        mbar_fence_init();
    }
    __syncthreads();
    float amax = 0.f;
    if (warp == NW) {
        // ---- producer: one lane streams the plane tiles through the ring.  An entering plane is read
        // again as the leaving plane p0 steps later: it is marked evict-last in L2 when it enters and
        // evict-first when it leaves (ncu: 1.23 -> 0.93 GB of DRAM reads for 0.81 GB of maps).
        if (lane == 0) {
            tma_prefetch_desc(&tmap);
            uint64_t pol_last, pol_first;
            asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
            for (int zi = 0; zi < nsteps; ++zi) {
                const int stage = zi % ST::NSTAGE;
                if (zi >= ST::NSTAGE) mbar_wait(empty_bar + stage, ((zi / ST::NSTAGE) & 1) ^ 1u);
                const uint32_t dst = stage_base + stage * 2 * kPlaneBytes;
                if (ym) {
                    mbar_expect_tx(full_bar + stage, kPlaneBytes);
                    tma_load_4d_addr(dst, &tmap, x0, y0 + zi * ST::ROWS, 0, (int)m, full_bar + stage);
                } else {
                    const bool leaving = zi >= p0;
                    mbar_expect_tx(full_bar + stage, leaving ? 2 * kPlaneBytes : kPlaneBytes);
                    tma_load_4d_addr_hint(dst, &tmap, x0, y0, (int)(zo0 + zi), (int)m, full_bar + stage, pol_last);
                    if (leaving)
                        tma_load_4d_addr_hint(dst + kPlaneBytes, &tmap, x0, y0, (int)(zo0 + zi - p0), (int)m,
                                              full_bar + stage, pol_first);
                }
            }
        }
    } else {
        // ---- consumers: lane = float4 column c4 of strip `strip` (K output rows, K + 9 input rows)
        const int c4 = lane % LW, strip = warp * ST::SUBS + lane / LW;
        const int r0 = strip * K;
        const int xo = 4 * c4, xlim = min(ST::XS, (int)prm.O2 - x0);        // valid outputs: xo + j < xlim
        const bool cv0 = xo < xlim, cv1 = xo + 1 < xlim, cv2 = xo + 2 < xlim, cv3 = xo + 3 < xlim;
        // every strip's own K rows feed the bound on max |input| (it is reduced over the CTAs of a map, and
        // the tail rows of a strip are the own rows of the next strip / CTA); only the last strip of the
        // bottom CTA row reads tail rows that are nobody's own
        const bool or_tail = strip == NW * ST::SUBS - 1;
        const uint32_t lane_off = (uint32_t)(r0 * ST::WF + xo) * 4u;
        // entry slot of this half-warp: march tile row (r0 / 32) x tile column (lanes 16.. of a 128-wide row)
        const int slot = (r0 / 32) * ST::TXC + (LW == 32 ? lane / 16 : 0);
        const float ninf_f = -__int_as_float(0x7f800000);
        float4 zs[RW];
#pragma unroll
        for (int r = 0; r < RW; ++r) zs[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        float tm0 = ninf_f, tm1 = ninf_f, tm2 = ninf_f, tm3 = ninf_f;   // four independent max chains
        int sub_idx = 0, sub_planes = 0, closes = 0;
        for (int zi = 0; zi < nsteps; ++zi) {
            const int stage = zi % ST::NSTAGE;
            mbar_wait(full_bar + stage, (zi / ST::NSTAGE) & 1);
            const uint32_t pn = stage_base + stage * 2 * kPlaneBytes + lane_off;
            if (ym) {
#pragma unroll
                for (int r = 0; r < RW; ++r) {
                    zs[r] = lds_f4(pn + r * ST::WF * 4);
                    if (r < K || or_tail) amax = absmax4(amax, zs[r]);
                }
            } else if (zi >= p0) {
#pragma unroll
                for (int r = 0; r < RW; ++r) {
                    const float4 nw = lds_f4(pn + r * ST::WF * 4);
                    const float4 od = lds_f4(pn + kPlaneBytes + r * ST::WF * 4);
                    zs[r] = add4(zs[r], sub4(nw, od));
                    if (r < K || or_tail) amax = absmax4(amax, nw);
                }
            } else {
#pragma unroll
                for (int r = 0; r < RW; ++r) {
                    const float4 nw = lds_f4(pn + r * ST::WF * 4);
                    zs[r] = add4(zs[r], nw);
                    if (r < K || or_tail) amax = absmax4(amax, nw);
                }
            }
            // the stage is free once every lane's loads have landed in registers: the arrive is issued
            // behind the last window sum (which needs them all) and a warp barrier
            asm volatile("" ::"f"(zs[0].x), "f"(zs[RW - 1].w) : "memory");
            __syncwarp();
            if (arrives(lane)) mbar_arrive(empty_bar + stage);
            if (!ym && zi < p0 - 1) continue;                   // z-window still filling
            // ---- y-stage (registers) and x-stage (shuffles) of this output plane / CTA row
            const int cy_cur = ym ? cy + zi : cy;
            const int nvr = max(0, min(K, (int)prm.O1 - (cy_cur * ST::ROWS + r0)));   // valid output rows of the strip
            const int nvr_w = __shfl_sync(0xffffffffu, nvr, 0);         // ... of the warp's first strip (the larger)
            float4 o = tree_sum4<PC>(zs);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if (k > 0) o = add4(o, sub4(zs[k + PC - 1], zs[k - 1]));
                if (k < nvr_w) {                                // warp-uniform (the shuffles need every lane)
                    const bool rv = LW == 32 || k < nvr;
                    const float A = o.x + o.y, B = o.z + o.w, Q = A + B;
                    const float Q1 = __shfl_down_sync(0xffffffffu, Q, 1);
                    const float A2 = __shfl_down_sync(0xffffffffu, A, 2);
                    const float e2z = __shfl_down_sync(0xffffffffu, o.z, 2);
                    const float e2w = __shfl_down_sync(0xffffffffu, o.w, 2);
                    const float e3x = __shfl_down_sync(0xffffffffu, o.x, 3);
                    const float S0 = (Q + Q1) + A2;
                    const float S1 = S0 + (e2z - o.x);
                    const float S2 = S1 + (e2w - o.y);
                    const float S3 = S2 + (e3x - o.z);
                    if (cv0 && rv) tm0 = fmaxf(tm0, S0);
                    if (cv1 && rv) tm1 = fmaxf(tm1, S1);
                    if (cv2 && rv) tm2 = fmaxf(tm2, S2);
                    if (cv3 && rv) tm3 = fmaxf(tm3, S3);
                }
            }
            // ---- one maximum per z sub-chunk of prm.zc_fine output planes (counted, not computed); in the
            // y-march every step closes the single entry of its own march tile row
            if (ym || ++sub_planes == prm.zc_fine || zi == nsteps - 1) {
                float mx = fmaxf(fmaxf(tm0, tm1), fmaxf(tm2, tm3));
#pragma unroll
                for (int o2 = 8; o2 > 0; o2 >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o2));
                const int q = closes & 3;
                if ((lane & 15) == 0) atomicMax(&s_key[q][slot], strip_fkey(mx));
                __syncwarp();
                if (lane == 0) {
                    __threadfence_block();
                    if (atomicAdd(&s_cnt[q], 1) == NW - 1) {    // last warp of the CTA to close this sub-chunk
                        __threadfence_block();
#pragma unroll
                        for (int i = 0; i < ST::NSLOT; ++i) {
                            const int key = atomicExch(&s_key[q][i], kStripKeyNinf);
                            const int ty_i = cy_cur * ST::TYC + i / ST::TXC, tx_i = xt * ST::TXC + i % ST::TXC;
                            if (ty_i < prm.tiles_y) {
                                const int64_t tile = ((int64_t)zc_i * prm.tiles_y + ty_i) * prm.tiles_x + tx_i;
                                prm.tile_max[m * prm.nent + tile * prm.zsub + sub_idx] = (double)strip_funkey(key);
                            }
                        }
                        atomicExch(&s_cnt[q], 0);
                    }
                }
                tm0 = tm1 = tm2 = tm3 = ninf_f; sub_planes = 0; ++closes;
                if (!ym) ++sub_idx;
            }
        }
    }
    if (tid == 0 && !ym) {   // sub-chunks past the end of a short last z-chunk hold nothing
        const double ninf = -__longlong_as_double(0x7ff0000000000000LL);
        const int written = (nout + prm.zc_fine - 1) / prm.zc_fine;
        for (int i = 0; i < ST::NSLOT; ++i) {
            const int ty_i = cy * ST::TYC + i / ST::TXC, tx_i = xt * ST::TXC + i % ST::TXC;
            if (ty_i >= prm.tiles_y) continue;
            const int64_t tile = ((int64_t)zc_i * prm.tiles_y + ty_i) * prm.tiles_x + tx_i;
            for (int f = written; f < prm.zsub; ++f) prm.tile_max[m * prm.nent + tile * prm.zsub + f] = ninf;
        }
    }
    __syncthreads();
    if (tid == 0) {   // the march is over for every warp: the ring's barriers are idle (common.cuh, mbar_inval)
#pragma unroll
        for (int s = 0; s < ST::NSTAGE; ++s) { mbar_inval(full_bar + s); mbar_inval(empty_bar + s); }
    }
    box_pass_finish<0, NT, true>(prm, m, 0.0, ~0ull, red, s_flag, s_count, fabsf(amax),
                                 ctas_per_map);
}

struct PatchPlan {
    int64_t O0, O1, O2;
    int tiles_x, tiles_y, chunks_z, zc;
    int64_t ntiles;
    size_t smem;
};

static int make_patch_plan(const int64_t* shape, const int64_t* patch, PatchPlan& pl) {
    for (int d = 0; d < 3; ++d) {
        if (shape[d] <= 0 || patch[d] <= 0)
            return set_error(VALUES_ERR_INVALID_ARG, "patch_max: non-positive shape/patch");
        if (patch[d] > shape[d])
            return set_error(VALUES_ERR_INVALID_ARG,
                             "For 'valid' mode, one must be at least as large as the other in "
                             "every dimension (axis %d: image %lld < patch %lld)",
                             d, (long long)shape[d], (long long)patch[d]);
    }
    pl.O0 = shape[0] - patch[0] + 1; pl.O1 = shape[1] - patch[1] + 1; pl.O2 = shape[2] - patch[2] + 1;
    pl.tiles_x = (int)ceil_div(pl.O2, kTX);
    pl.tiles_y = (int)ceil_div(pl.O1, kTY);
    pl.zc = 32;
    pl.chunks_z = (int)ceil_div(pl.O0, pl.zc);
    pl.ntiles = (int64_t)pl.tiles_x * pl.tiles_y * pl.chunks_z;
    const int64_t R = kTY + patch[1] - 1, W = kTX + patch[2] - 1;
    pl.smem = (size_t)(R * (W | 1) + R * kTX + patch[0] * kTX * kTY) * sizeof(double);
    if (pl.smem > 227 * 1024)
        return set_error(VALUES_ERR_UNSUPPORTED,
                         "patch_max: patch (%lld,%lld,%lld) needs %zu B of shared memory (> 227 KB)",
                         (long long)patch[0], (long long)patch[1], (long long)patch[2], pl.smem);
    if (pl.ntiles > 0x7fffffffLL) return set_error(VALUES_ERR_UNSUPPORTED, "patch_max: too many tiles");
    return VALUES_OK;
}

template <typename T>
static int run_patch(PatchParams prm, const PatchPlan& pl, int64_t M, double* gmax,
                     const PatchOut& out, cudaStream_t st) {
    auto k1 = patch_kernel<T, 1>;
    auto k2 = patch_kernel<T, 2>;
    if (pl.smem > 48 * 1024) {
        if (cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem) != cudaSuccess ||
            cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem) != cudaSuccess)
            return set_error(VALUES_ERR_CUDA, "patch_max: cudaFuncSetAttribute failed");
    }
    for (int64_t m0 = 0; m0 < M; m0 += 65535) {  // gridDim.y limit
        const int64_t mc = std::min<int64_t>(65535, M - m0);
        PatchParams q = prm;
        q.maps = reinterpret_cast<const T*>(prm.maps) + m0 * prm.stride_m;
        q.tile_max = prm.tile_max + m0 * pl.ntiles;
        q.gmax = gmax + m0;
        q.best = prm.best + m0;
        const dim3 grid((unsigned)pl.ntiles, (unsigned)mc);
        k1<<<grid, kThreads, pl.smem, st>>>(q);
        int rc = check_launch("patch_kernel<1>");
        if (rc) return rc;
        patch_select_kernel<<<(unsigned)mc, kThreads, 0, st>>>(q.tile_max, pl.ntiles, gmax + m0,
                                                                out.offset(m0), q.best);
        if ((rc = check_launch("patch_select_kernel"))) return rc;
        k2<<<grid, kThreads, pl.smem, st>>>(q);
        if ((rc = check_launch("patch_kernel<2>"))) return rc;
    }
    patch_finish_kernel<<<(unsigned)ceil_div(M, 128), 128, 0, st>>>(prm.best, M, pl.O1, pl.O2, out);
    return check_launch("patch_finish_kernel");
}

// ------------------------------------------------------------------ fused / march / strip paths: host side
struct FusedPlan {
    int march;             // 1 = box_march_kernel (32x64 tiles, z first), 0 = box_fused_kernel
    int strip_lw;          // > 0: the fp32 strip filter runs in front of the march (lanes per staged row: 32 / 16)
    int ty, tx;            // 32x64 (march), 16x64 or 8x32 (fused); 0 = neither applies
    int64_t O0, O1, O2;
    int tiles_x, tiles_y, chunks_z, zc, zsub, zc_fine;
    int tx_per, xs_stride; // x origin of march tile column tx: (tx / tx_per) * xs_stride + (tx % tx_per) * 64
    int cta_y, nxs;        // strip filter grid: CTA rows of march tiles, x tiles
    int ysteps;            // > 0: single-plane maps, the filter marches over ysteps CTA rows per CTA (y-march)
    int64_t ntiles;
};

static size_t fused_smem_bytes(int ty, int tx, const int64_t* patch) {
    const int64_t R = ty + patch[1] - 1, W = tx + patch[2] - 1;
    return (size_t)(R * (tx + 1) + patch[0] * ty * tx + R * (W | 1)) * sizeof(double);
}

// strip filter shapes: <K output rows per strip, consumer warps, lanes per staged row>
using StripWide = StripTile<8, 4, 32>;     // 128-wide staged rows: 32 output rows x 119 outputs per CTA
using StripNarrow = StripTile<8, 4, 16, 2, 3>;   // maps up to 64 wide: 64 output rows x 55 outputs per CTA, three CTAs per SM

static int check_patch_shape(const int64_t* shape, const int64_t* patch) {
    for (int d = 0; d < 3; ++d) {
        if (shape[d] <= 0 || patch[d] <= 0)
            return set_error(VALUES_ERR_INVALID_ARG, "patch_max: non-positive shape/patch");
        if (patch[d] > shape[d])
            return set_error(VALUES_ERR_INVALID_ARG,
                             "For 'valid' mode, one must be at least as large as the other in "
                             "every dimension (axis %d: image %lld < patch %lld)",
                             d, (long long)shape[d], (long long)patch[d]);
    }
    return VALUES_OK;
}

static bool march_applies(int path, const int64_t* shape, const int64_t* patch) {
    return (path == 0 || path == 5) && patch[1] == 10 && patch[2] == 10 && shape[2] - 9 > 32 && shape[1] - 9 > 16 &&
           shape[1] * shape[2] < 0x7fffffffLL;
}

// 0 on success (pl.ty == 0 when neither the march nor the fused kernel applies).
// strip: plan the fp32 strip filter in front of the march (the caller has checked dtype / alignment).
static int make_fused_plan(int64_t M, const int64_t* shape, const int64_t* patch, int path, bool strip, FusedPlan& pl) {
    pl = FusedPlan{};
    int rc = check_patch_shape(shape, patch);
    if (rc) return rc;
    pl.O0 = shape[0] - patch[0] + 1; pl.O1 = shape[1] - patch[1] + 1; pl.O2 = shape[2] - patch[2] + 1;
    pl.tx_per = 1; pl.xs_stride = 64;
    const size_t budget = 113 * 1024;  // two CTAs per SM
    if (march_applies(path, shape, patch)) {
        pl.march = 1; pl.ty = 32; pl.tx = 64;
        if (strip && path == 0) pl.strip_lw = shape[2] <= StripNarrow::WF ? 16 : 32;
    }
    else if (path != 0 && path != 4) return VALUES_OK;
    else if (fused_smem_bytes(16, 64, patch) <= budget && pl.O2 > 32) { pl.ty = 16; pl.tx = 64; }
    else if (fused_smem_bytes(8, 32, patch) <= budget) { pl.ty = 8; pl.tx = 32; }
    else return VALUES_OK;
    pl.tiles_y = (int)ceil_div(pl.O1, pl.ty);
    int64_t units;        // CTAs of the heaviest pass per z-chunk
    int slots = 2 * 148;  // CTAs resident at once
    if (pl.strip_lw == 32) {
        pl.nxs = pl.O2 <= StripWide::XS ? 1 : (int)ceil_div(pl.O2 - StripWide::XS, StripWide::XSTEP) + 1;
        pl.tx_per = StripWide::TXC; pl.xs_stride = StripWide::XSTEP;
        pl.tiles_x = pl.nxs * StripWide::TXC; pl.cta_y = (int)ceil_div(pl.tiles_y, StripWide::TYC);
        units = (int64_t)pl.nxs * pl.cta_y * std::max<int64_t>(M, 1);
        if (shape[0] == 1) {   // 2-D image: y-march; chunks of CTA rows chosen by waves x (steps + ring fill)
            const int64_t cols = (int64_t)pl.nxs * std::max<int64_t>(M, 1);
            int64_t best = -1;
            for (int yc = 1; yc <= pl.cta_y; ++yc) {
                const int64_t steps = ceil_div(pl.cta_y, yc);
                const int64_t cost = ceil_div(cols * ceil_div(pl.cta_y, steps), slots) * (steps + 2);
                if (best < 0 || cost < best) { best = cost; pl.ysteps = (int)steps; }
            }
        }
    } else if (pl.strip_lw == 16) {
        pl.nxs = 1; pl.tiles_x = 1; pl.cta_y = (int)ceil_div(pl.tiles_y, StripNarrow::TYC);
        units = (int64_t)pl.cta_y * std::max<int64_t>(M, 1);
        slots = 3 * 148;
    } else {
        pl.tiles_x = (int)ceil_div(pl.O2, pl.tx);
        units = (int64_t)pl.tiles_x * pl.tiles_y * std::max<int64_t>(M, 1);
    }
    // output planes per z-chunk: minimise (planes marched per SM by the heaviest pass) + (planes of
    // the serial re-walk of one sub-chunk by the listed passes); even splits of O0
    int64_t best_cost = -1;
    pl.zc = (int)pl.O0;
    for (int64_t n = 1; n <= 16; ++n) {
        const int64_t zc = ceil_div(pl.O0, n);
        const int64_t ctas = units * ceil_div(pl.O0, zc);
        // planes of pure warm-up per chunk, in units of a full plane (march / strip: register adds only)
        const int64_t warm = pl.strip_lw ? (patch[0] - 1) / 3 + 2 : pl.march ? (patch[0] - 1) / 4 + 2 : patch[0] - 1 + 2;
        // CTAs neither run in lock-step waves nor perfectly smoothly: average both models
        const int64_t smooth = std::max(ceil_div(ctas * (zc + warm), slots), zc + warm);
        const int64_t waves = ceil_div(ctas, slots) * (zc + warm);
        const int64_t pass2 = ceil_div(zc, pl.march ? 32 : 8) + warm;
        const int64_t cost = (smooth + waves) / 2 + pass2;
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; pl.zc = (int)zc; }
        if (zc <= 4) break;
    }
    // the listed passes (exact maxima of the filter's candidates, first index) walk single sub-chunks:
    // serial latency, so the march kernel's sub-chunks are short
    pl.zsub = (int)std::min<int64_t>(pl.march ? 32 : 8, pl.zc);
    pl.zc_fine = (int)ceil_div(pl.zc, pl.zsub);
    pl.chunks_z = (int)ceil_div(pl.O0, pl.zc);
    pl.ntiles = (int64_t)pl.tiles_x * pl.tiles_y * pl.chunks_z;
    if (pl.ntiles * pl.zsub > 0x7fffffffLL) return set_error(VALUES_ERR_UNSUPPORTED, "patch_max: too many tiles");
    return VALUES_OK;
}

template <typename T, int TY, int TX, int PC>
static int run_patch_fused_pc(FusedParams prm, const FusedPlan& pl, int64_t M, cudaStream_t st) {
    auto k1 = box_fused_kernel<T, TY, TX, PC, 1>;
    auto k2 = box_fused_kernel<T, TY, TX, PC, 2>;
    const int64_t patch[3] = {prm.p0, prm.p1, prm.p2};
    const size_t smem = fused_smem_bytes(TY, TX, patch);
    if (cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return set_error(VALUES_ERR_CUDA, "patch_max: cudaFuncSetAttribute(%zu) failed", smem);
    for (int64_t m0 = 0; m0 < M; m0 += 65535) {  // gridDim.y limit
        const int64_t mc = std::min<int64_t>(65535, M - m0);
        FusedParams q = prm;
        q.maps = reinterpret_cast<const T*>(prm.maps) + m0 * prm.stride_m;
        q.tile_max = prm.tile_max + m0 * prm.nent;
        q.best = prm.best + m0;
        q.out = prm.out.offset(m0);
        q.gmax = prm.gmax + m0; q.active = prm.active + m0 * (1 + kMaxActive);
        q.tickets = prm.tickets + 4 * m0;
        if (cudaMemsetAsync(q.tickets, 0, (size_t)mc * 4 * sizeof(unsigned int), st) != cudaSuccess)
            return set_error(VALUES_ERR_CUDA, "patch_max: cudaMemsetAsync failed");
        k1<<<dim3((unsigned)pl.ntiles, (unsigned)mc), kFusedThreads, smem, st>>>(q);
        int rc = check_launch("box_fused_kernel<1>");
        if (rc) return rc;
        // pass 2: a few CTAs per map walk the (normally one-entry) work list
        const unsigned g2 = (unsigned)std::min<int64_t>(pl.ntiles * pl.zsub, 4 * pl.zsub);
        k2<<<dim3(g2, (unsigned)mc), kFusedThreads, smem, st>>>(q);
        if ((rc = check_launch("box_fused_kernel<2>"))) return rc;
    }
    return VALUES_OK;
}

// err_coef of the fp32 filter: a bound on |fp32 box sum - exact box sum| / max |input| for the strip
// kernel's operation order (u = 2^-24, every fp32 add / subtract rounds to nearest, no underflow
// error in additions).  With a = max |input|, Z = p0 a (bound on a z-window sum), pc = 10:
//   z-slide, K = zc + p0 - 1 steps of zs = fl(zs + fl(new - old)):  ez <= K (p0 + 2) u a
//   y-stage, tree over pc (depth d) + at most ys slides s = fl(s + fl(in - out)):
//                                                                   ey <= pc ez + [d pc + ys (pc + 2)] u Z
//   x-stage, tree-shaped start over pc (depth d) + at most xs slides: ex <= pc ey + [d pc + xs (pc + 2)] u pc Z
// doubled to cover the second-order terms and the exact pass's own fp64 rounding.  (The kernel's a is
// the running maximum of |input| over the rows the strips own, NaN-propagating.)
static double filter_err_coef(int zc, int p0, int pc, int s1, int s2) {
    const double u = 5.9604644775390625e-8;
    const double K = zc + p0 - 1, d = tree_depth(pc);
    const double ez = K * (p0 + 2);
    const double e1 = pc * ez + (d * pc + s1 * (pc + 2)) * p0;
    const double e2 = pc * e1 + (d * pc + s2 * (pc + 2)) * (double)pc * p0;
    return 2.0 * u * e2;
}

// fp32 maps [M][D0][D1][D2] as a 4-D tensor map with a [1][1][rows][cols] box
static int make_map_tensor(const FusedParams& prm, int64_t M, int rows, int cols, CUtensorMap* tm) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return set_error(VALUES_ERR_CUDA, "patch_max: cuTensorMapEncodeTiled is not available");
    const cuuint64_t gdim[4] = {(cuuint64_t)prm.D2, (cuuint64_t)prm.D1, (cuuint64_t)prm.D0, (cuuint64_t)M};
    const cuuint64_t gstr[3] = {(cuuint64_t)prm.pitch * 4, (cuuint64_t)prm.D1 * prm.pitch * 4, (cuuint64_t)prm.stride_m * 4};
    const cuuint32_t box[4] = {(cuuint32_t)cols, (cuuint32_t)rows, 1, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(prm.maps), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(VALUES_ERR_CUDA, "patch_max: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return VALUES_OK;
}

// can the strip filter stream these maps by TMA?  (16-byte aligned rows, planes and maps)
// Maps whose rows are not 16-byte aligned (an innermost extent that is not a multiple of 4: 478-wide images,
// 127^3 volumes) cannot be described by a tensor map, and the exact march on every window is 3x the filter
// path.  They are copied once into a scratch with the rows pitched to a multiple of 4 floats (a warp per row,
// lane-consecutive loads and stores; the pad is never read as data: the tensor map's extent stays D2 and the
// march clamps to D2 - 1), and filter + listed passes run on the copy: 2 x 4V bytes for the copy against
// ~8V saved.
__global__ void __launch_bounds__(kThreads) pitch_rows_kernel(const float* __restrict__ maps, int64_t stride_m,
                                                              int64_t rows_per_map, int64_t n_rows, int D2, int pitch,
                                                              float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * kThreads + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * kThreads) >> 5;
    for (int64_t r = warp; r < n_rows; r += nwarps) {
        const int64_t m = r / rows_per_map;
        const float* src = maps + m * stride_m + (r - m * rows_per_map) * D2;
        float* dst = out + r * pitch;
        for (int x = lane; x < pitch; x += 32) dst[x] = x < D2 ? __ldg(src + x) : 0.f;
    }
}
static int64_t pitched_width(int64_t d2) { return (d2 + 3) / 4 * 4; }
static size_t pitched_scratch_bytes(int64_t M, const int64_t* shape) {
    return (size_t)M * (size_t)shape[0] * (size_t)shape[1] * (size_t)pitched_width(shape[2]) * sizeof(float) + 256;
}

static bool strip_filter_ok(const void* maps, int dtype, int64_t stride_m, const int64_t* shape) {
    return dtype == VALUES_F32 && shape[2] % 4 == 0 && stride_m % 4 == 0 &&
           (reinterpret_cast<uintptr_t>(maps) & 15) == 0 && shape[0] < 0x7fffffffLL && encode_tiled_fn() != nullptr;
}

template <typename ST>
static int launch_strip_filter(const FusedParams& prm, const FusedPlan& pl, int64_t M, cudaStream_t st) {
    auto k0 = box_strip_filter_kernel<ST>;
    if (cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ST::smem) != cudaSuccess)
        return set_error(VALUES_ERR_CUDA, "patch_max: cudaFuncSetAttribute(%zu) failed", ST::smem);
    CUtensorMap tm;
    int rc = make_map_tensor(prm, M, ST::RIN, ST::WF, &tm);
    if (rc) return rc;
    const int64_t per_map = pl.ysteps ? ceil_div(pl.cta_y, pl.ysteps) * pl.nxs : (int64_t)pl.chunks_z * pl.cta_y * pl.nxs;
    if (M * per_map > 0x7fffffffLL) return set_error(VALUES_ERR_UNSUPPORTED, "patch_max: grid too large");
    k0<<<(unsigned)(M * per_map), ST::NT, ST::smem, st>>>(tm, prm, pl.cta_y, pl.nxs, pl.ysteps, (unsigned)per_map);
    return check_launch("box_strip_filter_kernel");
}

template <typename T>
static int run_patch_march(FusedParams prm, const FusedPlan& pl, int64_t M, cudaStream_t st) {
    using MT = MarchTile<32, 64, 10>;
    const bool filter = pl.strip_lw != 0;
    // behind the filter pass 1 walks a handful of entries per map: the instantiation that can finish a
    // single-candidate map on its own (window sums kept in 32 KB more shared memory, one CTA per SM)
    // (it runs one CTA per SM: with many maps per call -- measured 768 maps of 64^3 -- the two-launch form
    // with two CTAs per SM walks the entries faster than the fused form saves)
    const bool fuse = filter && M <= 2 * 148;
    auto k1 = fuse ? box_march_kernel<T, double, 32, 64, 10, 1, 1, true> : box_march_kernel<T, double, 32, 64, 10, 1, 2>;
    auto k2 = box_march_kernel<T, double, 32, 64, 10, 2, 2>;
    const size_t smem = MT::smem_bytes<double>();
    const size_t smem1 = smem + (fuse ? (size_t)kFusePlanes * 32 * 64 * sizeof(double) : 0);
    if (cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1) != cudaSuccess ||
        cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return set_error(VALUES_ERR_CUDA, "patch_max: cudaFuncSetAttribute(%zu) failed", smem1);
    prm.use_list = filter ? 1 : 0;
    prm.err_coef = filter_err_coef(prm.zc, prm.p0, 10, StripWide::K - 1, 3);   // == values_patch_filter_err_coef
    if (cudaMemsetAsync(prm.tickets, 0, (size_t)M * 4 * sizeof(unsigned int), st) != cudaSuccess)
        return set_error(VALUES_ERR_CUDA, "patch_max: cudaMemsetAsync failed");
    int rc;
    if (filter) {
        if constexpr (std::is_same<T, float>::value) {
            rc = pl.strip_lw == 32 ? launch_strip_filter<StripWide>(prm, pl, M, st)
                                   : launch_strip_filter<StripNarrow>(prm, pl, M, st);
            if (rc) return rc;
        } else {
            return set_error(VALUES_ERR_UNSUPPORTED, "patch_max: the strip filter runs on fp32 maps");
        }
    }
    for (int64_t m0 = 0; m0 < M; m0 += 65535) {  // gridDim.y limit
        const int64_t mc = std::min<int64_t>(65535, M - m0);
        FusedParams q = prm;
        q.maps = reinterpret_cast<const T*>(prm.maps) + m0 * prm.stride_m;
        q.tile_max = prm.tile_max + m0 * prm.nent;
        q.best = prm.best + m0;
        q.out = prm.out.offset(m0);
        q.gmax = prm.gmax + m0; q.active = prm.active + m0 * (1 + kMaxActive);
        q.tickets = prm.tickets + 4 * m0;
        // the work list normally holds one or two (tile, z sub-chunk) entries per map
        // (every CTA of the grid pays the finish protocol: with hundreds of maps two CTAs per map are enough)
        const unsigned g2 = (unsigned)std::min<int64_t>(prm.nent, mc >= 256 ? 2 : mc >= 16 ? 8 : 32);
        // pass 1: every tile, or (after the filter) the few listed sub-chunks
        const unsigned g1 = filter ? (unsigned)std::min<int64_t>(pl.ntiles, g2) : (unsigned)pl.ntiles;
        k1<<<dim3(g1, (unsigned)mc), kFusedThreads, smem1, st>>>(q);
        if ((rc = check_launch("box_march_kernel<1>"))) return rc;
        k2<<<dim3(g2, (unsigned)mc), kFusedThreads, smem, st>>>(q);
        if ((rc = check_launch("box_march_kernel<2>"))) return rc;
    }
    return VALUES_OK;
}

template <typename T, int TY, int TX>
static int run_patch_fused(FusedParams prm, const FusedPlan& pl, int64_t M, cudaStream_t st) {
    if (prm.p1 == 10 && prm.p2 == 10) return run_patch_fused_pc<T, TY, TX, 10>(prm, pl, M, st);
    return run_patch_fused_pc<T, TY, TX, 0>(prm, pl, M, st);
}

// tile_max [M, nent] | gmax [M] | best [M] | tickets [M, 4] (two 8-byte slots) | active [M, 1 + kMaxActive] ints
static int64_t fused_entries(const FusedPlan& pl) { return pl.march ? pl.ntiles * pl.zsub : pl.ntiles; }
static size_t fused_workspace_bytes(int64_t M, const FusedPlan& pl) {
    return (size_t)(M * fused_entries(pl) + 4 * M) * sizeof(double) + (size_t)M * (1 + kMaxActive) * sizeof(int);
}

}  // namespace vb

using namespace vb;

extern "C" size_t values_map_reduce_workspace_bytes(int64_t M, int64_t V) {
    if (M <= 0 || V <= 0) return 0;
    return (size_t)(M * ceil_div(V, kThreads * kReduceEPT) * 3) * sizeof(double);
}

extern "C" int values_map_reduce(const void* maps, int dtype, int64_t M, int64_t V,
                                 int64_t stride_m, const double* thresholds_host, int n_thresholds,
                                 double* out, void* workspace, size_t workspace_bytes,
                                 void* stream) {
    if (!maps || !out) return set_error(VALUES_ERR_INVALID_ARG, "map_reduce: NULL pointer");
    if (M < 0 || V < 0 || n_thresholds < 0 || n_thresholds > 16)
        return set_error(VALUES_ERR_INVALID_ARG, "map_reduce: bad sizes");
    if (M == 0) return VALUES_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (V == 0) {
        if (cudaMemsetAsync(out, 0, (size_t)M * 3 * sizeof(double), st) != cudaSuccess)
            return set_error(VALUES_ERR_CUDA, "map_reduce: memset failed");
        return VALUES_OK;
    }
    const size_t need = values_map_reduce_workspace_bytes(M, V);
    if (!workspace || workspace_bytes < need)
        return set_error(VALUES_ERR_WORKSPACE, "map_reduce: workspace %zu < %zu", workspace_bytes, need);
    ThrTable thr{};
    thr.n = thresholds_host ? n_thresholds : 0;
    // numpy compares an fp32 image with a Python-float threshold in fp32 (`image >= threshold`,
    // aggregate_uncertainties.py:61-62), and so do K1's fused threshold sums: for fp32 maps the
    // threshold is rounded to fp32 first (a voxel equal to fl32(0.7) counts for threshold 0.7)
    for (int i = 0; i < thr.n; ++i)
        thr.v[i] = dtype == VALUES_F32 ? (double)(float)thresholds_host[i] : thresholds_host[i];
    const int64_t bpm = ceil_div(V, kThreads * kReduceEPT);
    if (bpm * M > 0x7fffffffLL) return set_error(VALUES_ERR_UNSUPPORTED, "map_reduce: grid too large");
    const unsigned grid = (unsigned)(bpm * M);
    double* partials = reinterpret_cast<double*>(workspace);
    switch (dtype) {
        case VALUES_F32:
            map_reduce_kernel<float><<<grid, kThreads, 0, st>>>((const float*)maps, V, stride_m, bpm, thr, partials);
            break;
        case VALUES_F64:
            map_reduce_kernel<double><<<grid, kThreads, 0, st>>>((const double*)maps, V, stride_m, bpm, thr, partials);
            break;
        default: return set_error(VALUES_ERR_INVALID_ARG, "map_reduce: dtype must be f32 or f64");
    }
    int rc = check_launch("map_reduce_kernel");
    if (rc) return rc;
    return launch_reduce_partials(partials, M, bpm, 3, out, st);
}

extern "C" int values_normalize_maps(const void* maps, int dtype, int64_t M, int64_t V,
                                     int64_t stride_m, const double* count, double clip_min,
                                     double* out, void* stream) {
    if (!maps || !count || !out) return set_error(VALUES_ERR_INVALID_ARG, "normalize: NULL pointer");
    if (M < 0 || V < 0) return set_error(VALUES_ERR_INVALID_ARG, "normalize: bad sizes");
    if (M == 0 || V == 0) return VALUES_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype != VALUES_F32 && dtype != VALUES_F64)
        return set_error(VALUES_ERR_INVALID_ARG, "normalize: dtype must be f32 or f64");
    const size_t es = dtype == VALUES_F32 ? 4 : 8;
    if (V % 4 == 0 && stride_m % 4 == 0 && reinterpret_cast<uintptr_t>(maps) % (4 * es) == 0 &&
        reinterpret_cast<uintptr_t>(count) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0) {
        const int64_t blocks = ceil_div(V / 4, kThreads);
        if (blocks > 0x7fffffffLL) return set_error(VALUES_ERR_UNSUPPORTED, "normalize: grid too large");
        if (dtype == VALUES_F32)
            normalize_vec_kernel<float><<<(unsigned)blocks, kThreads, 0, st>>>((const float*)maps, V, M, stride_m, count, clip_min, out);
        else
            normalize_vec_kernel<double><<<(unsigned)blocks, kThreads, 0, st>>>((const double*)maps, V, M, stride_m, count, clip_min, out);
        return check_launch("normalize_vec_kernel");
    }
    const int64_t bpm = ceil_div(V, kThreads);
    if (bpm * M > 0x7fffffffLL) return set_error(VALUES_ERR_UNSUPPORTED, "normalize: grid too large");
    const unsigned grid = (unsigned)(bpm * M);
    switch (dtype) {
        case VALUES_F32:
            normalize_kernel<float><<<grid, kThreads, 0, st>>>((const float*)maps, V, stride_m, bpm, count, clip_min, out);
            break;
        case VALUES_F64:
            normalize_kernel<double><<<grid, kThreads, 0, st>>>((const double*)maps, V, stride_m, bpm, count, clip_min, out);
            break;
        default: return set_error(VALUES_ERR_INVALID_ARG, "normalize: dtype must be f32 or f64");
    }
    return check_launch("normalize_kernel");
}

// workspace layout (march / fused): tile_max [M, nent] | gmax [M] | best [M] | tickets [M, 4] | active [M, 1 + 64]
//                  (tiled fallback): tile_max [M, ntiles] | gmax [M] | best [M]
static bool patch_path_ok(int path) { return path == 0 || path == 2 || path == 4 || path == 5; }

extern "C" size_t values_patch_max_workspace_bytes(int64_t M, const int64_t* shape3_host,
                                                   const int64_t* patch3_host, int path) {
    if (M <= 0 || !shape3_host || !patch3_host || !patch_path_ok(path)) return 0;
    if (path != 2) {
        // dtype and alignment (which decide whether the strip filter and its tiling apply) are not
        // known here: the larger of the two layouts
        size_t need = 0;
        for (int strip = 0; strip < 2; ++strip) {
            FusedPlan fp;
            if (make_fused_plan(M, shape3_host, patch3_host, path, strip != 0, fp) != VALUES_OK) return 0;
            if (fp.ty) need = std::max(need, fused_workspace_bytes(M, fp));
        }
        // rows that are not 16-byte aligned: room for the pitched copy the strip filter runs on
        if (need && path == 0 && shape3_host[2] % 4 != 0 && march_applies(path, shape3_host, patch3_host))
            need = (need + 255) / 256 * 256 + pitched_scratch_bytes(M, shape3_host);
        if (need) return need;
    }
    PatchPlan pl;
    if (make_patch_plan(shape3_host, patch3_host, pl) != VALUES_OK) return 0;
    return (size_t)(M * pl.ntiles + 2 * M) * sizeof(double);
}

extern "C" int values_patch_max(const void* maps, int dtype, int64_t M, int64_t stride_m,
                                const int64_t* shape3_host, const int64_t* patch3_host,
                                int mean_flag, double rtol, double atol, double* max_score,
                                int64_t score_stride, void* bbox_lo, int bbox_dtype, int64_t bbox_stride,
                                void* workspace, size_t workspace_bytes, int path, void* stream) {
    if (!shape3_host || !patch3_host) return set_error(VALUES_ERR_INVALID_ARG, "patch_max: NULL shape");
    if (M < 0) return set_error(VALUES_ERR_INVALID_ARG, "patch_max: M < 0");
    if (dtype != VALUES_F32 && dtype != VALUES_F64)
        return set_error(VALUES_ERR_INVALID_ARG, "patch_max: dtype must be f32 or f64");
    if (!patch_path_ok(path)) return set_error(VALUES_ERR_INVALID_ARG, "patch_max: unknown path %d (0, 2, 4, 5)", path);
    if (bbox_dtype != VALUES_I64 && bbox_dtype != VALUES_F64)
        return set_error(VALUES_ERR_INVALID_ARG, "patch_max: bbox_dtype must be i64 or f64");
    if (score_stride < 1 || bbox_stride < 3)
        return set_error(VALUES_ERR_INVALID_ARG, "patch_max: score_stride >= 1 and bbox_stride >= 3");
    PatchOut out{};
    out.max_score = max_score; out.score_stride = score_stride; out.bbox_stride = bbox_stride;
    if (bbox_dtype == VALUES_F64) out.bbox_f64 = reinterpret_cast<double*>(bbox_lo);
    else out.bbox_i64 = reinterpret_cast<int64_t*>(bbox_lo);
    cudaStream_t st = (cudaStream_t)stream;
    double* ws = reinterpret_cast<double*>(workspace);
    const double denom =
        mean_flag ? (double)patch3_host[0] * (double)patch3_host[1] * (double)patch3_host[2] : 1.0;
    if (path != 2) {
        FusedPlan fp;
        bool strip = path == 0 && M > 0 && maps && march_applies(path, shape3_host, patch3_host) &&
                     strip_filter_ok(maps, dtype, stride_m, shape3_host);
        // fp32 maps with an innermost extent that is not a multiple of 4: the strip filter on a pitched copy,
        // if the caller's workspace has room for it (values_patch_max_workspace_bytes asks for it)
        bool pitched = false;
        if (!strip && path == 0 && M > 0 && maps && dtype == VALUES_F32 && shape3_host[2] % 4 != 0 &&
            march_applies(path, shape3_host, patch3_host) && encode_tiled_fn() != nullptr) {
            FusedPlan probe;
            int rcp = make_fused_plan(M, shape3_host, patch3_host, path, true, probe);
            if (rcp) return rcp;
            const size_t base = (fused_workspace_bytes(M, probe) + 255) / 256 * 256;
            pitched = probe.ty && workspace && workspace_bytes >= base + pitched_scratch_bytes(M, shape3_host) &&
                      (int64_t)M * shape3_host[0] * shape3_host[1] < (1LL << 40);
            strip = pitched;
        }
        int rc = make_fused_plan(M, shape3_host, patch3_host, path, strip, fp);
        if (rc) return rc;
        if (fp.ty) {
            if (M == 0) return VALUES_OK;
            if (!maps || !max_score || !bbox_lo) return set_error(VALUES_ERR_INVALID_ARG, "patch_max: NULL pointer");
            const size_t need = fused_workspace_bytes(M, fp);
            if (!workspace || workspace_bytes < need)
                return set_error(VALUES_ERR_WORKSPACE, "patch_max: workspace %zu < %zu", workspace_bytes, need);
            FusedParams prm{};
            prm.maps = maps; prm.stride_m = stride_m;
            prm.D0 = shape3_host[0]; prm.D1 = shape3_host[1]; prm.D2 = shape3_host[2];
            prm.pitch = prm.D2;
            if (pitched) {
                const size_t base = (need + 255) / 256 * 256;
                uintptr_t sp = (reinterpret_cast<uintptr_t>(workspace) + base + 255) / 256 * 256;
                float* scratch = reinterpret_cast<float*>(sp);
                const int64_t rows_per_map = prm.D0 * prm.D1, n_rows = M * rows_per_map;
                const int pw = (int)pitched_width(prm.D2);
                const int64_t want = ceil_div(n_rows * 32, kThreads);
                const unsigned grid = (unsigned)std::min<int64_t>(want, 148 * 16);
                pitch_rows_kernel<<<grid, kThreads, 0, st>>>((const float*)maps, stride_m, rows_per_map, n_rows,
                                                            (int)prm.D2, pw, scratch);
                if ((rc = check_launch("pitch_rows_kernel"))) return rc;
                prm.maps = scratch; prm.pitch = pw; prm.stride_m = rows_per_map * pw;
            }
            prm.O0 = fp.O0; prm.O1 = fp.O1; prm.O2 = fp.O2;
            prm.p0 = (int)patch3_host[0]; prm.p1 = (int)patch3_host[1]; prm.p2 = (int)patch3_host[2];
            prm.tiles_x = fp.tiles_x; prm.tiles_y = fp.tiles_y; prm.chunks_z = fp.chunks_z; prm.zc = fp.zc;
            prm.zsub = fp.zsub; prm.zc_fine = fp.zc_fine; prm.ntiles = fp.ntiles;
            prm.tx_per = fp.tx_per; prm.xs_stride = fp.xs_stride;
            prm.nent = fused_entries(fp);
            prm.denom = denom; prm.mean_flag = mean_flag ? 1 : 0; prm.rtol = rtol; prm.atol = atol;
            prm.tile_max = ws;
            prm.gmax = ws + M * prm.nent;
            prm.best = reinterpret_cast<unsigned long long*>(ws + M * prm.nent + M);
            prm.tickets = reinterpret_cast<unsigned int*>(ws + M * prm.nent + 2 * M);
            prm.active = reinterpret_cast<int*>(ws + M * prm.nent + 4 * M);
            prm.out = out;
            if (fp.march) {
                if (dtype == VALUES_F32) return run_patch_march<float>(prm, fp, M, st);
                return run_patch_march<double>(prm, fp, M, st);
            }
            if (fp.ty == 16) {
                if (dtype == VALUES_F32) return run_patch_fused<float, 16, 64>(prm, fp, M, st);
                return run_patch_fused<double, 16, 64>(prm, fp, M, st);
            }
            if (dtype == VALUES_F32) return run_patch_fused<float, 8, 32>(prm, fp, M, st);
            return run_patch_fused<double, 8, 32>(prm, fp, M, st);
        }
    }
    PatchPlan pl;
    int rc = make_patch_plan(shape3_host, patch3_host, pl);
    if (rc) return rc;
    if (M == 0) return VALUES_OK;
    if (!maps || !max_score || !bbox_lo) return set_error(VALUES_ERR_INVALID_ARG, "patch_max: NULL pointer");
    const size_t need = (size_t)(M * pl.ntiles + 2 * M) * sizeof(double);
    if (!workspace || workspace_bytes < need)
        return set_error(VALUES_ERR_WORKSPACE, "patch_max: workspace %zu < %zu", workspace_bytes, need);
    PatchParams prm{};
    prm.maps = maps; prm.stride_m = stride_m;
    prm.D0 = shape3_host[0]; prm.D1 = shape3_host[1]; prm.D2 = shape3_host[2];
    prm.O0 = pl.O0; prm.O1 = pl.O1; prm.O2 = pl.O2;
    prm.p0 = (int)patch3_host[0]; prm.p1 = (int)patch3_host[1]; prm.p2 = (int)patch3_host[2];
    prm.tiles_x = pl.tiles_x; prm.tiles_y = pl.tiles_y; prm.chunks_z = pl.chunks_z; prm.zc = pl.zc;
    prm.ntiles = pl.ntiles;
    prm.mean_flag = mean_flag ? 1 : 0;
    prm.denom = denom;
    prm.rtol = rtol; prm.atol = atol;
    prm.tile_max = ws;
    double* gmax = ws + M * pl.ntiles;
    prm.best = reinterpret_cast<unsigned long long*>(gmax + M);
    if (dtype == VALUES_F32) return run_patch<float>(prm, pl, M, gmax, out, st);
    return run_patch<double>(prm, pl, M, gmax, out, st);
}

extern "C" double values_patch_filter_err_coef(int zc, int p0) {
    if (zc <= 0 || p0 <= 0) return 0.0;
    return filter_err_coef(zc, p0, 10, StripWide::K - 1, 3);
}
