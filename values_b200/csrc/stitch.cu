// K3: sliding-window stitch accumulator -- atomic-free, output-stationary.
//
// The reference (DataCarrier3D.concat_data, uncertainty_modeling/data_carrier_3D.py:154-179)
// does one device->host copy and one strided numpy `+=` per (patch, sample).  Here every
// CTA owns a tile of the OUTPUT volume, finds the patches that overlap it (ordered
// compaction, so the summation order is the reference's list order), sums them in registers
// and writes each output voxel exactly once.  No atomics, bit-reproducible.
// Algorithmic bytes: every patch element read once + every output element written once.
#include "common.cuh"
#include "tma_host.cuh"

namespace vb {

constexpr int kSZ = 64;  // tile extent along the contiguous axis (z)
constexpr int kSY = 4;   // tile extent along axis 1 (y); kSZ*kSY == kThreads
constexpr int kMaxList = 1024;  // overlapping patches kept per tile chunk

struct StitchParams {
    const void* patches;
    int64_t stride_n, stride_p;
    const int32_t* patch_index;
    const int32_t* crop_lo;
    const double* weight;   // [p0, p1, p2] importance map or NULL (uniform, the reference)
    const double* wsep[3];  // separable importance map: weight[x][y][z] = fl(fl(wx[x] * wy[y]) * wz[z]), or NULLs
    int64_t n_sel, N, C;
    int p0, p1, p2;
    int64_t X, Y, Z;
    void* out_sum;
    double* out_count;
    int accumulate;
    int tiles_z, tiles_y;
};

template <typename TP, typename TO>
__global__ void __launch_bounds__(kThreads) stitch_kernel(const StitchParams prm) {
    __shared__ int s_list[kMaxList];
    __shared__ int s_warp_cnt[kThreads / 32];
    __shared__ int s_total;

    int tile = blockIdx.x;
    const int tz = tile % prm.tiles_z; tile /= prm.tiles_z;
    const int ty = tile % prm.tiles_y; tile /= prm.tiles_y;
    const int64_t x = tile;
    const int64_t n = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t y_lo = (int64_t)ty * kSY, z_lo = (int64_t)tz * kSZ;
    const int64_t y = y_lo + tid / kSZ, z = z_lo + tid % kSZ;
    const bool inside = y < prm.Y && z < prm.Z;
    const int64_t vox = (x * prm.Y + y) * prm.Z + z;
    const int64_t vol = prm.X * prm.Y * prm.Z;
    const int64_t pvol = (int64_t)prm.p0 * prm.p1 * prm.p2;
    TO* out = reinterpret_cast<TO*>(prm.out_sum) + n * prm.C * vol;
    const TP* pin = reinterpret_cast<const TP*>(prm.patches) + n * prm.stride_n;

    // registers: one accumulator per class would need compile-time C; instead classes are the
    // outer loop and the (short) overlap list is re-walked per class.
    // the count continues from the stored value, patch by patch, as numpy's `count[crop] += w` does (a local
    // sum added at the end rounds differently once the weights are not integers)
    double cnt = (inside && n == 0 && prm.out_count && prm.accumulate) ? prm.out_count[vox] : 0.0;
    for (int64_t base = 0; base < prm.n_sel; base += kMaxList) {
        // ---- ordered compaction of the patches overlapping this tile (chunk of kMaxList)
        const int64_t chunk = min((int64_t)kMaxList, prm.n_sel - base);
        if (tid == 0) s_total = 0;
        __syncthreads();
        for (int64_t off = 0; off < chunk; off += kThreads) {
            const int64_t i = base + off + tid;
            bool hit = false;
            if (off + tid < chunk) {
                const int cx = prm.crop_lo[3 * i], cy = prm.crop_lo[3 * i + 1], cz = prm.crop_lo[3 * i + 2];
                hit = x >= cx && x < cx + prm.p0 && y_lo < cy + prm.p1 && y_lo + kSY > cy &&
                      z_lo < cz + prm.p2 && z_lo + kSZ > cz;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, hit);
            if (lane == 0) s_warp_cnt[warp] = __popc(bal);
            __syncthreads();
            int before = s_total;
            for (int w = 0; w < warp; ++w) before += s_warp_cnt[w];
            if (hit) s_list[before + __popc(bal & ((1u << lane) - 1u))] = (int)(off + tid);
            __syncthreads();
            if (tid == 0) {
                int t = s_total;
                for (int w = 0; w < kThreads / 32; ++w) t += s_warp_cnt[w];
                s_total = t;
            }
            __syncthreads();
        }
        const int total = s_total;
        if (inside) {
            for (int64_t c = 0; c < prm.C; ++c) {
                double acc = 0.0;
                TO* dst = out + c * vol + vox;
                if (base > 0 || prm.accumulate) acc = (double)*dst;
                for (int k = 0; k < total; ++k) {
                    const int64_t i = base + s_list[k];
                    const int cx = prm.crop_lo[3 * i], cy = prm.crop_lo[3 * i + 1], cz = prm.crop_lo[3 * i + 2];
                    if (y >= cy && y < cy + prm.p1 && z >= cz && z < cz + prm.p2) {
                        const int64_t pi = prm.patch_index ? prm.patch_index[i] : i;
                        const int64_t local = ((x - cx) * prm.p1 + (y - cy)) * prm.p2 + (z - cz);
                        const double v = (double)In<TP>::load_one(pin + pi * prm.stride_p + c * pvol + local);
                        if (prm.weight) {
                            const double w = __ldg(prm.weight + local);
                            acc = __dadd_rn(acc, __dmul_rn(w, v));   // two roundings, as `sum += w * patch`
                            if (c == 0) cnt += w;
                        } else {
                            acc += v;
                            if (c == 0) cnt += 1.0;
                        }
                    }
                }
                *dst = (TO)acc;
            }
        }
        __syncthreads();
    }
    if (inside && n == 0 && prm.out_count) {
        prm.out_count[vox] = cnt;
    }
}


// ------------------------------------------------------------------ K3 vector kernel
// Same output-stationary scheme, but a thread owns FOUR consecutive z voxels of one (x, y) row
// and kCB classes at once: per overlapping patch it issues kCB independent 16-byte loads (fp32
// patches), so a CTA keeps ~8 KB in flight instead of 1 KB -- the scalar kernel is latency
// bound at ~0.13 of HBM peak.  Patches whose z origin is not a multiple of 4 are handled element
// by element inside the same kernel (uniform per patch, no divergence).  Requires Z % 4 == 0,
// p2 % 4 == 0 and 16-byte aligned patch rows / outputs (checked on the host).
constexpr int kCB = 2;   // classes per sweep of the overlap list (the reference has C == 2)

// four consecutive elements: raw load (kept in flight as bits) and widening to fp64
template <typename T> struct Raw4;
template <> struct Raw4<float> {
    using type = uint4;
    __device__ static __forceinline__ type load(const float* p) { return ldg_stream_128(p); }
    __device__ static __forceinline__ void widen(const type& r, double (&o)[4]) {
        o[0] = (double)__uint_as_float(r.x); o[1] = (double)__uint_as_float(r.y);
        o[2] = (double)__uint_as_float(r.z); o[3] = (double)__uint_as_float(r.w);
    }
};
template <> struct Raw4<double> {
    struct type { uint4 a, b; };
    __device__ static __forceinline__ type load(const double* p) { return {ldg_stream_128(p), ldg_stream_128(p + 2)}; }
    __device__ static __forceinline__ void widen(const type& r, double (&o)[4]) {
        o[0] = __hiloint2double((int)r.a.y, (int)r.a.x); o[1] = __hiloint2double((int)r.a.w, (int)r.a.z);
        o[2] = __hiloint2double((int)r.b.y, (int)r.b.x); o[3] = __hiloint2double((int)r.b.w, (int)r.b.z);
    }
};
template <> struct Raw4<__nv_bfloat16> {
    using type = uint2;
    __device__ static __forceinline__ type load(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint2*>(p)); }
    __device__ static __forceinline__ void widen(const type& r, double (&o)[4]) {
        o[0] = (double)__uint_as_float(r.x << 16); o[1] = (double)__uint_as_float(r.x & 0xffff0000u);
        o[2] = (double)__uint_as_float(r.y << 16); o[3] = (double)__uint_as_float(r.y & 0xffff0000u);
    }
};
template <typename TO> __device__ __forceinline__ void store4(TO* p, const double (&v)[4]);
template <> __device__ __forceinline__ void store4<double>(double* p, const double (&v)[4]) {
    reinterpret_cast<double2*>(p)[0] = make_double2(v[0], v[1]);
    reinterpret_cast<double2*>(p)[1] = make_double2(v[2], v[3]);
}
template <> __device__ __forceinline__ void store4<float>(float* p, const double (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4((float)v[0], (float)v[1], (float)v[2], (float)v[3]);
}
template <typename TO> __device__ __forceinline__ void read4(const TO* p, double (&v)[4]) {
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] = (double)p[e];
}

constexpr int kXB = 8;   // x planes per CTA of the vector kernel (one overlap list serves them all)

template <typename TP, typename TO, int TZ, bool WEIGHTED>
__global__ void __launch_bounds__(kThreads, 3) stitch_vec_kernel(const StitchParams prm) {
    constexpr int TYV = kThreads / TZ;          // rows per tile
    __shared__ int4 s_list[kMaxList];           // {cx, cy, cz, patch index} of the overlapping patches
    __shared__ long long s_base[64];            // element offset of patch voxel (0,0,0) minus its crop origin
    __shared__ int s_warp_cnt[kThreads / 32];
    __shared__ int s_total;

    int tile = blockIdx.x;
    const int tz = tile % prm.tiles_z; tile /= prm.tiles_z;
    const int ty = tile % prm.tiles_y; tile /= prm.tiles_y;
    const int64_t x_lo = (int64_t)tile * kXB;
    const int64_t x_hi = min(x_lo + kXB, prm.X);
    const int64_t n = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t y_lo = (int64_t)ty * TYV, z_lo = (int64_t)tz * (TZ * 4);
    const int64_t y = y_lo + tid / TZ, z = z_lo + (int64_t)(tid % TZ) * 4;
    const bool inside = y < prm.Y && z < prm.Z;     // Z % 4 == 0: a vector is inside or outside
    const int64_t vol = prm.X * prm.Y * prm.Z;
    const int64_t pvol = (int64_t)prm.p0 * prm.p1 * prm.p2;
    TO* out = reinterpret_cast<TO*>(prm.out_sum) + n * prm.C * vol;
    const TP* pin = reinterpret_cast<const TP*>(prm.patches) + n * prm.stride_n;

    for (int64_t base = 0; base < prm.n_sel; base += kMaxList) {
        // ---- ordered compaction of the patches overlapping this CTA's box (chunk of kMaxList)
        const int64_t chunk = min((int64_t)kMaxList, prm.n_sel - base);
        if (tid == 0) s_total = 0;
        __syncthreads();
        for (int64_t off = 0; off < chunk; off += kThreads) {
            const int64_t i = base + off + tid;
            bool hit = false;
            int4 e = make_int4(0, 0, 0, 0);
            if (off + tid < chunk) {
                e.x = prm.crop_lo[3 * i]; e.y = prm.crop_lo[3 * i + 1]; e.z = prm.crop_lo[3 * i + 2];
                e.w = prm.patch_index ? prm.patch_index[i] : (int)i;
                hit = x_lo < e.x + prm.p0 && x_hi > e.x && y_lo < e.y + prm.p1 && y_lo + TYV > e.y &&
                      z_lo < e.z + prm.p2 && z_lo + TZ * 4 > e.z;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, hit);
            if (lane == 0) s_warp_cnt[warp] = __popc(bal);
            __syncthreads();
            int before = s_total;
            for (int w = 0; w < warp; ++w) before += s_warp_cnt[w];
            if (hit) s_list[before + __popc(bal & ((1u << lane) - 1u))] = e;
            __syncthreads();
            if (tid == 0) {
                int t = s_total;
                for (int w = 0; w < kThreads / 32; ++w) t += s_warp_cnt[w];
                s_total = t;
            }
            __syncthreads();
        }
        const int total = s_total;
        const bool readback = base > 0 || prm.accumulate;
        // one patch: add its four voxels (all classes of the chunk) to the accumulators
        auto add_patch = [&](const int4& e, int64_t local, const TP* src, int64_t c0, double (&acc)[kCB][4],
                             double (&cnt)[4]) {
            if ((e.z & 3) == 0) {         // aligned z origin: the vector is fully inside the patch
                typename Raw4<TP>::type raw[kCB];
#pragma unroll
                for (int cc = 0; cc < kCB; ++cc)
                    if (c0 + cc < prm.C) raw[cc] = Raw4<TP>::load(src + cc * pvol);
                double w[4] = {1.0, 1.0, 1.0, 1.0};
                if (WEIGHTED) Raw4<double>::widen(Raw4<double>::load(prm.weight + local), w);
#pragma unroll
                for (int cc = 0; cc < kCB; ++cc) {
                    if (c0 + cc < prm.C) {
                        double v[4];
                        Raw4<TP>::widen(raw[cc], v);
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            acc[cc][q] = WEIGHTED ? __dadd_rn(acc[cc][q], __dmul_rn(w[q], v[q])) : acc[cc][q] + v[q];
                    }
                }
                if (c0 == 0) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) cnt[q] += w[q];
                }
            } else {                      // unaligned z origin: element by element
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (z + q < e.z || z + q >= e.z + prm.p2) continue;
                    const double w = WEIGHTED ? __ldg(prm.weight + local + q) : 1.0;
#pragma unroll
                    for (int cc = 0; cc < kCB; ++cc) {
                        if (c0 + cc < prm.C) {
                            const double v = (double)In<TP>::load_one(src + cc * pvol + q);
                            acc[cc][q] = WEIGHTED ? __dadd_rn(acc[cc][q], __dmul_rn(w, v)) : acc[cc][q] + v;
                        }
                    }
                    if (c0 == 0) cnt[q] += w;
                }
            }
        };
        if (total <= 64) {
            // Fast path: the (y, z) test of this thread against every listed patch is folded into one
            // 64-bit mask, and each patch's address arithmetic into one shared 64-bit base, so the
            // x / class loops only walk set bits.
            if (tid < total) {
                const int4 e = s_list[tid];
                s_base[tid] = (long long)e.w * prm.stride_p - (((long long)e.x * prm.p1 + e.y) * prm.p2 + e.z);
            }
            __syncthreads();
            unsigned long long mask = 0;
            if (inside) {
                for (int k = 0; k < total; ++k) {
                    const int4 e = s_list[k];
                    if (y >= e.y && y < e.y + prm.p1 && z + 3 >= e.z && z < e.z + prm.p2) mask |= 1ull << k;
                }
            }
            for (int64_t x = x_lo; x < x_hi && mask; ++x) {
                const int64_t vox = (x * prm.Y + y) * prm.Z + z;
                const int64_t toff = (x * prm.p1 + y) * prm.p2 + z;
                double cnt[4] = {0.0, 0.0, 0.0, 0.0};
                // the count continues from the stored value, patch by patch (numpy's `count[crop] += w`)
                if (readback && n == 0 && prm.out_count) read4(prm.out_count + vox, cnt);
                for (int64_t c0 = 0; c0 < prm.C; c0 += kCB) {
                    double acc[kCB][4];
#pragma unroll
                    for (int cc = 0; cc < kCB; ++cc) {
                        if (readback && c0 + cc < prm.C) read4(out + (c0 + cc) * vol + vox, acc[cc]);
                        else { acc[cc][0] = acc[cc][1] = acc[cc][2] = acc[cc][3] = 0.0; }
                    }
                    unsigned long long m = mask;
                    while (m) {
                        const int k = __ffsll((long long)m) - 1;
                        m &= m - 1;
                        const int4 e = s_list[k];
                        if (x < e.x || x >= e.x + prm.p0) continue;
                        const int64_t sb = s_base[k] + toff;
                        add_patch(e, sb - (int64_t)e.w * prm.stride_p, pin + sb + c0 * pvol, c0, acc, cnt);
                    }
#pragma unroll
                    for (int cc = 0; cc < kCB; ++cc)
                        if (c0 + cc < prm.C) store4<TO>(out + (c0 + cc) * vol + vox, acc[cc]);
                }
                if (n == 0 && prm.out_count) store4<double>(prm.out_count + vox, cnt);
            }
            if (inside && !mask && !readback) {   // no patch covers this thread's voxels: zeros
                const double zero[4] = {0.0, 0.0, 0.0, 0.0};
                for (int64_t x = x_lo; x < x_hi; ++x) {
                    const int64_t vox = (x * prm.Y + y) * prm.Z + z;
                    for (int64_t c = 0; c < prm.C; ++c) store4<TO>(out + c * vol + vox, zero);
                    if (n == 0 && prm.out_count) store4<double>(prm.out_count + vox, zero);
                }
            }
        } else if (inside) {
            // Dense overlap (more than 64 patches touch the box): walk the whole list per voxel.
            for (int64_t x = x_lo; x < x_hi; ++x) {
                const int64_t vox = (x * prm.Y + y) * prm.Z + z;
                double cnt[4] = {0.0, 0.0, 0.0, 0.0};
                // the count continues from the stored value, patch by patch (numpy's `count[crop] += w`)
                if (readback && n == 0 && prm.out_count) read4(prm.out_count + vox, cnt);
                for (int64_t c0 = 0; c0 < prm.C; c0 += kCB) {
                    double acc[kCB][4];
#pragma unroll
                    for (int cc = 0; cc < kCB; ++cc) {
                        if (readback && c0 + cc < prm.C) read4(out + (c0 + cc) * vol + vox, acc[cc]);
                        else { acc[cc][0] = acc[cc][1] = acc[cc][2] = acc[cc][3] = 0.0; }
                    }
                    for (int k = 0; k < total; ++k) {
                        const int4 e = s_list[k];
                        if (x < e.x || x >= e.x + prm.p0 || y < e.y || y >= e.y + prm.p1 || z + 3 < e.z ||
                            z >= e.z + prm.p2)
                            continue;
                        const int64_t local = ((x - e.x) * prm.p1 + (y - e.y)) * prm.p2 + (z - e.z);
                        add_patch(e, local, pin + (int64_t)e.w * prm.stride_p + c0 * pvol + local, c0, acc, cnt);
                    }
#pragma unroll
                    for (int cc = 0; cc < kCB; ++cc)
                        if (c0 + cc < prm.C) store4<TO>(out + (c0 + cc) * vol + vox, acc[cc]);
                }
                if (n == 0 && prm.out_count) store4<double>(prm.out_count + vox, cnt);
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ K3 box kernel (TMA ring)
// The same output-stationary scheme with the patch data arriving through the copy engine: a CTA owns
// an output box of kBX x kBY x kBZ voxels of one sample; for every class and every patch that
// overlaps the box (list order = the reference's summation order) ONE cp.async.bulk.tensor copy
// brings the box-shaped window of that patch into a shared-memory ring -- the part of the window
// that lies outside the patch arrives as zeros -- while the eight consumer warps add the previous
// window into their fp64 accumulators (four float4 groups of voxels per thread, conflict-free
// LDS.128).  kStages windows (16 KB each for fp32 patches) are in flight per CTA regardless of how
// far the arithmetic has got, which is what the register-staged vector kernel lacked (ncu r01e:
// long-scoreboard bound, 32 B in flight per thread).  Every output voxel is still written exactly
// once and sums are bit-identical to the other kernels.
// A patch whose z origin is not a multiple of the 16-byte copy granularity cannot be fetched this
// way (the innermost tensor coordinate must be 16-byte aligned); the consumers read such a patch
// straight from global memory, element by element, in its place in the list.
constexpr int kBX = 8, kBY = 16, kBZ = 32;             // output box of a CTA
constexpr int kBoxVox = kBX * kBY * kBZ;               // 4096 voxels = 1024 groups of 4 = 4 groups per thread
constexpr int kGroups = kBoxVox / 4 / kThreads;
constexpr int kBoxList = 512;                           // overlapping patches kept per round (8 KB: two CTAs per SM)

template <typename TP> struct StitchRing {
    static constexpr int kBytes = kBoxVox * (int)sizeof(TP);          // one window
    static constexpr int kStages = sizeof(TP) == 8 ? 3 : 6;
    static constexpr int kAlign = sizeof(TP) == 2 ? 8 : 4;            // elements: 16-byte copies and whole groups of 4
    static constexpr size_t smem = (size_t)kStages * kBytes + 128;
};

// four consecutive elements from shared memory, widened to fp64
template <typename TP> __device__ __forceinline__ void lds_group(uint32_t addr, double (&o)[4]);
template <> __device__ __forceinline__ void lds_group<float>(uint32_t addr, double (&o)[4]) {
    const uint4 r = lds128(addr);
    Raw4<float>::widen(r, o);
}
template <> __device__ __forceinline__ void lds_group<__nv_bfloat16>(uint32_t addr, double (&o)[4]) {
    uint2 r;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(addr) : "memory");
    Raw4<__nv_bfloat16>::widen(r, o);
}
template <> __device__ __forceinline__ void lds_group<double>(uint32_t addr, double (&o)[4]) {
    Raw4<double>::type r;
    r.a = lds128(addr); r.b = lds128(addr + 16);
    Raw4<double>::widen(r, o);
}

constexpr int kXSteps = 4;                              // boxes a CTA walks along x (one overlap list, one ring)

// WMODE: 0 uniform (the reference), 1 importance map read from global memory (any map), 2 separable
// importance map: the three 1-D factors sit in shared memory and the weight of a voxel is two multiplies
// away -- no weight traffic at all (the map read from L2 next to the data ran at 0.33 of the HBM peak).
constexpr int kSepMax = 128;                            // longest patch edge of the separable path
template <typename TP, typename TO, int WMODE>
__global__ void __launch_bounds__(kThreads + 32, 2)
stitch_box_kernel(const __grid_constant__ CUtensorMap tmap, const StitchParams prm, int rows_per_sample) {
    using SR = StitchRing<TP>;
    constexpr bool WEIGHTED = WMODE != 0;
    __shared__ __align__(16) double s_wx[WMODE == 2 ? kSepMax : 1], s_wy[WMODE == 2 ? kSepMax : 1];
    __shared__ __align__(32) double s_wz[WMODE == 2 ? kSepMax + 4 : 4];
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[SR::kStages];
    __shared__ __align__(8) uint64_t empty_bar[SR::kStages];
    __shared__ int4 s_list[kBoxList];           // {cx, cy, cz, patch index} of the overlapping patches
    __shared__ unsigned char s_flag[kBoxList];  // bit 0: fetched by the copy engine (aligned z origin), bit 1: covers the box in y and z
    __shared__ int s_warp_cnt[kThreads / 32];
    __shared__ int s_total;

    int tile = blockIdx.x;
    const int tz = tile % prm.tiles_z; tile /= prm.tiles_z;
    const int ty = tile % prm.tiles_y; tile /= prm.tiles_y;
    const int col_lo = tile * (kBX * kXSteps), y_lo = ty * kBY, z_lo = tz * kBZ;
    const int col_hi = (int)min((int64_t)col_lo + kBX * kXSteps, prm.X);
    const int64_t n = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool producer = warp == kThreads / 32;
    const int64_t vol = prm.X * prm.Y * prm.Z;
    const int64_t pvol = (int64_t)prm.p0 * prm.p1 * prm.p2;
    TO* out = reinterpret_cast<TO*>(prm.out_sum) + n * prm.C * vol;
    const TP* pin = reinterpret_cast<const TP*>(prm.patches) + n * prm.stride_n;
    const uint32_t ring = (smem_u32(smem_raw) + 127u) & ~127u;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < SR::kStages; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, kThreads / 32 * kArriveLanes); }
        mbar_fence_init();
    }
    if constexpr (WMODE == 2) {   // visible after the first block barrier of the list compaction below
        for (int i = threadIdx.x; i < kSepMax; i += kThreads + 32) {
            s_wx[i] = i < prm.p0 ? prm.wsep[0][i] : 0.0;
            s_wy[i] = i < prm.p1 ? prm.wsep[1][i] : 0.0;
            s_wz[i] = i < prm.p2 ? prm.wsep[2][i] : 0.0;
        }
        if (threadIdx.x < 4) s_wz[kSepMax + threadIdx.x] = 0.0;
    }
    // this thread's voxel groups: group g = tid + kThreads * i -> (dx, y, z) inside a box
    int gdx[kGroups], gy[kGroups], gz[kGroups];
    bool gyz[kGroups];
#pragma unroll
    for (int i = 0; i < kGroups; ++i) {
        const int g = (tid & (kThreads - 1)) + kThreads * i, row = g / (kBZ / 4);
        gdx[i] = row / kBY; gy[i] = y_lo + row % kBY; gz[i] = z_lo + (g % (kBZ / 4)) * 4;
        gyz[i] = !producer && gy[i] < prm.Y && gz[i] < prm.Z;                    // Z % 4 == 0: whole groups
    }
    const uint32_t lds_off = (uint32_t)(tid & (kThreads - 1)) * 4u * (uint32_t)sizeof(TP);
    // Z % 4 != 0: the volume's rows are not whole (or aligned) groups of four voxels -- the sums are read back
    // and written element by element, the last group of a row guarded; the arithmetic between is the same
    const bool zr = (prm.Z & 3) != 0;
    unsigned int it = 0;        // windows that went through the ring so far (producer and consumers count alike)
    // Rounds: the candidate list is scanned in order until kBoxList overlapping patches are found (one
    // round in every practical case: a box is overlapped by a few dozen patches at most); a further
    // round continues from the sums of the previous one.
    int64_t cursor = 0;
    for (bool first = true; first || cursor < prm.n_sel; first = false) {
        // ---- ordered compaction of the patches overlapping this CTA's column of boxes
        if (tid == 0) s_total = 0;
        __syncthreads();
        while (cursor < prm.n_sel && s_total + kThreads <= kBoxList) {
            const int64_t i = cursor + tid;
            cursor += kThreads;
            bool hit = false;
            int4 e = make_int4(0, 0, 0, 0);
            if (!producer && i < prm.n_sel) {
                e.x = prm.crop_lo[3 * i]; e.y = prm.crop_lo[3 * i + 1]; e.z = prm.crop_lo[3 * i + 2];
                e.w = prm.patch_index ? prm.patch_index[i] : (int)i;
                hit = col_lo < e.x + prm.p0 && col_hi > e.x && y_lo < e.y + prm.p1 && y_lo + kBY > e.y &&
                      z_lo < e.z + prm.p2 && z_lo + kBZ > e.z;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, hit);
            if (lane == 0 && !producer) s_warp_cnt[warp] = __popc(bal);
            __syncthreads();
            int before = s_total;
            for (int w = 0; w < warp && w < kThreads / 32; ++w) before += s_warp_cnt[w];
            if (hit) {
                const int slot = before + __popc(bal & ((1u << lane) - 1u));
                s_list[slot] = e;
                s_flag[slot] = (unsigned char)((e.z % SR::kAlign == 0 ? 1 : 0) |
                                               ((e.y <= y_lo && e.y + prm.p1 >= y_lo + kBY && e.z <= z_lo &&
                                                 e.z + prm.p2 >= z_lo + kBZ) ? 2 : 0));
            }
            __syncthreads();
            if (tid == 0) {
                int t = s_total;
                for (int w = 0; w < kThreads / 32; ++w) t += s_warp_cnt[w];
                s_total = t;
            }
            __syncthreads();
        }
        const int total = s_total;
        const bool readback = !first || prm.accumulate;
        for (int x_lo = col_lo; x_lo < col_hi; x_lo += kBX) {
            const int x_hi = min(x_lo + kBX, col_hi);
            if (producer) {
                // ---- one lane streams the windows of this box: classes outer, listed patches inner
                if (lane == 0) {
                    for (int c = 0; c < (int)prm.C; ++c) {
                        for (int k = 0; k < total; ++k) {
                            const int4 e = s_list[k];
                            if (!(s_flag[k] & 1) || !(x_lo < e.x + prm.p0 && x_hi > e.x)) continue;
                            const int stage = it % SR::kStages;
                            if (it >= SR::kStages) mbar_wait(empty_bar + stage, ((it / SR::kStages) & 1) ^ 1u);
                            mbar_expect_tx(full_bar + stage, SR::kBytes);
                            tma_load_5d_addr(ring + stage * SR::kBytes, &tmap, z_lo - e.z, y_lo - e.y, x_lo - e.x, c,
                                             (int)(n * rows_per_sample + e.w), full_bar + stage);
                            ++it;
                        }
                    }
                }
                continue;
            }
            // accumulating on top of earlier sums (DataCarrier3D.concat_data: a handful of patches per call):
            // a box no listed patch reaches keeps its sums -- no read-modify-write of the untouched volume
            if (readback) {
                bool any = false;
                for (int k = 0; k < total && !any; ++k) {
                    const int ex = s_list[k].x;
                    any = x_lo < ex + prm.p0 && x_hi > ex;
                }
                if (!any) continue;
            }
            bool gin[kGroups];
#pragma unroll
            for (int i = 0; i < kGroups; ++i) gin[i] = gyz[i] && x_lo + gdx[i] < prm.X;
            // pass c < C: class c; pass C (sample 0 only): the count -- same walk, no data
            const int passes = (int)prm.C + ((n == 0 && prm.out_count) ? 1 : 0);
            for (int c = 0; c < passes; ++c) {
                const bool count_pass = c == (int)prm.C;
                double acc[kGroups][4];
#pragma unroll
                for (int i = 0; i < kGroups; ++i) {
                    const int64_t vox = ((int64_t)(x_lo + gdx[i]) * prm.Y + gy[i]) * prm.Z + gz[i];
                    acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.0;
                    if (readback && gin[i]) {
                        if (zr) {
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                if (gz[i] + q < prm.Z)
                                    acc[i][q] = count_pass ? prm.out_count[vox + q] : (double)out[c * vol + vox + q];
                        } else if (count_pass) {
                            read4(prm.out_count + vox, acc[i]);
                        } else {
                            read4(out + c * vol + vox, acc[i]);
                        }
                    }
                }
                for (int k = 0; k < total; ++k) {
                    const int4 e = s_list[k];
                    if (!(x_lo < e.x + prm.p0 && x_hi > e.x)) continue;             // not over this box
                    const int flag = s_flag[k];
                    const bool aligned = flag & 1;
                    const bool covers = (flag & 2) && e.x <= x_lo && e.x + prm.p0 >= x_lo + kBX;   // the whole box
                    int stage = 0;
                    if (aligned && !count_pass) {
                        stage = it % SR::kStages;
                        mbar_wait(full_bar + stage, (it / SR::kStages) & 1);
                        ++it;
                    }
                    const uint32_t src = ring + stage * SR::kBytes + lds_off;
                    if (WMODE == 2 && covers && aligned) {
                        // whole box inside the patch, separable weight: w = fl(fl(wx * wy) * wz), then the two
                        // roundings of numpy's `sum += w * patch` (no contraction into an FMA)
#pragma unroll
                        for (int i = 0; i < kGroups; ++i) {
                            const int lx = x_lo + gdx[i] - e.x, ly = gy[i] - e.y, lz = gz[i] - e.z;
                            const double wxy = __dmul_rn(s_wx[lx & (kSepMax - 1)], s_wy[ly & (kSepMax - 1)]);
                            const double2 za = *reinterpret_cast<const double2*>(&s_wz[lz & (kSepMax - 1)]);
                            const double2 zb = *reinterpret_cast<const double2*>(&s_wz[(lz & (kSepMax - 1)) + 2]);
                            const double wq[4] = {__dmul_rn(wxy, za.x), __dmul_rn(wxy, za.y), __dmul_rn(wxy, zb.x),
                                                  __dmul_rn(wxy, zb.y)};
                            if (count_pass) {
#pragma unroll
                                for (int q = 0; q < 4; ++q) acc[i][q] += wq[q];
                            } else {
                                double v[4];
                                lds_group<TP>(src + kThreads * i * 4 * (int)sizeof(TP), v);
#pragma unroll
                                for (int q = 0; q < 4; ++q) acc[i][q] = __dadd_rn(acc[i][q], __dmul_rn(wq[q], v[q]));
                            }
                        }
                    } else if (covers && aligned && !WEIGHTED) {
                        // the usual case (patch grids aligned to the boxes): every voxel of the box is inside
                        if (count_pass) {
#pragma unroll
                            for (int i = 0; i < kGroups; ++i) {
#pragma unroll
                                for (int q = 0; q < 4; ++q) acc[i][q] += 1.0;
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < kGroups; ++i) {
                                double v[4];
                                lds_group<TP>(src + kThreads * i * 4 * (int)sizeof(TP), v);
#pragma unroll
                                for (int q = 0; q < 4; ++q) acc[i][q] += v[q];
                            }
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < kGroups; ++i) {
                            const int lx = x_lo + gdx[i] - e.x, ly = gy[i] - e.y, lz = gz[i] - e.z;
                            const bool in_xy = gin[i] && lx >= 0 && lx < prm.p0 && ly >= 0 && ly < prm.p1;
                            const int64_t local = ((int64_t)lx * prm.p1 + ly) * prm.p2 + lz;
                            if (aligned) {
                                const bool in = in_xy && lz >= 0 && lz < prm.p2;         // whole group in or out
                                double v[4] = {1.0, 1.0, 1.0, 1.0};
                                if (!count_pass) lds_group<TP>(src + kThreads * i * 4 * (int)sizeof(TP), v);
                                if (in) {
                                    if (WEIGHTED) {
                                        double w[4];
                                        if constexpr (WMODE == 2) {
                                            const double wxy = __dmul_rn(s_wx[lx], s_wy[ly]);
#pragma unroll
                                            for (int q = 0; q < 4; ++q) w[q] = __dmul_rn(wxy, s_wz[lz + q]);
                                        } else {
                                            Raw4<double>::widen(Raw4<double>::load(prm.weight + local), w);
                                        }
#pragma unroll
                                        for (int q = 0; q < 4; ++q)
                                            acc[i][q] = count_pass ? acc[i][q] + w[q] : __dadd_rn(acc[i][q], __dmul_rn(w[q], v[q]));
                                    } else {
#pragma unroll
                                        for (int q = 0; q < 4; ++q) acc[i][q] += v[q];
                                    }
                                }
                            } else if (in_xy) {
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    if (lz + q < 0 || lz + q >= prm.p2) continue;
                                    const double w = WMODE == 2 ? __dmul_rn(__dmul_rn(s_wx[lx], s_wy[ly]), s_wz[lz + q])
                                                   : WMODE == 1 ? __ldg(prm.weight + local + q) : 1.0;
                                    const double v = count_pass ? 1.0 : (double)In<TP>::load_one(pin + (int64_t)e.w * prm.stride_p + c * pvol + local + q);
                                    acc[i][q] = !WEIGHTED ? acc[i][q] + v : count_pass ? acc[i][q] + w : __dadd_rn(acc[i][q], __dmul_rn(w, v));
                                }
                            }
                        }
                    }
                    if (aligned && !count_pass) {
                        // hand the slot back once the loads above have returned: the arrive sits behind
                        // arithmetic that consumed them (the same ordering as K1's ring) and a warp barrier
                        asm volatile("" ::"d"(acc[0][0]), "d"(acc[kGroups - 1][3]) : "memory");
                        __syncwarp();
                        if (arrives(lane)) mbar_arrive(empty_bar + stage);
                    }
                }
#pragma unroll
                for (int i = 0; i < kGroups; ++i) {
                    if (!gin[i]) continue;
                    const int64_t vox = ((int64_t)(x_lo + gdx[i]) * prm.Y + gy[i]) * prm.Z + gz[i];
                    if (zr) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if (gz[i] + q >= prm.Z) continue;
                            if (count_pass) prm.out_count[vox + q] = acc[i][q];
                            else out[c * vol + vox + q] = (TO)acc[i][q];
                        }
                    } else if (count_pass) {
                        store4<double>(prm.out_count + vox, acc[i]);
                    } else {
                        store4<TO>(out + c * vol + vox, acc[i]);
                    }
                }
            }
        }
        __syncthreads();
    }
    if (tid == 0) {   // all copies consumed, all threads past the last block barrier: see mbar_inval
#pragma unroll
        for (int s = 0; s < SR::kStages; ++s) { mbar_inval(full_bar + s); mbar_inval(empty_bar + s); }
    }
}

// patches [rows = (sample, patch)][C][p0][p1][p2] as a 5-D tensor map with a [1][1][kBX][kBY][kBZ] box
template <typename TP>
static int make_patch_tensor(const StitchParams& prm, CUtensorMap* tm) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return set_error(VALUES_ERR_CUDA, "stitch: cuTensorMapEncodeTiled is not available");
    const cuuint64_t es = sizeof(TP);
    const cuuint64_t pvol = (cuuint64_t)prm.p0 * prm.p1 * prm.p2;
    // (sample, patch) rows behind the pointer: exact when the patches are taken in order (rows 0 .. n_sel - 1
    // of every sample) or the samples are whole multiples of a row apart; with a patch_index on a single
    // sample the row count is not known here: as many rows as the allocation holds
    const cuuint64_t rps = prm.N > 1 ? (cuuint64_t)(prm.stride_n / prm.stride_p) : 0;
    cuuint64_t rows = !prm.patch_index ? (cuuint64_t)(prm.N - 1) * rps + (cuuint64_t)prm.n_sel
                    : prm.N > 1        ? (cuuint64_t)prm.N * rps
                                       : 0x7fffffffull;
    // ... and never past the end of the allocation the patches live in (see bytes_to_allocation_end)
    const size_t room = bytes_to_allocation_end(prm.patches);
    if (room >= pvol * prm.C * es) {
        const cuuint64_t fit = (room - pvol * prm.C * es) / ((cuuint64_t)prm.stride_p * es) + 1;
        if (rows == 0 || fit < rows) rows = fit;
    }
    const cuuint64_t gdim[5] = {(cuuint64_t)prm.p2, (cuuint64_t)prm.p1, (cuuint64_t)prm.p0, (cuuint64_t)prm.C,
                                rows > 0 && rows < 0x7fffffffull ? rows : 0x7fffffffull};
    const cuuint64_t gstr[4] = {(cuuint64_t)prm.p2 * es, (cuuint64_t)prm.p1 * prm.p2 * es, pvol * es,
                                (cuuint64_t)prm.stride_p * es};
    const cuuint32_t box[5] = {kBZ, kBY, kBX, 1, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const CUtensorMapDataType dt = sizeof(TP) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64
                                 : sizeof(TP) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    const CUresult r = enc(tm, dt, 5, const_cast<void*>(prm.patches), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(VALUES_ERR_CUDA, "stitch: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return VALUES_OK;
}

template <typename TP, typename TO>
static int launch_stitch_box(StitchParams prm, cudaStream_t st) {
    using SR = StitchRing<TP>;
    CUtensorMap tm;
    int rc = make_patch_tensor<TP>(prm, &tm);
    if (rc) return rc;
    prm.tiles_z = (int)ceil_div(prm.Z, kBZ);
    prm.tiles_y = (int)ceil_div(prm.Y, kBY);
    const int64_t tiles = (int64_t)prm.tiles_z * prm.tiles_y * ceil_div(prm.X, kBX * kXSteps);
    if (tiles > 0x7fffffffLL || prm.N > 65535) return set_error(VALUES_ERR_UNSUPPORTED, "stitch: volume too large");
    const int rows_per_sample = prm.N > 1 ? (int)(prm.stride_n / prm.stride_p) : 0;
    const dim3 grid((unsigned)tiles, (unsigned)prm.N);
    auto launch = [&](auto kern) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SR::smem) != cudaSuccess)
            return set_error(VALUES_ERR_CUDA, "stitch: cudaFuncSetAttribute(%zu) failed", SR::smem);
        kern<<<grid, kThreads + 32, SR::smem, st>>>(tm, prm, rows_per_sample);
        return check_launch("stitch_box_kernel");
    };
    if (prm.wsep[0]) return launch(stitch_box_kernel<TP, TO, 2>);
    return prm.weight ? launch(stitch_box_kernel<TP, TO, 1>) : launch(stitch_box_kernel<TP, TO, 0>);
}

// can the box kernel fetch these patches by tensor-map copies?
static bool stitch_box_ok(const StitchParams& prm, int patch_dtype, int out_dtype, size_t pes, size_t oes) {
    const int al = pes == 2 ? 8 : 4;
    const bool rows_mergeable = prm.N == 1 || (prm.stride_p > 0 && prm.stride_n % prm.stride_p == 0 &&
                                               prm.stride_n / prm.stride_p < 0x7fffffffLL / 65536);
    // the sums leave as 4-voxel vectors when the volume's rows are whole groups of four (Z % 4 == 0), else
    // element by element: only then do the output pointers need vector alignment
    const size_t ovec = prm.Z % 4 == 0 ? 4 : 1;
    return prm.p2 % al == 0 && rows_mergeable && (prm.stride_p * (int64_t)pes) % 16 == 0 &&
           (prm.stride_p * (int64_t)pes) < (1LL << 40) && (reinterpret_cast<uintptr_t>(prm.patches) % 16) == 0 &&
           (reinterpret_cast<uintptr_t>(prm.out_sum) % (ovec * oes)) == 0 &&
           (!prm.out_count || reinterpret_cast<uintptr_t>(prm.out_count) % (ovec * 8) == 0) &&
           (!prm.weight || reinterpret_cast<uintptr_t>(prm.weight) % 32 == 0) &&
           prm.X < 0x7fffffffLL && prm.Y < 0x7fffffffLL && prm.Z < 0x7fffffffLL && prm.C < 0x7fffffffLL &&
           (patch_dtype != VALUES_F64 || out_dtype == VALUES_F64) && encode_tiled_fn() != nullptr;
}

}  // namespace vb

using namespace vb;

static int stitch_dispatch(const void* patches, int patch_dtype, int64_t patch_stride_n, int64_t patch_stride_p,
                           const int32_t* patch_index, const int32_t* crop_lo, const double* weight,
                           const double* const* wsep, int64_t n_sel, int64_t N, int64_t C,
                           const int64_t* patch3_host, const int64_t* vol3_host, void* out_sum, int out_dtype,
                           double* out_count, int accumulate, int path, void* stream) {
    if (!patches || !crop_lo || !patch3_host || !vol3_host || !out_sum)
        return set_error(VALUES_ERR_INVALID_ARG, "stitch: NULL pointer");
    if (path < 0 || path > 2) return set_error(VALUES_ERR_INVALID_ARG, "stitch: unknown path %d (0, 1, 2)", path);
    if (n_sel < 0 || N <= 0 || C <= 0)
        return set_error(VALUES_ERR_INVALID_ARG, "stitch: bad sizes");
    for (int d = 0; d < 3; ++d)
        if (patch3_host[d] <= 0 || vol3_host[d] <= 0 || patch3_host[d] > 0x7fffffff)
            return set_error(VALUES_ERR_INVALID_ARG, "stitch: bad patch/volume shape");
    if (N > 65535) return set_error(VALUES_ERR_UNSUPPORTED, "stitch: N > 65535");
    if (n_sel == 0 && accumulate) return VALUES_OK;
    if (n_sel == 0) {  // nothing covers anything: zeros (uncovered voxels stay 0 in the reference)
        const size_t vol = (size_t)vol3_host[0] * vol3_host[1] * vol3_host[2];
        const size_t es = out_dtype == VALUES_F64 ? 8 : 4;
        if (cudaMemsetAsync(out_sum, 0, (size_t)N * C * vol * es, (cudaStream_t)stream) != cudaSuccess ||
            (out_count && cudaMemsetAsync(out_count, 0, vol * 8, (cudaStream_t)stream) != cudaSuccess))
            return set_error(VALUES_ERR_CUDA, "stitch: memset failed");
        return VALUES_OK;
    }
    StitchParams prm{};
    prm.patches = patches; prm.stride_n = patch_stride_n; prm.stride_p = patch_stride_p;
    prm.patch_index = patch_index; prm.crop_lo = crop_lo; prm.weight = weight;
    prm.n_sel = n_sel; prm.N = N; prm.C = C;
    prm.p0 = (int)patch3_host[0]; prm.p1 = (int)patch3_host[1]; prm.p2 = (int)patch3_host[2];
    prm.X = vol3_host[0]; prm.Y = vol3_host[1]; prm.Z = vol3_host[2];
    prm.out_sum = out_sum; prm.out_count = out_count; prm.accumulate = accumulate ? 1 : 0;
    cudaStream_t st = (cudaStream_t)stream;
    // vector kernel: 4 consecutive z voxels per thread (16-byte loads / stores)
    const size_t pes = patch_dtype == VALUES_F64 ? 8 : (patch_dtype == VALUES_F32 ? 4 : 2);
    const size_t oes = out_dtype == VALUES_F64 ? 8 : 4;
    if (wsep) {
        // the factors live in the shared memory of the box kernel: anything it cannot take is the caller's
        // to materialise (weight[x][y][z] = fl(fl(wx[x] * wy[y]) * wz[z])) and pass as a map
        for (int d = 0; d < 3; ++d) prm.wsep[d] = wsep[d];
        if (path != 0 || patch3_host[0] > kSepMax || patch3_host[1] > kSepMax || patch3_host[2] > kSepMax ||
            !stitch_box_ok(prm, patch_dtype, out_dtype, pes, oes))
            return set_error(VALUES_ERR_UNSUPPORTED, "stitch: separable weights need the box kernel (path 0, "
                             "16-byte aligned rows, patch edges <= %d)", kSepMax);
    }
    if (path == 0 && stitch_box_ok(prm, patch_dtype, out_dtype, pes, oes)) {
        // default: output boxes fed by tensor-map copies through a shared-memory ring
        if (out_dtype == VALUES_F64) {
            if (patch_dtype == VALUES_F64) return launch_stitch_box<double, double>(prm, st);
            if (patch_dtype == VALUES_F32) return launch_stitch_box<float, double>(prm, st);
            if (patch_dtype == VALUES_BF16) return launch_stitch_box<__nv_bfloat16, double>(prm, st);
            return set_error(VALUES_ERR_INVALID_ARG, "stitch: unknown patch dtype");
        }
        if (out_dtype == VALUES_F32) {
            if (patch_dtype == VALUES_F32) return launch_stitch_box<float, float>(prm, st);
            if (patch_dtype == VALUES_BF16) return launch_stitch_box<__nv_bfloat16, float>(prm, st);
            return set_error(VALUES_ERR_INVALID_ARG, "stitch: f32 output needs f32/bf16 patches");
        }
        return set_error(VALUES_ERR_INVALID_ARG, "stitch: out dtype must be f64 or f32");
    }
    const bool vec_ok = path != 1 && prm.Z % 4 == 0 && prm.p2 % 4 == 0 &&
                        (reinterpret_cast<uintptr_t>(patches) % (4 * pes)) == 0 &&
                        (patch_stride_n % 4) == 0 && (patch_stride_p % 4) == 0 &&
                        (reinterpret_cast<uintptr_t>(out_sum) % (4 * oes)) == 0 &&
                        (!out_count || reinterpret_cast<uintptr_t>(out_count) % 32 == 0) &&
                        (!weight || reinterpret_cast<uintptr_t>(weight) % 32 == 0);
    if (vec_ok) {
        // tile extent along z: short tiles keep the overlap list short (every thread walks all of it)
        constexpr int tzv = 16;
        prm.tiles_z = (int)ceil_div(prm.Z, tzv * 4);
        prm.tiles_y = (int)ceil_div(prm.Y, kThreads / tzv);
        const int64_t vtiles = (int64_t)prm.tiles_z * prm.tiles_y * ceil_div(prm.X, kXB);
        if (vtiles > 0x7fffffffLL) return set_error(VALUES_ERR_UNSUPPORTED, "stitch: volume too large");
        const dim3 vgrid((unsigned)vtiles, (unsigned)N);
#define VB_STITCHV2(TP, TO, TZV) do { if (weight) stitch_vec_kernel<TP, TO, TZV, true><<<vgrid, kThreads, 0, st>>>(prm); \
        else stitch_vec_kernel<TP, TO, TZV, false><<<vgrid, kThreads, 0, st>>>(prm); } while (0)
#define VB_STITCHV3(TP, TO) VB_STITCHV2(TP, TO, tzv)
        if (out_dtype == VALUES_F64) {
            if (patch_dtype == VALUES_F64) VB_STITCHV3(double, double);
            else if (patch_dtype == VALUES_F32) VB_STITCHV3(float, double);
            else if (patch_dtype == VALUES_BF16) VB_STITCHV3(__nv_bfloat16, double);
            else return set_error(VALUES_ERR_INVALID_ARG, "stitch: unknown patch dtype");
        } else if (out_dtype == VALUES_F32) {
            if (patch_dtype == VALUES_F32) VB_STITCHV3(float, float);
            else if (patch_dtype == VALUES_BF16) VB_STITCHV3(__nv_bfloat16, float);
            else return set_error(VALUES_ERR_INVALID_ARG, "stitch: f32 output needs f32/bf16 patches");
        } else {
            return set_error(VALUES_ERR_INVALID_ARG, "stitch: out dtype must be f64 or f32");
        }
#undef VB_STITCHV3
#undef VB_STITCHV2
        return check_launch("stitch_vec_kernel");
    }
    prm.tiles_z = (int)ceil_div(prm.Z, kSZ);
    prm.tiles_y = (int)ceil_div(prm.Y, kSY);
    const int64_t tiles = (int64_t)prm.tiles_z * prm.tiles_y * prm.X;
    if (tiles > 0x7fffffffLL) return set_error(VALUES_ERR_UNSUPPORTED, "stitch: volume too large");
    const dim3 grid((unsigned)tiles, (unsigned)N);
#define VB_STITCH(TP, TO) stitch_kernel<TP, TO><<<grid, kThreads, 0, st>>>(prm)
    if (out_dtype == VALUES_F64) {
        if (patch_dtype == VALUES_F64) VB_STITCH(double, double);
        else if (patch_dtype == VALUES_F32) VB_STITCH(float, double);
        else if (patch_dtype == VALUES_BF16) VB_STITCH(__nv_bfloat16, double);
        else return set_error(VALUES_ERR_INVALID_ARG, "stitch: unknown patch dtype");
    } else if (out_dtype == VALUES_F32) {
        if (patch_dtype == VALUES_F32) VB_STITCH(float, float);
        else if (patch_dtype == VALUES_BF16) VB_STITCH(__nv_bfloat16, float);
        else return set_error(VALUES_ERR_INVALID_ARG, "stitch: f32 output needs f32/bf16 patches");
    } else {
        return set_error(VALUES_ERR_INVALID_ARG, "stitch: out dtype must be f64 or f32");
    }
#undef VB_STITCH
    return check_launch("stitch_kernel");
}

extern "C" int values_stitch_accumulate_weighted(const void* patches, int patch_dtype,
                                                 int64_t patch_stride_n, int64_t patch_stride_p,
                                                 const int32_t* patch_index, const int32_t* crop_lo,
                                                 const double* weight, int64_t n_sel, int64_t N,
                                                 int64_t C, const int64_t* patch3_host,
                                                 const int64_t* vol3_host, void* out_sum,
                                                 int out_dtype, double* out_count, int accumulate,
                                                 int path, void* stream) {
    return stitch_dispatch(patches, patch_dtype, patch_stride_n, patch_stride_p, patch_index, crop_lo, weight,
                           nullptr, n_sel, N, C, patch3_host, vol3_host, out_sum, out_dtype, out_count,
                           accumulate, path, stream);
}

extern "C" int values_stitch_accumulate_separable(const void* patches, int patch_dtype,
                                                  int64_t patch_stride_n, int64_t patch_stride_p,
                                                  const int32_t* patch_index, const int32_t* crop_lo,
                                                  const double* wx, const double* wy, const double* wz,
                                                  int64_t n_sel, int64_t N, int64_t C,
                                                  const int64_t* patch3_host, const int64_t* vol3_host,
                                                  void* out_sum, int out_dtype, double* out_count,
                                                  int accumulate, int path, void* stream) {
    if (!wx || !wy || !wz) return set_error(VALUES_ERR_INVALID_ARG, "stitch: NULL weight factor");
    const double* wsep[3] = {wx, wy, wz};
    return stitch_dispatch(patches, patch_dtype, patch_stride_n, patch_stride_p, patch_index, crop_lo, nullptr,
                           wsep, n_sel, N, C, patch3_host, vol3_host, out_sum, out_dtype, out_count,
                           accumulate, path, stream);
}

extern "C" int values_stitch_accumulate(const void* patches, int patch_dtype,
                                        int64_t patch_stride_n, int64_t patch_stride_p,
                                        const int32_t* patch_index, const int32_t* crop_lo,
                                        int64_t n_sel, int64_t N, int64_t C,
                                        const int64_t* patch3_host, const int64_t* vol3_host,
                                        void* out_sum, int out_dtype, double* out_count,
                                        int accumulate, int path, void* stream) {
    return values_stitch_accumulate_weighted(patches, patch_dtype, patch_stride_n, patch_stride_p,
                                             patch_index, crop_lo, nullptr, n_sel, N, C, patch3_host,
                                             vol3_host, out_sum, out_dtype, out_count, accumulate,
                                             path, stream);
}

