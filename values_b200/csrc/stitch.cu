// K3: sliding-window stitch accumulator -- atomic-free, output-stationary.
//
// The reference (DataCarrier3D.concat_data, uncertainty_modeling/data_carrier_3D.py:154-179)
// does one device->host copy and one strided numpy `+=` per (patch, sample).  Here every
// CTA owns a tile of the OUTPUT volume, finds the patches that overlap it (ordered
// compaction, so the summation order is the reference's list order), sums them in registers
// and writes each output voxel exactly once.  No atomics, bit-reproducible.
// Algorithmic bytes: every patch element read once + every output element written once.
#include "common.cuh"

namespace vb {

constexpr int kSZ = 64;  // tile extent along the contiguous axis (z)
constexpr int kSY = 4;   // tile extent along axis 1 (y); kSZ*kSY == kThreads
constexpr int kMaxList = 1024;  // overlapping patches kept per tile chunk

struct StitchParams {
    const void* patches;
    int64_t stride_n, stride_p;
    const int32_t* patch_index;
    const int32_t* crop_lo;
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
    double cnt = 0.0;
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
                        acc += (double)In<TP>::load_one(pin + pi * prm.stride_p + c * pvol + local);
                        if (c == 0) cnt += 1.0;
                    }
                }
                *dst = (TO)acc;
            }
        }
        __syncthreads();
    }
    if (inside && n == 0 && prm.out_count) {
        double* d = prm.out_count + vox;
        *d = prm.accumulate ? *d + cnt : cnt;
    }
}

}  // namespace vb

using namespace vb;

extern "C" int values_stitch_accumulate(const void* patches, int patch_dtype,
                                        int64_t patch_stride_n, int64_t patch_stride_p,
                                        const int32_t* patch_index, const int32_t* crop_lo,
                                        int64_t n_sel, int64_t N, int64_t C,
                                        const int64_t* patch3_host, const int64_t* vol3_host,
                                        void* out_sum, int out_dtype, double* out_count,
                                        int accumulate, void* stream) {
    if (!patches || !crop_lo || !patch3_host || !vol3_host || !out_sum)
        return set_error(VALUES_ERR_INVALID_ARG, "stitch: NULL pointer");
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
    prm.patch_index = patch_index; prm.crop_lo = crop_lo;
    prm.n_sel = n_sel; prm.N = N; prm.C = C;
    prm.p0 = (int)patch3_host[0]; prm.p1 = (int)patch3_host[1]; prm.p2 = (int)patch3_host[2];
    prm.X = vol3_host[0]; prm.Y = vol3_host[1]; prm.Z = vol3_host[2];
    prm.out_sum = out_sum; prm.out_count = out_count; prm.accumulate = accumulate ? 1 : 0;
    prm.tiles_z = (int)ceil_div(prm.Z, kSZ);
    prm.tiles_y = (int)ceil_div(prm.Y, kSY);
    const int64_t tiles = (int64_t)prm.tiles_z * prm.tiles_y * prm.X;
    if (tiles > 0x7fffffffLL) return set_error(VALUES_ERR_UNSUPPORTED, "stitch: volume too large");
    const dim3 grid((unsigned)tiles, (unsigned)N);
    cudaStream_t st = (cudaStream_t)stream;
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
