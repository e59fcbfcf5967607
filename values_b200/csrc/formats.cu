// Axis-order conversion between the on-disk layout of the reference's hand-off files and the
// in-memory layout its Python code indexes.
//
// The reference writes its maps with medpy.io.save and reads them back with medpy.io.load
// (data_carrier_3D.py:233-371, experiment_dataloader.py:38-49, aggregate_uncertainties.py:77-79).
// A NIfTI payload is stored x-fastest, i.e. as a C-order array [Z][Y][X]; medpy hands Python an
// array indexed [x][y][z].  On the host that is a strided view; the reference then makes strided
// passes over it.  Here the raw payload is uploaded as it lies in the file and one HBM-bound
// kernel reverses the axis order on the device (and back before a save): algorithmic bytes =
// 2 * elem_bytes per element, 32x32 tiles through shared memory, both sides coalesced.
// The same reversal turns cv2's [H][W] image into medpy's [W][H] (n1 == 1).
#include "common.cuh"

namespace vb {

template <typename T>
__global__ void __launch_bounds__(256) reverse_axes_kernel(const T* __restrict__ in, T* __restrict__ out,
                                                           int64_t n0, int64_t n1, int64_t n2) {
    __shared__ T tile[32][33];
    const int64_t c0 = (int64_t)blockIdx.x * 32, a0 = (int64_t)blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int64_t b = blockIdx.z; b < n1; b += gridDim.z) {
#pragma unroll
        for (int r = 0; r < 32; r += 8) {
            const int64_t a = a0 + ty + r, c = c0 + tx;
            if (a < n0 && c < n2) tile[ty + r][tx] = in[(a * n1 + b) * n2 + c];
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 32; r += 8) {
            const int64_t c = c0 + ty + r, a = a0 + tx;
            if (a < n0 && c < n2) out[(c * n1 + b) * n0 + a] = tile[tx][ty + r];
        }
        __syncthreads();
    }
}

template <typename T>
static int launch_reverse(const void* in, void* out, int64_t n0, int64_t n1, int64_t n2, cudaStream_t st) {
    const int64_t gx = ceil_div(n2, 32), gy = ceil_div(n0, 32);
    if (gx > 0x7fffffffLL || gy > 65535) return set_error(VALUES_ERR_UNSUPPORTED, "reverse_axes: grid too large");
    const dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)std::min<int64_t>(n1, 65535));
    reverse_axes_kernel<T><<<grid, 256, 0, st>>>(reinterpret_cast<const T*>(in), reinterpret_cast<T*>(out), n0, n1, n2);
    return check_launch("reverse_axes_kernel");
}

}  // namespace vb

using namespace vb;

extern "C" int values_reverse_axes(const void* in, void* out, int elem_bytes, int64_t n0, int64_t n1,
                                   int64_t n2, void* stream) {
    if (n0 < 0 || n1 < 0 || n2 < 0) return set_error(VALUES_ERR_INVALID_ARG, "reverse_axes: negative size");
    if (n0 == 0 || n1 == 0 || n2 == 0) return VALUES_OK;
    if (!in || !out || in == out) return set_error(VALUES_ERR_INVALID_ARG, "reverse_axes: NULL or aliased pointers");
    cudaStream_t st = (cudaStream_t)stream;
    switch (elem_bytes) {
        case 1: return launch_reverse<uint8_t>(in, out, n0, n1, n2, st);
        case 2: return launch_reverse<uint16_t>(in, out, n0, n1, n2, st);
        case 4: return launch_reverse<uint32_t>(in, out, n0, n1, n2, st);
        case 8: return launch_reverse<uint64_t>(in, out, n0, n1, n2, st);
        default: return set_error(VALUES_ERR_INVALID_ARG, "reverse_axes: elem_bytes must be 1, 2, 4 or 8");
    }
}
