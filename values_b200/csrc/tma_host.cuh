// Host side of the tensor-map (TMA) copies: cuTensorMapEncodeTiled through the runtime's driver entry
// point table, so the library does not link against libcuda.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace vb {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}

// Bytes from `ptr` to the end of the device allocation that contains it (cuMemGetAddressRange), or 0 when
// the driver cannot say.  A tensor map's declared extent is kept inside this range: the copies never touch
// what lies beyond the caller's rows either way, but an extent that runs past the allocation is something
// compute-sanitizer's synccheck / racecheck instrumentation of the tensor-map copies does not survive
// (r02: "illegal memory access" with 0 errors reported on a 2^31-row map over a 737 KB tensor).
typedef CUresult (*AddressRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
// [base, base + size) of the device allocation that contains `ptr`; false when the driver cannot say
static inline bool allocation_range(const void* ptr, uintptr_t* base_out, size_t* size_out) {
    static AddressRangeFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return reinterpret_cast<AddressRangeFn>(f);
    }();
    if (!fn) return false;
    CUdeviceptr base = 0;
    size_t size = 0;
    const CUdeviceptr p = reinterpret_cast<CUdeviceptr>(ptr);
    if (fn(&base, &size, p) != CUDA_SUCCESS || p < base || p - base >= size) return false;
    *base_out = (uintptr_t)base;
    *size_out = size;
    return true;
}
static inline size_t bytes_to_allocation_end(const void* ptr) {
    uintptr_t base = 0;
    size_t size = 0;
    if (!allocation_range(ptr, &base, &size)) return 0;
    return size - (size_t)(reinterpret_cast<uintptr_t>(ptr) - base);
}

}  // namespace vb
