// morton3D / morton3D_invert / packbits  (ref: models/csrc/raymarching.cu:35-161)
// HBM-bound integer work: 16 B/elt for the morton ops, 4.125 B/cell for packbits.
#include "common.cuh"
#include "../../include/mfnerf_b200.h"

namespace mfn {

__global__ void morton3d_kernel(const int32_t* __restrict__ coords, int64_t n, int32_t* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const uint32_t x = coords[3 * i + 0], y = coords[3 * i + 1], z = coords[3 * i + 2];
        out[i] = (int32_t)morton_encode(x, y, z);
    }
}

__global__ void morton3d_invert_kernel(const int32_t* __restrict__ idx, int64_t n, int32_t* __restrict__ coords) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        // the reference shifts the *signed* int (ind >> 1) before the uint conversion (raymarching.cu:97-100)
        const int32_t v = idx[i];
        coords[3 * i + 0] = (int32_t)compact3((uint32_t)(v >> 0));
        coords[3 * i + 1] = (int32_t)compact3((uint32_t)(v >> 1));
        coords[3 * i + 2] = (int32_t)compact3((uint32_t)(v >> 2));
    }
}

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }

// one thread per output byte, 8 consecutive cells -> two 16-B loads for fp32.
// bit i of byte n  =  grid[8n+i] > thr   (ref: raymarching.cu:136-138).  The reference compares in the
// grid's own dtype for fp64; fp32/fp16 compare after promotion to float, which is what we do.
template <typename T>
__global__ void packbits_kernel(const T* __restrict__ grid, int64_t n_bytes, float thr, const float* __restrict__ thr_dev, uint8_t* __restrict__ bits) {
    if (thr_dev) thr = fminf(*thr_dev, thr);   // device-side min(mean_density, density_threshold), networks.py:270
    int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; n < n_bytes; n += stride) {
        uint32_t b = 0;
        if constexpr (sizeof(T) == 4) {
            const float4 lo = __ldg(reinterpret_cast<const float4*>(grid) + 2 * n);
            const float4 hi = __ldg(reinterpret_cast<const float4*>(grid) + 2 * n + 1);
            b |= (lo.x > thr) << 0; b |= (lo.y > thr) << 1; b |= (lo.z > thr) << 2; b |= (lo.w > thr) << 3;
            b |= (hi.x > thr) << 4; b |= (hi.y > thr) << 5; b |= (hi.z > thr) << 6; b |= (hi.w > thr) << 7;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) b |= (to_f32<T>(grid[8 * n + i]) > thr) << i;
        }
        bits[n] = (uint8_t)b;
    }
}
__global__ void packbits_f64_kernel(const double* __restrict__ grid, int64_t n_bytes, float thr, uint8_t* __restrict__ bits) {
    int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; n < n_bytes; n += stride) {
        uint32_t b = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) b |= (grid[8 * n + i] > (double)thr) << i;
        bits[n] = (uint8_t)b;
    }
}

static inline int grid_for(int64_t n, int threads) {
    int64_t blocks = ceil_div(n, threads);
    const int64_t cap = (int64_t)kNumSMs * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace mfn

using namespace mfn;

extern "C" int mfn_morton3d(const int32_t* coords, int64_t n, int32_t* indices, void* stream) {
    if (n < 0 || (n > 0 && (!coords || !indices))) { set_error("mfn_morton3d: bad argument"); return MFN_ERR_ARG; }
    if (n == 0) return MFN_OK;
    morton3d_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(coords, n, indices);
    return check_launch("mfn_morton3d", (cudaStream_t)stream);
}

extern "C" int mfn_morton3d_invert(const int32_t* indices, int64_t n, int32_t* coords, void* stream) {
    if (n < 0 || (n > 0 && (!coords || !indices))) { set_error("mfn_morton3d_invert: bad argument"); return MFN_ERR_ARG; }
    if (n == 0) return MFN_OK;
    morton3d_invert_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(indices, n, coords);
    return check_launch("mfn_morton3d_invert", (cudaStream_t)stream);
}

extern "C" int mfn_packbits(const void* grid, int dtype, int64_t n_bytes, float thr, uint8_t* bitfield, void* stream) {
    if (n_bytes < 0 || (n_bytes > 0 && (!grid || !bitfield))) { set_error("mfn_packbits: bad argument"); return MFN_ERR_ARG; }
    if (n_bytes == 0) return MFN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int g = grid_for(n_bytes, 256);
    switch (dtype) {
        case MFN_DTYPE_F32:
            if (((uintptr_t)grid & 15) != 0) { set_error("mfn_packbits: fp32 grid must be 16-byte aligned"); return MFN_ERR_ARG; }
            packbits_kernel<float><<<g, 256, 0, st>>>((const float*)grid, n_bytes, thr, nullptr, bitfield); break;
        case MFN_DTYPE_F16: packbits_kernel<__half><<<g, 256, 0, st>>>((const __half*)grid, n_bytes, thr, nullptr, bitfield); break;
        case MFN_DTYPE_F64: packbits_f64_kernel<<<g, 256, 0, st>>>((const double*)grid, n_bytes, thr, bitfield); break;
        default: set_error("mfn_packbits: unsupported dtype %d", dtype); return MFN_ERR_ARG;
    }
    return check_launch("mfn_packbits", st);
}

extern "C" int mfn_packbits_dev_thr(const float* grid, int64_t n_bytes, float max_threshold, const float* threshold_dev, uint8_t* bitfield, void* stream) {
    if (n_bytes < 0 || (n_bytes > 0 && (!grid || !bitfield || !threshold_dev))) { set_error("mfn_packbits_dev_thr: bad argument"); return MFN_ERR_ARG; }
    if (n_bytes == 0) return MFN_OK;
    if (((uintptr_t)grid & 15) != 0) { set_error("mfn_packbits_dev_thr: grid must be 16-byte aligned"); return MFN_ERR_ARG; }
    packbits_kernel<float><<<grid_for(n_bytes, 256), 256, 0, (cudaStream_t)stream>>>(grid, n_bytes, max_threshold, threshold_dev, bitfield);
    return check_launch("mfn_packbits_dev_thr", (cudaStream_t)stream);
}
