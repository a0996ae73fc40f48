// mfnerf_b200 -- shared device/host helpers for the sm_100a hot path.
// Everything here is written from scratch; reference behaviour is cited as
// (ref: models/csrc/<file>:<lines>) so the parity tests can be audited.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#define MFN_OK 0
#define MFN_ERR_CUDA (-1)
#define MFN_ERR_ARG (-2)
#define MFN_ERR_CAPACITY (-3)

namespace mfn {

void set_error(const char* fmt, ...);
int check_launch(const char* what, cudaStream_t stream);   // also counts one kernel launch
void note_launch(int k);                                    // extra launches of a multi-kernel entry point
// Brackets a kernel launch with the CUDA events registered through mfn_profile_set() when `name` matches; works
// inside stream capture (event-record nodes), so a kernel can be timed inside a replayed CUDA graph.
struct ProfScope {
    cudaStream_t st; int slot;
    ProfScope(const char* name, cudaStream_t stream);
    ~ProfScope();
};

constexpr int kWarp = 32;
constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

__host__ __device__ __forceinline__ int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- morton codes (ref: raymarching.cu:35-60).  10 bits/axis, x | y<<1 | z<<2. ----
// The reference expands with multiplies (v*0x00010001 & 0xFF0000FF ...).  For grid coords (< 1024)
// that equals shift-or; for arbitrary 32-bit inputs the carries differ, so the multiply form is
// kept to stay bit-exact with the reference on any input.
__host__ __device__ __forceinline__ uint32_t spread3_mul(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__host__ __device__ __forceinline__ uint32_t morton_encode(uint32_t x, uint32_t y, uint32_t z) {
    return spread3_mul(x) | (spread3_mul(y) << 1) | (spread3_mul(z) << 2);
}
__host__ __device__ __forceinline__ uint32_t compact3(uint32_t x) {
    x &= 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}

// uniform integer in [0, n), n < 2^32, from the high 32 bits of a 64-bit random word (32 x 32 -> 64-bit product: cannot overflow)
__host__ __device__ __forceinline__ uint32_t uniform_below(uint64_t r, uint32_t n) { return (uint32_t)(((r >> 32) * (uint64_t)n) >> 32); }

// ---- warp helpers ----
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_incl_scan(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    return v;
}
__device__ __forceinline__ int warp_incl_scan_i(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    return v;
}

}  // namespace mfn
