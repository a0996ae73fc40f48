// Warp-level tensor-core building blocks (ldmatrix + mma.sync m16n8k16, fp16 inputs, fp32 accumulate) used by
// the v1 MLP kernels.  The fragment/addressing conventions are spelled out once here.
#pragma once
#include "common.cuh"

namespace mfn {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// D(16x8,f32) += A(16x16,f16,row) * B(16x8,f16,col)
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

// ---- operand loaders.  `ld` = row stride of the smem matrix in halfs (padded so that 8 consecutive rows fall in
// distinct 16-byte bank groups).  lane = threadIdx.x & 31.
// A fragment (rows m0..m0+15, k0..k0+15) from a matrix stored [m][k]
__device__ __forceinline__ void load_a(uint32_t (&a)[4], const __half* s, int ld, int m0, int k0, int lane) {
    const int r = m0 + (lane & 7) + ((lane >> 3) & 1) * 8, c = k0 + (lane >> 4) * 8;
    ldsm_x4(a, smem_u32(s + r * ld + c));
}
// A fragment from a matrix stored transposed, [k][m]
__device__ __forceinline__ void load_a_t(uint32_t (&a)[4], const __half* s, int ld, int m0, int k0, int lane) {
    const int k = k0 + (lane & 7) + (lane >> 4) * 8, m = m0 + ((lane >> 3) & 1) * 8;
    ldsm_x4_t(a, smem_u32(s + k * ld + m));
}
// B fragments of two adjacent n8 tiles (n0..n0+15, k0..k0+15) from a matrix stored [n][k]: b[0],b[1] -> tile n0, b[2],b[3] -> tile n0+8
__device__ __forceinline__ void load_b(uint32_t (&b)[4], const __half* s, int ld, int n0, int k0, int lane) {
    const int n = n0 + (lane & 7) + (lane >> 4) * 8, k = k0 + ((lane >> 3) & 1) * 8;
    ldsm_x4(b, smem_u32(s + n * ld + k));
}
// same from a matrix stored [k][n]
__device__ __forceinline__ void load_b_t(uint32_t (&b)[4], const __half* s, int ld, int n0, int k0, int lane) {
    const int k = k0 + (lane & 7) + ((lane >> 3) & 1) * 8, n = n0 + (lane >> 4) * 8;
    ldsm_x4_t(b, smem_u32(s + k * ld + n));
}

}  // namespace mfn
