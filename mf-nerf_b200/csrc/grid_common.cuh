// Device-side pieces of the multiresolution hash-grid encoding shared by the stand-alone encoder kernels (encoder.cu) and the
// fused field kernels (field_fused.cu).  Algorithm: see the header comment of encoder.cu.
#pragma once
#include "field_internal.h"

namespace mfn {

__device__ __forceinline__ uint32_t grid_index(uint32_t x, uint32_t y, uint32_t z, uint32_t res, uint32_t size, bool hashed) {
    uint32_t idx;
    if (hashed) idx = x ^ (y * 2654435761u) ^ (z * 805459861u);
    else idx = x + y * res + z * res * res;
    return idx % size;
}

// MixedFeature index transformation: vertex v of a level sits at x = (v - 0.5) / scale_l; its canonical-grid vertex is the one
// nearest to that point, round((v - 0.5) * scale_c / scale_l + 0.5).  Separately rounded multiply and add (no FMA) so that the torch
// restatement reproduces the integer result bit for bit.  For the canonical level itself (ratio 1) this is the identity.
__device__ __forceinline__ uint32_t canon_vertex(uint32_t v, float ratio) {
    return (uint32_t)(int)floorf(__fadd_rn(__fmul_rn(__fsub_rn((float)v, 0.5f), ratio), 1.0f));
}

template <int F> struct FeatVec;
template <> struct FeatVec<1> { using T = unsigned short; };
template <> struct FeatVec<2> { using T = uint32_t; };
template <> struct FeatVec<4> { using T = uint2; };
template <> struct FeatVec<8> { using T = uint4; };

template <int F>
__device__ __forceinline__ void add_weighted(float (&acc)[F], const typename FeatVec<F>::T& raw, float w) {
    const __half* h = reinterpret_cast<const __half*>(&raw);
#pragma unroll
    for (int f = 0; f < F; ++f) acc[f] = fmaf(w, __half2float(h[f]), acc[f]);
}

// interpolate one level for one sample
template <int F>
__device__ __forceinline__ void encode_level(const __half* __restrict__ table, const GridMeta& m, int l, float x, float y, float z, float (&acc)[F]) {
    const float s = m.scale[l];
    const uint32_t res = m.res[l], size = m.size[l];
    const bool hashed = (m.hashed >> l) & 1u;
    const float px = fmaf(x, s, 0.5f), py = fmaf(y, s, 0.5f), pz = fmaf(z, s, 0.5f);
    const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
    const float wx = px - fx, wy = py - fy, wz = pz - fz;
    const uint32_t gx = (uint32_t)(int)fx, gy = (uint32_t)(int)fy, gz = (uint32_t)(int)fz;
    uint32_t cx[2] = {gx, gx + 1u}, cy[2] = {gy, gy + 1u}, cz[2] = {gz, gz + 1u};
    if (m.mixed) {
        const float r = m.canon[l];
#pragma unroll
        for (int k = 0; k < 2; ++k) { cx[k] = canon_vertex(cx[k], r); cy[k] = canon_vertex(cy[k], r); cz[k] = canon_vertex(cz[k], r); }
    }
    const typename FeatVec<F>::T* lvl = reinterpret_cast<const typename FeatVec<F>::T*>(table) + m.offset[l];
    typename FeatVec<F>::T v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c)
        v[c] = __ldg(lvl + grid_index(cx[c & 1], cy[(c >> 1) & 1], cz[c >> 2], res, size, hashed));
#pragma unroll
    for (int f = 0; f < F; ++f) acc[f] = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float w = ((c & 1) ? wx : 1.f - wx) * (((c >> 1) & 1) ? wy : 1.f - wy) * ((c >> 2) ? wz : 1.f - wz);
        add_weighted<F>(acc, v[c], w);
    }
}

// sample position in [0,1]^3; with e.normalize the world position is mapped exactly like networks.py:105,
// x = (x - xyz_min) / (xyz_max - xyz_min), in IEEE fp32
__device__ __forceinline__ void load_pos(const EncArgs& e, int64_t i, float& x, float& y, float& z) {
    x = e.x[3 * i]; y = e.x[3 * i + 1]; z = e.x[3 * i + 2];
    if (e.normalize) {
        x = __fdiv_rn(__fsub_rn(x, e.mn[0]), __fsub_rn(e.mx[0], e.mn[0]));
        y = __fdiv_rn(__fsub_rn(y, e.mn[1]), __fsub_rn(e.mx[1], e.mn[1]));
        z = __fdiv_rn(__fsub_rn(z, e.mn[2]), __fsub_rn(e.mx[2], e.mn[2]));
    }
}


}  // namespace mfn
