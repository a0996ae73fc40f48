// Multiresolution hash-grid encoding, forward and backward, and the degree-4 spherical-harmonics encoding.
// Replaces the tiny-cuda-nn "HashGrid" / "SphericalHarmonics" encodings the reference instantiates in
// models/networks.py:36-47,60-67.  tcnn's source is not part of the reference tree (parity unpinned);
// the arithmetic below follows tcnn's published algorithm as restated in SURVEY.md section 8c:
//   per level l: scale_l = 2^(l*log2(b)) * N_min - 1, res_l = ceil(scale_l) + 1,
//                entries_l = min(round_up(res_l^3, 8), 2^T), levels concatenated;
//   pos = x*scale_l + 0.5, cell = floor(pos), w = pos - cell, 8 corners weighted by prod(bit ? w : 1-w);
//   index = x + y*res + z*res^2 when res^3 <= entries_l, else (x*1 ^ y*2654435761 ^ z*805459861), both mod entries_l;
//   features fp16, interpolation in fp32, one rounding to fp16 on output; output order [level][feature].
// Forward: one thread per sample walks all levels (128 independent 4-byte gathers in flight per thread for L=16,F=2),
// writes its 2*L*F output bytes with 16-byte stores.  Backward: grid.y = level (concurrent CTAs hit the same level's
// table -> L2 locality), fp32 vector atomics (red.global.add.v2.f32 on sm_100) into a dense fp32 gradient table.
// Algorithmic bytes/sample (L=16,F=2): fwd 12 + 512 gathered + 64 written; bwd 12 + 64 + 512 scattered.
#include "field_internal.h"
#include "grid_common.cuh"
#include "sh4.cuh"
#include "umma.cuh"
#include <math.h>
#include <stdlib.h>

namespace mfn {

int build_grid_meta(const mfn_grid_cfg* cfg, GridMeta* m, const char* who) {
    if (!cfg || cfg->n_levels < 1 || cfg->n_levels > kMaxLevels || cfg->log2_hashmap_size < 1 || cfg->log2_hashmap_size > 30 ||
        cfg->base_resolution < 1 || !(cfg->per_level_scale > 0)) { set_error("%s: bad grid config", who); return MFN_ERR_ARG; }
    if (cfg->n_features != 1 && cfg->n_features != 2 && cfg->n_features != 4 && cfg->n_features != 8) {
        set_error("%s: n_features_per_level must be 1, 2, 4 or 8", who); return MFN_ERR_ARG;
    }
    if (cfg->grid_type != MFN_GRID_HASH && cfg->grid_type != MFN_GRID_MIXED) { set_error("%s: unsupported grid_type %d", who, cfg->grid_type); return MFN_ERR_ARG; }
    if (cfg->grid_type == MFN_GRID_MIXED && (cfg->n_tables < 1 || cfg->n_tables > cfg->n_levels)) {
        set_error("%s: MixedFeature grid needs 1 <= n_tables <= n_levels (got %d)", who, cfg->n_tables); return MFN_ERR_ARG;
    }
    // tcnn keeps per_level_scale as a float; the level scale is evaluated in double from that float and rounded to
    // float once, which makes res_l robust against last-bit noise when scale_l is (nearly) an integer.
    const double log2b = log2((double)(float)cfg->per_level_scale);
    uint64_t off = 0;
    m->hashed = 0; m->n_levels = cfg->n_levels;
    for (int l = 0; l < cfg->n_levels; ++l) {
        const float s = (float)(exp2((double)l * log2b) * (double)cfg->base_resolution - 1.0);
        const uint32_t res = (uint32_t)ceilf(s) + 1u;
        uint64_t cells = (uint64_t)res * res * res;
        const uint64_t cap = 0x7fffffffu;
        uint64_t entries = cells > cap ? cap : cells;
        entries = (entries + 7) / 8 * 8;
        const uint64_t T = 1ull << cfg->log2_hashmap_size;
        if (entries > T) entries = T;
        if (cells > entries) m->hashed |= 1u << l;
        m->offset[l] = (uint32_t)off; m->res[l] = res; m->scale[l] = s; m->size[l] = (uint32_t)entries; m->canon[l] = 0.f;
        off += entries;
        if (off > 0xffffffffull) { set_error("%s: grid too large", who); return MFN_ERR_ARG; }
    }
    m->offset[cfg->n_levels] = (uint32_t)off;
    m->mixed = 0;
    if (cfg->grid_type == MFN_GRID_MIXED) {
        // MixedFeature (the MF-NeRF fork's grid; semantics defined HERE): level l lives in table k = l * n_tables / L,
        // every table has 2^T entries and is always hashed; vertices are hashed by their coordinates in the table's canonical grid
        // (its finest level), so coincident vertices of the levels of one table share a feature
        const uint64_t T = 1ull << cfg->log2_hashmap_size;
        if (T * (uint64_t)cfg->n_tables > 0xffffffffull) { set_error("%s: grid too large", who); return MFN_ERR_ARG; }
        m->mixed = 1;
        m->hashed = cfg->n_levels >= 32 ? 0xffffffffu : ((1u << cfg->n_levels) - 1u);
        for (int l = 0; l < cfg->n_levels; ++l) {
            const int k = l * cfg->n_tables / cfg->n_levels;
            int lc = l;
            while (lc + 1 < cfg->n_levels && (lc + 1) * cfg->n_tables / cfg->n_levels == k) ++lc;
            m->offset[l] = (uint32_t)((uint64_t)k * T); m->size[l] = (uint32_t)T;
            m->canon[l] = m->scale[lc] / m->scale[l];
        }
        m->offset[cfg->n_levels] = (uint32_t)(T * (uint64_t)cfg->n_tables);
    }
    return MFN_OK;
}

template <int F>
__global__ void __launch_bounds__(128)
grid_encode_fwd_kernel(const __grid_constant__ EncArgs e, const __half* __restrict__ table, const __grid_constant__ GridMeta m,
                       __half* __restrict__ out) {
    const int64_t n = e.n_dev ? min((int64_t)*e.n_dev, e.n_max) : e.n_max;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float x, y, z;
    load_pos(e, i, x, y, z);
    constexpr int G = 8 / F;  // levels per 16-byte store
    __half* o = out + (size_t)i * m.n_levels * F;
    int l = 0;
    for (; l + G <= m.n_levels; l += G) {
        __align__(16) __half pack[8];
#pragma unroll
        for (int j = 0; j < G; ++j) {
            float acc[F];
            encode_level<F>(table, m, l + j, x, y, z, acc);
#pragma unroll
            for (int f = 0; f < F; ++f) pack[j * F + f] = __float2half_rn(acc[f]);
        }
        if (((m.n_levels * F) % 8) == 0) *reinterpret_cast<uint4*>(o + l * F) = *reinterpret_cast<const uint4*>(pack);
        else {
#pragma unroll
            for (int k = 0; k < 8; ++k) o[l * F + k] = pack[k];
        }
    }
    for (; l < m.n_levels; ++l) {
        float acc[F];
        encode_level<F>(table, m, l, x, y, z, acc);
#pragma unroll
        for (int f = 0; f < F; ++f) o[l * F + f] = __float2half_rn(acc[f]);
    }
    }
}

template <int F>
__device__ __forceinline__ void atomic_add_vec(float* p, const float (&v)[F]) {
    if constexpr (F == 2) atomicAdd(reinterpret_cast<float2*>(p), make_float2(v[0], v[1]));
    else if constexpr (F == 4) atomicAdd(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
    else {
#pragma unroll
        for (int f = 0; f < F; ++f) atomicAdd(p + f, v[f]);
    }
}

// Backward.  grid.y = level (all CTAs resident at one time work on the same level's table -> L2 locality); one thread per sample,
// so a warp holds 32 consecutive samples -- samples that follow one another along a ray.  At every level whose cell is larger
// than the marching step, consecutive samples sit in the SAME cell and address the same 8 table entries: each run of lanes with
// an identical cell is reduced to its head lane with a segmented shuffle reduction (warp-uniform early exit once no run is longer
// than the shuffle distance), and only the head issues the 8 vector atomics.  On the Lego-shaped workload this removes ~2/3 of
// the L2 reductions (measured: 213 us -> 65 us for 223k samples, tools/encbwd_bench.cu; the L2 sustains ~182 G lane-red/s).
template <int F>
__global__ void __launch_bounds__(256)
grid_encode_bwd_kernel(const __grid_constant__ EncArgs e, const __half* __restrict__ dL_dout, const __grid_constant__ GridMeta m,
                       float* __restrict__ dgrid, int32_t* __restrict__ overflow_flag) {
    const int64_t n = e.n_dev ? min((int64_t)*e.n_dev, e.n_max) : e.n_max;
    const int l = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const float s = m.scale[l];
    const uint32_t res = m.res[l], size = m.size[l];
    const bool hashed = (m.hashed >> l) & 1u;
    float* lvl = dgrid + (size_t)m.offset[l] * F;
    const int64_t n_pad = (n + 31) / 32 * 32;   // whole warps stay in the loop (shuffles below)
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += (int64_t)gridDim.x * blockDim.x) {
        float g[F];
        bool live = false, bad = false;
        if (i < n) {
#pragma unroll
            for (int f = 0; f < F; ++f) { g[f] = __half2float(dL_dout[(size_t)i * m.n_levels * F + l * F + f]); live |= (g[f] != 0.f); bad |= !isfinite(g[f]); }
        }
        if (bad && overflow_flag) *overflow_flag = 1;
        const uint32_t live_mask = __ballot_sync(0xffffffffu, live);
        if (live_mask == 0) continue;
        uint32_t gx = 0, gy = 0, gz = 0;
        float wx = 0.f, wy = 0.f, wz = 0.f;
        if (live) {
            float x, y, z;
            load_pos(e, i, x, y, z);
            const float px = fmaf(x, s, 0.5f), py = fmaf(y, s, 0.5f), pz = fmaf(z, s, 0.5f);
            const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
            wx = px - fx; wy = py - fy; wz = pz - fz;
            gx = (uint32_t)(int)fx; gy = (uint32_t)(int)fy; gz = (uint32_t)(int)fz;
        }
        float v[8][F];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float w = ((c & 1) ? wx : 1.f - wx) * (((c >> 1) & 1) ? wy : 1.f - wy) * ((c >> 2) ? wz : 1.f - wz);
#pragma unroll
            for (int f = 0; f < F; ++f) v[c][f] = live ? w * g[f] : 0.f;
        }
        // runs of consecutive live lanes in the same cell (21 bits per coordinate: resolutions up to 2^21)
        const unsigned long long key = live ? ((unsigned long long)gx | ((unsigned long long)gy << 21) | ((unsigned long long)gz << 42)) : ~0ull;
        const unsigned long long prev = __shfl_up_sync(0xffffffffu, key, 1);
        const bool head = live && (lane == 0 || prev != key);
        const uint32_t heads = __ballot_sync(0xffffffffu, head);
        const int my_run = __popc(heads & (0xffffffffu >> (31 - lane)));
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int other_run = __shfl_down_sync(0xffffffffu, my_run, d);
            const bool take = live && (lane + d < 32) && ((live_mask >> ((lane + d) & 31)) & 1u) && other_run == my_run;
            if (!__any_sync(0xffffffffu, take)) break;    // no run reaches this far: none reaches farther either
#pragma unroll
            for (int c = 0; c < 8; ++c)
#pragma unroll
                for (int f = 0; f < F; ++f) { const float o = __shfl_down_sync(0xffffffffu, v[c][f], d); if (take) v[c][f] += o; }
        }
        if (head) {
            uint32_t cx[2] = {gx, gx + 1u}, cy[2] = {gy, gy + 1u}, cz[2] = {gz, gz + 1u};
            if (m.mixed) {
                const float r = m.canon[l];
#pragma unroll
                for (int k = 0; k < 2; ++k) { cx[k] = canon_vertex(cx[k], r); cy[k] = canon_vertex(cy[k], r); cz[k] = canon_vertex(cz[k], r); }
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint32_t idx = grid_index(cx[c & 1], cy[(c >> 1) & 1], cz[c >> 2], res, size, hashed);
                atomic_add_vec<F>(lvl + (size_t)idx * F, v[c]);
            }
        }
    }
}

// Backward, second generation (F = 2; used by the fused field path).  Differences from grid_encode_bwd_kernel:
//   * TWO lanes per sample: lane parity picks the x-corner, so the two corners (x, x+1) of one (y, z) pair -- neighbours in
//     memory for dense levels and, 7 times out of 8, in the same 32-byte sector for hashed levels (the hash only XORs x into
//     the low bits) -- travel in the SAME red instruction and are coalesced into one L2 reduction packet: ~45% fewer L2
//     reduction sectors at the fine levels, which is what bounds this kernel;
//   * the incoming gradient is read level-major, dT[l][i] (half2), written that way by field_bwd_fused: a warp reads 64
//     contiguous bytes per level instead of 16 sectors;
//   * same run aggregation over consecutive samples in the same cell (shuffle distances 2, 4, 8, 16 lanes).
__global__ void __launch_bounds__(256)
grid_scatter_pair_kernel(const float4* __restrict__ x01, int n_max, const int32_t* __restrict__ n_dev, const uint32_t* __restrict__ dT, int64_t dT_stride,
                         const __grid_constant__ GridMeta m, float* __restrict__ dgrid, const WgradReduce wr) {
    constexpr uint32_t FULL = 0xffffffffu;
    if ((int)blockIdx.y == m.n_levels) {
        // extra grid row: sums the MLP weight-gradient partials of field_bwd_fused (one per CTA) into the parameter gradient while the
        // level rows scatter -- the two are independent, and a separate launch would sit on the step's critical path.
        // 32 outputs x 8 slices of the partial list per pass, combined through shared memory.
        __shared__ float sm[8][33];
        const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5, n_out = wr.n_sigma + wr.n_rgb;
        for (int j0 = blockIdx.x * 32; j0 < n_out; j0 += gridDim.x * 32) {
            const int j = j0 + lane;
            float s0 = 0.f, s1 = 0.f;
            if (j < n_out) {
                int c = slice;
                for (; c + 8 < wr.n_parts; c += 16) { s0 += wr.partials[(size_t)c * wr.stride + j]; s1 += wr.partials[(size_t)(c + 8) * wr.stride + j]; }
                if (c < wr.n_parts) s0 += wr.partials[(size_t)c * wr.stride + j];
            }
            sm[slice][lane] = s0 + s1;
            __syncthreads();
            if (slice == 0 && j < n_out) {
                float t = 0.f;
#pragma unroll
                for (int k = 0; k < 8; ++k) t += sm[k][lane];
                if (j < wr.n_sigma) wr.d_sigma[j] += t; else wr.d_rgb[j - wr.n_sigma] += t;
            }
            __syncthreads();
        }
        return;
    }
    const int n = n_dev ? min(*n_dev, n_max) : n_max;
    const int l = blockIdx.y;
    const int lane = threadIdx.x & 31, xb = lane & 1;
    const float s = m.scale[l];
    const uint32_t res = m.res[l], size = m.size[l];
    const bool hashed = (m.hashed >> l) & 1u;
    const uint32_t mask = size - 1u, r2 = res * res;
    float2* lvl = reinterpret_cast<float2*>(dgrid) + m.offset[l];
    const uint32_t* dl = dT + (size_t)l * dT_stride;
    const int n_pad = (n + 15) & ~15;   // whole warps stay in the loop (shuffles below)
    const uint64_t pol_keep = umma::policy_evict_last();   // gradient table lines stay in L2 while the other levels / kernels stream
    const int step = (int)gridDim.x * (int)(blockDim.x >> 1);
    // software pipeline: the gradient and the position of the NEXT sample of this lane pair are requested before the current one is
    // processed (the kernel used to spend ~40% of its issue slots waiting on these two loads at the top of every iteration)
    int i = (int)blockIdx.x * (int)(blockDim.x >> 1) + (int)(threadIdx.x >> 1);
    uint32_t raw_n = 0u;
    float4 p_n = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n) { raw_n = __ldg(dl + i); p_n = __ldg(x01 + i); }
    for (; i < n_pad; i += step) {
        const uint32_t raw = raw_n;
        const float4 p = p_n;
        raw_n = 0u;
        if (i + step < n) { raw_n = __ldg(dl + i + step); p_n = __ldg(x01 + i + step); }
        const bool live = (raw & 0x7fff7fffu) != 0u;          // either half non-zero
        const uint32_t live_mask = __ballot_sync(FULL, live);
        if (live_mask == 0) continue;
        uint32_t gx = 0, gy = 0, gz = 0;
        float v[4][2];   // corners (x = gx + xb, y = gy + (c & 1), z = gz + (c >> 1))
        if (live) {
            const float2 g = __half22float2(*reinterpret_cast<const __half2*>(&raw));
            const float px = fmaf(p.x, s, 0.5f), py = fmaf(p.y, s, 0.5f), pz = fmaf(p.z, s, 0.5f);
            const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
            const float wx = px - fx, wy = py - fy, wz = pz - fz;
            gx = (uint32_t)(int)fx; gy = (uint32_t)(int)fy; gz = (uint32_t)(int)fz;
            const float wxs = xb ? wx : 1.f - wx;
            const float a0 = wxs * (1.f - wy), a1 = wxs * wy;
            const float w0 = a0 * (1.f - wz), w1 = a1 * (1.f - wz), w2 = a0 * wz, w3 = a1 * wz;
            v[0][0] = w0 * g.x; v[0][1] = w0 * g.y; v[1][0] = w1 * g.x; v[1][1] = w1 * g.y;
            v[2][0] = w2 * g.x; v[2][1] = w2 * g.y; v[3][0] = w3 * g.x; v[3][1] = w3 * g.y;
        } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) { v[c][0] = 0.f; v[c][1] = 0.f; }
        }
        // runs of consecutive live samples in the same cell (the two lanes of a sample share the key)
        const unsigned long long key = live ? ((unsigned long long)gx | ((unsigned long long)gy << 21) | ((unsigned long long)gz << 42)) : ~0ull;
        const unsigned long long prev = __shfl_up_sync(FULL, key, 2);
        const bool head = live && (lane < 2 || prev != key);
        const uint32_t heads = __ballot_sync(FULL, head && xb == 0);
        if (heads != (live_mask & 0x55555555u)) {             // at least one run longer than one sample
            const int my_run = __popc(heads & (FULL >> (31 - lane)));
#pragma unroll
            for (int d = 2; d < 32; d <<= 1) {
                const int other_run = __shfl_down_sync(FULL, my_run, d);
                const bool take = live && (lane + d < 32) && ((live_mask >> ((lane + d) & 31)) & 1u) && other_run == my_run;
                if (!__any_sync(FULL, take)) break;    // no run reaches this far: none reaches farther either
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float o0 = __shfl_down_sync(FULL, v[c][0], d), o1 = __shfl_down_sync(FULL, v[c][1], d);
                    if (take) { v[c][0] += o0; v[c][1] += o1; }
                }
            }
        }
        if (head) {
            const uint32_t cx = gx + xb;
            uint32_t idx[4];
            if (m.mixed) {   // MixedFeature: hash of the canonical-grid vertices (always hashed, table of 2^T entries)
                const float r = m.canon[l];
                const uint32_t ccx = canon_vertex(cx, r);
                const uint32_t hy0 = canon_vertex(gy, r) * 2654435761u, hy1 = canon_vertex(gy + 1u, r) * 2654435761u;
                const uint32_t hz0 = canon_vertex(gz, r) * 805459861u, hz1 = canon_vertex(gz + 1u, r) * 805459861u;
                idx[0] = (ccx ^ hy0 ^ hz0) & mask; idx[1] = (ccx ^ hy1 ^ hz0) & mask; idx[2] = (ccx ^ hy0 ^ hz1) & mask; idx[3] = (ccx ^ hy1 ^ hz1) & mask;
            } else if (hashed) {
                const uint32_t hy0 = gy * 2654435761u, hy1 = hy0 + 2654435761u, hz0 = gz * 805459861u, hz1 = hz0 + 805459861u;
                idx[0] = (cx ^ hy0 ^ hz0) & mask; idx[1] = (cx ^ hy1 ^ hz0) & mask; idx[2] = (cx ^ hy0 ^ hz1) & mask; idx[3] = (cx ^ hy1 ^ hz1) & mask;
            } else {
                const uint32_t b = cx + gy * res + gz * r2;
                idx[0] = b; idx[1] = b + res; idx[2] = b + r2; idx[3] = b + res + r2;
                if (idx[3] >= size) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) idx[c] %= size;
                }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) umma::red_add_v2_hint(lvl + idx[c], v[c][0], v[c][1], pol_keep);
        }
    }
}

// degree-4 real spherical harmonics of v = 2*d01 - 1 (16 outputs)
__device__ __forceinline__ void sh4(float x, float y, float z, float (&o)[16]) {  // body shared via sh4.cuh
    sh4_eval(x, y, z, o);
}

__global__ void sh4_fwd_kernel(const float* __restrict__ d01, int64_t n, __half* __restrict__ out, int out_stride, int out_offset) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float o[16];
    sh4(fmaf(d01[3 * i], 2.f, -1.f), fmaf(d01[3 * i + 1], 2.f, -1.f), fmaf(d01[3 * i + 2], 2.f, -1.f), o);
    __align__(16) __half h[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) h[k] = __float2half_rn(o[k]);
    uint4* dst = reinterpret_cast<uint4*>(out + (size_t)i * out_stride + out_offset);
    dst[0] = reinterpret_cast<const uint4*>(h)[0];
    dst[1] = reinterpret_cast<const uint4*>(h)[1];
}

// Gradient w.r.t. the sample POSITION (the reference's --optimize_ext path: custom_functions.py:102-112 sums dL/dxyz over each ray's
// samples into dL/drays_o, dL/drays_d; tcnn: kernel_grid_backward_input).  Trilinear interpolation is piecewise linear in x:
//   d f_l / dx = scale_l * sum_{cy,cz} w_y(cy) w_z(cz) (v[x1,cy,cz] - v[x0,cy,cz])       (same for y, z)
// dx01 = sum over levels and features of dL_dout[l][f] * d f_l[f] / dx.  One thread per sample, fp32 accumulation.
template <int F>
__global__ void __launch_bounds__(128)
grid_encode_bwd_input_kernel(const __grid_constant__ EncArgs e, const __half* __restrict__ table, const __half* __restrict__ dL_dout,
                             const __grid_constant__ GridMeta m, float* __restrict__ dx01) {
    const int64_t n = e.n_max;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float x, y, z;
        load_pos(e, i, x, y, z);
        float gx = 0.f, gy = 0.f, gz = 0.f;
        for (int l = 0; l < m.n_levels; ++l) {
            float g[F];
            bool live = false;
#pragma unroll
            for (int f = 0; f < F; ++f) { g[f] = __half2float(dL_dout[(size_t)i * m.n_levels * F + l * F + f]); live |= (g[f] != 0.f); }
            if (!live) continue;
            const float s = m.scale[l];
            const uint32_t res = m.res[l], size = m.size[l];
            const bool hashed = (m.hashed >> l) & 1u;
            const float px = fmaf(x, s, 0.5f), py = fmaf(y, s, 0.5f), pz = fmaf(z, s, 0.5f);
            const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
            const float wx = px - fx, wy = py - fy, wz = pz - fz;
            const uint32_t cx0 = (uint32_t)(int)fx, cy0 = (uint32_t)(int)fy, cz0 = (uint32_t)(int)fz;
            uint32_t cx[2] = {cx0, cx0 + 1u}, cy[2] = {cy0, cy0 + 1u}, cz[2] = {cz0, cz0 + 1u};
            if (m.mixed) {
                const float r = m.canon[l];
#pragma unroll
                for (int k = 0; k < 2; ++k) { cx[k] = canon_vertex(cx[k], r); cy[k] = canon_vertex(cy[k], r); cz[k] = canon_vertex(cz[k], r); }
            }
            const typename FeatVec<F>::T* lvl = reinterpret_cast<const typename FeatVec<F>::T*>(table) + m.offset[l];
            float d[8];   // d[c] = <dL_dout_l, v_c>
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const typename FeatVec<F>::T raw = __ldg(lvl + grid_index(cx[c & 1], cy[(c >> 1) & 1], cz[c >> 2], res, size, hashed));
                const __half* h = reinterpret_cast<const __half*>(&raw);
                float acc = 0.f;
#pragma unroll
                for (int f = 0; f < F; ++f) acc = fmaf(g[f], __half2float(h[f]), acc);
                d[c] = acc;
            }
            const float ux = 1.f - wx, uy = 1.f - wy, uz = 1.f - wz;
            // corner c = x | y << 1 | z << 2
            gx = fmaf(s, uy * uz * (d[1] - d[0]) + wy * uz * (d[3] - d[2]) + uy * wz * (d[5] - d[4]) + wy * wz * (d[7] - d[6]), gx);
            gy = fmaf(s, ux * uz * (d[2] - d[0]) + wx * uz * (d[3] - d[1]) + ux * wz * (d[6] - d[4]) + wx * wz * (d[7] - d[5]), gy);
            gz = fmaf(s, ux * uy * (d[4] - d[0]) + wx * uy * (d[5] - d[1]) + ux * wy * (d[6] - d[2]) + wx * wy * (d[7] - d[3]), gz);
        }
        dx01[3 * i] = gx; dx01[3 * i + 1] = gy; dx01[3 * i + 2] = gz;
    }
}

// d SH4(2 d01 - 1) / d d01 contracted with dL_dout (n,16) fp16 -> (n,3) f32   (--optimize_ext: gradient w.r.t. the view direction)
__global__ void sh4_bwd_kernel(const float* __restrict__ d01, const __half* __restrict__ dL_dout, int64_t n, float* __restrict__ dd01) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = fmaf(d01[3 * i], 2.f, -1.f), y = fmaf(d01[3 * i + 1], 2.f, -1.f), z = fmaf(d01[3 * i + 2], 2.f, -1.f);
    float g[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) g[k] = __half2float(dL_dout[(size_t)i * 16 + k]);
    const float c1 = 0.48860251190291987f, c2 = 1.0925484305920792f, c3 = 0.94617469575755997f, c5 = 0.54627421529603959f, c6 = 0.59004358992664352f,
                c7 = 2.8906114426405538f, c8 = 0.45704579946446572f, c9 = 0.3731763325901154f, c10 = 1.4453057213202769f;
    const float x2 = x * x, y2 = y * y, z2 = z * z;
    const float gx = -c1 * g[3] + c2 * y * g[4] - c2 * z * g[7] + 2.f * c5 * x * g[8] - 6.f * c6 * x * y * g[9] + c7 * y * z * g[10] + c8 * (1.f - 5.f * z2) * g[13] +
                     2.f * c10 * x * z * g[14] + c6 * (-3.f * x2 + 3.f * y2) * g[15];
    const float gy = -c1 * g[1] + c2 * x * g[4] - c2 * z * g[5] - 2.f * c5 * y * g[8] + c6 * (-3.f * x2 + 3.f * y2) * g[9] + c7 * x * z * g[10] +
                     c8 * (1.f - 5.f * z2) * g[11] - 2.f * c10 * y * z * g[14] + 6.f * c6 * x * y * g[15];
    const float gz = c1 * g[2] - c2 * y * g[5] + 2.f * c3 * z * g[6] - c2 * x * g[7] + c7 * x * y * g[10] - 10.f * c8 * y * z * g[11] + c9 * (15.f * z2 - 3.f) * g[12] -
                     10.f * c8 * x * z * g[13] + c10 * (x2 - y2) * g[14];
    dd01[3 * i] = 2.f * gx; dd01[3 * i + 1] = 2.f * gy; dd01[3 * i + 2] = 2.f * gz;      // v = 2 d01 - 1
}

}  // namespace mfn

using namespace mfn;

#define MFN_F_DISPATCH(F_, BODY) \
    switch (F_) { case 1: { constexpr int F = 1; BODY } break; case 2: { constexpr int F = 2; BODY } break; \
                  case 4: { constexpr int F = 4; BODY } break; default: { constexpr int F = 8; BODY } break; }

extern "C" int64_t mfn_grid_layout(const mfn_grid_cfg* cfg, uint32_t* offsets_host, uint32_t* resolutions_host, float* scales_host) {
    GridMeta m;
    if (build_grid_meta(cfg, &m, "mfn_grid_layout") != MFN_OK) return -1;
    for (int l = 0; l < cfg->n_levels; ++l) {
        if (offsets_host) offsets_host[l] = m.offset[l];
        if (resolutions_host) resolutions_host[l] = m.res[l];
        if (scales_host) scales_host[l] = m.scale[l];
    }
    if (offsets_host) offsets_host[cfg->n_levels] = m.offset[cfg->n_levels];
    return (int64_t)m.offset[cfg->n_levels];
}



static inline unsigned enc_grid(int64_t n_max, int threads, int per_sm) {
    int64_t blocks = ceil_div(n_max, threads);
    const int64_t cap = (int64_t)kNumSMs * per_sm;
    return (unsigned)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

namespace mfn {
int grid_encode_forward(const EncArgs& e, const __half* table, const GridMeta& m, int F_, __half* out, cudaStream_t st) {
    if (e.n_max <= 0) return MFN_OK;
    ProfScope ps("grid_encode_fwd", st);
    MFN_F_DISPATCH(F_, (grid_encode_fwd_kernel<F><<<enc_grid(e.n_max, 128, 16), 128, 0, st>>>(e, table, m, out));)
    return check_launch("mfn_grid_encode_fwd", st);
}
int grid_scatter_level_major(const float4* x01, int64_t n_max, const int32_t* n_dev, const __half* dT, int64_t dT_stride, const GridMeta& m, float* dgrid,
                             const WgradReduce& wr, cudaStream_t st) {
    if (n_max <= 0) return MFN_OK;
    if (n_max > 0x7fffffff) { set_error("mfn_field_bwd: more than 2^31 samples"); return MFN_ERR_ARG; }
    static int per_sm = 0;
    if (per_sm == 0) { const char* e = getenv("MFN_SCATTER_BPS"); per_sm = e ? atoi(e) : 2; if (per_sm < 1) per_sm = 2; }       // measured: 2 -> 174 us, 8 -> 178 us, 64 -> 235 us (per level)
    dim3 grid(enc_grid(n_max, 128, per_sm), (unsigned)m.n_levels + (wr.partials ? 1u : 0u));
    ProfScope ps("grid_encode_bwd", st);
    grid_scatter_pair_kernel<<<grid, 256, 0, st>>>(x01, (int)n_max, n_dev, reinterpret_cast<const uint32_t*>(dT), dT_stride, m, dgrid, wr);
    return check_launch("mfn_grid_encode_bwd(level-major)", st);
}
int grid_encode_backward(const EncArgs& e, const __half* dL_dout, const GridMeta& m, int F_, float* dgrid, int32_t* overflow_flag, cudaStream_t st) {
    if (e.n_max <= 0) return MFN_OK;
    dim3 grid(enc_grid(e.n_max, 256, 8), (unsigned)m.n_levels);
    ProfScope ps("grid_encode_bwd", st);
    MFN_F_DISPATCH(F_, (grid_encode_bwd_kernel<F><<<grid, 256, 0, st>>>(e, dL_dout, m, dgrid, overflow_flag));)
    return check_launch("mfn_grid_encode_bwd", st);
}
}  // namespace mfn

extern "C" int mfn_grid_encode_fwd(const float* x01, const void* table, const mfn_grid_cfg* cfg, int64_t n, void* out, void* stream) {
    GridMeta m;
    int rc = build_grid_meta(cfg, &m, "mfn_grid_encode_fwd");
    if (rc != MFN_OK) return rc;
    if (n < 0) { set_error("mfn_grid_encode_fwd: bad n"); return MFN_ERR_ARG; }
    if (n > 0 && (!x01 || !table || !out)) { set_error("mfn_grid_encode_fwd: null pointer"); return MFN_ERR_ARG; }
    EncArgs e{}; e.x = x01; e.normalize = false; e.n_max = n; e.n_dev = nullptr;
    return grid_encode_forward(e, (const __half*)table, m, cfg->n_features, (__half*)out, (cudaStream_t)stream);
}

extern "C" int mfn_grid_encode_bwd(const float* x01, const void* dL_dout, const mfn_grid_cfg* cfg, int64_t n, float* dgrid, void* stream) {
    GridMeta m;
    int rc = build_grid_meta(cfg, &m, "mfn_grid_encode_bwd");
    if (rc != MFN_OK) return rc;
    if (n < 0) { set_error("mfn_grid_encode_bwd: bad n"); return MFN_ERR_ARG; }
    if (n > 0 && (!x01 || !dL_dout || !dgrid)) { set_error("mfn_grid_encode_bwd: null pointer"); return MFN_ERR_ARG; }
    EncArgs e{}; e.x = x01; e.normalize = false; e.n_max = n; e.n_dev = nullptr;
    return grid_encode_backward(e, (const __half*)dL_dout, m, cfg->n_features, dgrid, nullptr, (cudaStream_t)stream);
}

extern "C" int mfn_sh4_fwd(const float* dirs01, int64_t n, void* out, int out_stride, int out_offset, void* stream) {
    if (n < 0 || out_stride < 16 || out_offset < 0 || (out_stride % 8) || (out_offset % 8)) { set_error("mfn_sh4_fwd: bad argument"); return MFN_ERR_ARG; }
    if (n == 0) return MFN_OK;
    if (!dirs01 || !out) { set_error("mfn_sh4_fwd: null pointer"); return MFN_ERR_ARG; }
    sh4_fwd_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(dirs01, n, (__half*)out, out_stride, out_offset);
    return check_launch("mfn_sh4_fwd", (cudaStream_t)stream);
}

extern "C" int mfn_grid_encode_bwd_input(const float* x01, const void* table, const void* dL_dout, const mfn_grid_cfg* cfg, int64_t n, float* dx01, void* stream) {
    GridMeta m;
    int rc = build_grid_meta(cfg, &m, "mfn_grid_encode_bwd_input");
    if (rc != MFN_OK) return rc;
    if (n < 0) { set_error("mfn_grid_encode_bwd_input: bad n"); return MFN_ERR_ARG; }
    if (n == 0) return MFN_OK;
    if (!x01 || !table || !dL_dout || !dx01) { set_error("mfn_grid_encode_bwd_input: null pointer"); return MFN_ERR_ARG; }
    EncArgs e{}; e.x = x01; e.normalize = false; e.n_max = n; e.n_dev = nullptr;
    MFN_F_DISPATCH(cfg->n_features, (grid_encode_bwd_input_kernel<F><<<enc_grid(n, 128, 16), 128, 0, (cudaStream_t)stream>>>(e, (const __half*)table, (const __half*)dL_dout, m, dx01));)
    return check_launch("mfn_grid_encode_bwd_input", (cudaStream_t)stream);
}

extern "C" int mfn_sh4_bwd(const float* dirs01, const void* dL_dout, int64_t n, float* ddirs01, void* stream) {
    if (n < 0) { set_error("mfn_sh4_bwd: bad argument"); return MFN_ERR_ARG; }
    if (n == 0) return MFN_OK;
    if (!dirs01 || !dL_dout || !ddirs01) { set_error("mfn_sh4_bwd: null pointer"); return MFN_ERR_ARG; }
    sh4_bwd_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(dirs01, (const __half*)dL_dout, n, ddirs01);
    return check_launch("mfn_sh4_bwd", (cudaStream_t)stream);
}
