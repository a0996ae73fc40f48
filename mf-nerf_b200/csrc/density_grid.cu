// NGP.update_density_grid (ref: models/networks.py:157-197, 242-271) as a handful of fused kernels without host round trips:
//   cells -> jittered world positions          (get_all_cells / sample_uniform_and_occupied_cells + the xyzs_w arithmetic, l.157-197, 254-260)
//   [density query = mfn_density_fwd]
//   decay / max update of the grid              (l.263-266; cells < 0, i.e. invisible ones, are never touched)
//   mean over the positive cells                (l.268)  -> feeds mfn_packbits_dev_thr (l.270-271)
// The occupied-cell list the reference builds with nonzero() is replaced, in engine.py, by a cumsum + searchsorted draw (no sync).
// Random numbers are counter-based (splitmix64 of seed and element index); the reference draws them with torch.rand / randint, so
// the streams differ -- the distributions are the reference's (uniform cell, uniform jitter inside the cell).
#include "common.cuh"
#include "../../include/mfnerf_b200.h"

namespace mfn {

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__device__ __forceinline__ float u01(uint32_t bits) { return (float)(bits >> 8) * (1.0f / 16777216.0f); }   // [0, 1)

// idx == nullptr && random == 0 : cell i (all cells, morton order)     (get_all_cells)
// random != 0                   : uniformly random cell, its morton index is written to idx_out  (first half of the sampled cells)
// idx != nullptr                : the given morton indices                                        (the occupied half)
__global__ void grid_positions_kernel(const int32_t* __restrict__ idx, int random, int64_t n, float s, int grid_size, uint64_t seed,
                                      int32_t* __restrict__ idx_out, float* __restrict__ xyz) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const float half = s / (float)grid_size;                       // half_grid_size, networks.py:255
    const float span = s - half;
    const float inv = 1.0f / (float)(grid_size - 1);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t r0 = splitmix64(seed ^ (uint64_t)i * 0xD1342543DE82EF95ull), r1 = splitmix64(r0);
        uint32_t cx, cy, cz;
        if (random) {
            const uint64_t r2 = splitmix64(r1);
            cx = (uint32_t)((r2 & 0xfffff) * (uint64_t)grid_size >> 20);
            cy = (uint32_t)(((r2 >> 20) & 0xfffff) * (uint64_t)grid_size >> 20);
            cz = (uint32_t)(((r2 >> 40) & 0xfffff) * (uint64_t)grid_size >> 20);
            idx_out[i] = (int32_t)morton_encode(cx, cy, cz);
        } else {
            const uint32_t mi = idx ? (uint32_t)idx[i] : (uint32_t)i;
            cx = compact3(mi); cy = compact3(mi >> 1); cz = compact3(mi >> 2);
        }
        // xyzs_w = (coords / (G-1) * 2 - 1) * (s - half) + (rand * 2 - 1) * half      (networks.py:256-260)
        const float jx = u01((uint32_t)r0) * 2.f - 1.f, jy = u01((uint32_t)(r0 >> 32)) * 2.f - 1.f, jz = u01((uint32_t)r1) * 2.f - 1.f;
        xyz[3 * i] = ((float)cx * inv * 2.f - 1.f) * span + jx * half;
        xyz[3 * i + 1] = ((float)cy * inv * 2.f - 1.f) * span + jy * half;
        xyz[3 * i + 2] = ((float)cz * inv * 2.f - 1.f) * span + jz * half;
    }
}

// occupied-cell draw (networks.py:186-193: indices2 = nonzero(grid > thr); indices2[randint(len, (M,))]) without materialising the
// list: `cs` is the inclusive prefix count of occupied cells, draw k uniformly in [0, total) and binary-search the k-th occupied cell
__global__ void grid_draw_occupied_kernel(const int32_t* __restrict__ cs, int64_t n_cells, int64_t n, uint64_t seed, int32_t* __restrict__ idx_out) {
    const int32_t total = cs[n_cells - 1];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t r = splitmix64(seed ^ (uint64_t)i * 0xA24BAED4963EE407ull);
        if (total <= 0) {      // no occupied cell yet (the reference then adds no second half, l.188): spend the draws on uniform cells
            idx_out[i] = (int32_t)uniform_below(r, (uint32_t)n_cells);
            continue;
        }
        const int32_t k = (int32_t)uniform_below(r, (uint32_t)total);     // uniform in [0, total)
        int64_t lo = 0, hi = n_cells - 1;                  // smallest j with cs[j] > k
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (__ldg(cs + mid) > k) hi = mid; else lo = mid + 1;
        }
        idx_out[i] = (int32_t)lo;
    }
}

// all cells were queried: grid = grid < 0 ? grid : max(grid * decay, sigma)
__global__ void grid_update_all_kernel(float* __restrict__ grid, const float* __restrict__ sig, int64_t n, float decay) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float g = grid[i];
        if (!(g < 0.f)) grid[i] = fmaxf(g * decay, sig[i]);
    }
}
// sampled cells: first the decay over the whole cascade, then a scatter-max of the queried densities (non-negative floats order like
// their bit patterns, so atomicMax on the int view is exact; duplicates resolve to their maximum, where torch's index_put keeps an
// arbitrary one)
__global__ void grid_decay_kernel(float* __restrict__ grid, int64_t n, float decay) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float g = grid[i];
        if (!(g < 0.f)) grid[i] = g * decay;
    }
}
__global__ void grid_scatter_max_kernel(float* __restrict__ grid, const int32_t* __restrict__ idx, const float* __restrict__ sig, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float s = sig[i];
        float* p = grid + idx[i];
        if (s > 0.f && !(*p < 0.f)) atomicMax(reinterpret_cast<int*>(p), __float_as_int(s));
    }
}

// acc[0] += sum of positive cells, acc[1] += their count (fp32 block partials, one atomic pair per block); the last block writes
// mean = acc[0] / max(acc[1], 1) and clears acc and the ticket for the next call
__global__ void __launch_bounds__(256)
grid_mean_positive_kernel(const float* __restrict__ grid, int64_t n, float* __restrict__ acc, unsigned int* __restrict__ ticket, float* __restrict__ mean_out) {
    __shared__ float ssum[8], scnt[8];
    float s = 0.f, c = 0.f;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float g = grid[i];
        if (g > 0.f) { s += g; c += 1.f; }
    }
    s = warp_sum(s); c = warp_sum(c);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { ssum[w] = s; scnt[w] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float S = 0.f, C = 0.f;
        for (int k = 0; k < 8; ++k) { S += ssum[k]; C += scnt[k]; }
        atomicAdd(acc, S); atomicAdd(acc + 1, C);
        __threadfence();
        if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
            __threadfence();
            const float St = atomicAdd(acc, 0.f), Ct = atomicAdd(acc + 1, 0.f);
            *mean_out = St / fmaxf(Ct, 1.f);
            acc[0] = 0.f; acc[1] = 0.f; *ticket = 0u;
        }
    }
}

// mark_invisible_cells (networks.py:199-240, run once before training, train.py:159-162): a cell is valid when at least one camera
// sees its centre at depth >= near inside the image and no camera sees it closer than near; valid -> density 0, else -1 (cells at -1
// are never updated nor marched).  One thread per cell (morton order, like get_all_cells l.157-170), cameras staged through shared
// memory as world-to-camera rows [R^T | -R^T t] (l.215-216).
constexpr int kCamChunk = 128;
__global__ void __launch_bounds__(256)
grid_mark_invisible_kernel(const float* __restrict__ K, const float* __restrict__ poses, int n_cams, float img_w, float img_h, float span, int grid_size,
                           float near_d, int64_t n_cells, float* __restrict__ density, float* __restrict__ count_grid) {
    __shared__ float cam[kCamChunk][12];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < n_cells;
    float x = 0.f, y = 0.f, z = 0.f;
    if (live) {
        const uint32_t mi = (uint32_t)i;
        // xyzs = coords / (G-1) * 2 - 1 ;  xyzs_w = xyzs * (s - half_grid_size)      (l.221-224).  torch's CUDA division of a tensor by a
        // python scalar multiplies by the fp32 reciprocal, and the reference runs this on the GPU: same here, to stay bit-compatible
        const float inv = __fdiv_rn(1.f, (float)(grid_size - 1));
        x = __fmul_rn(__fsub_rn(__fmul_rn(__fmul_rn((float)compact3(mi), inv), 2.f), 1.f), span);
        y = __fmul_rn(__fsub_rn(__fmul_rn(__fmul_rn((float)compact3(mi >> 1), inv), 2.f), 1.f), span);
        z = __fmul_rn(__fsub_rn(__fmul_rn(__fmul_rn((float)compact3(mi >> 2), inv), 2.f), 1.f), span);
    }
    float k[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) k[j] = __ldg(K + j);
    int covered = 0;
    bool too_near = false;
    for (int c0 = 0; c0 < n_cams; c0 += kCamChunk) {
        const int nc = min(kCamChunk, n_cams - c0);
        __syncthreads();
        for (int t = threadIdx.x; t < nc; t += blockDim.x) {
            const float* P = poses + 12 * (size_t)(c0 + t);          // c2w rows [R_a0 R_a1 R_a2 t_a]
            float R[3][3], T[3];
            for (int a = 0; a < 3; ++a) { for (int b = 0; b < 3; ++b) R[a][b] = P[4 * a + b]; T[a] = P[4 * a + 3]; }
            for (int a = 0; a < 3; ++a) {                            // w2c_R = R^T ; w2c_T = -(R^T t)
                cam[t][4 * a] = R[0][a]; cam[t][4 * a + 1] = R[1][a]; cam[t][4 * a + 2] = R[2][a];
                cam[t][4 * a + 3] = -fmaf(R[2][a], T[2], fmaf(R[1][a], T[1], R[0][a] * T[0]));
            }
        }
        __syncthreads();
        if (!live) continue;
        for (int t = 0; t < nc; ++t) {
            const float* w = cam[t];
            const float xc = fmaf(w[2], z, fmaf(w[1], y, w[0] * x)) + w[3];        // xyzs_c = w2c_R @ xyzs_w + w2c_T   (l.225)
            const float yc = fmaf(w[6], z, fmaf(w[5], y, w[4] * x)) + w[7];
            const float zc = fmaf(w[10], z, fmaf(w[9], y, w[8] * x)) + w[11];
            const float ud = fmaf(k[2], zc, fmaf(k[1], yc, k[0] * xc));            // uvd = K @ xyzs_c                  (l.226)
            const float vd = fmaf(k[5], zc, fmaf(k[4], yc, k[3] * xc));
            const float d = fmaf(k[8], zc, fmaf(k[7], yc, k[6] * xc));
            const float u = __fdiv_rn(ud, d), v = __fdiv_rn(vd, d);                // uv = uvd[:2] / uvd[2]             (l.227)
            const bool in_image = (d >= 0.f) && (u >= 0.f) && (u < img_w) && (v >= 0.f) && (v < img_h);
            covered += (in_image && d >= near_d) ? 1 : 0;                          // l.231
            too_near |= in_image && (d < near_d);                                  // l.236-238
        }
    }
    if (!live) return;
    const float count = __fmul_rn((float)covered, __fdiv_rn(1.f, (float)n_cams)); // l.233-234 (sum / N_cams, scalar divisor: reciprocal multiply as above)
    if (count_grid) count_grid[i] = count;
    density[i] = (count > 0.f && !too_near) ? 0.f : -1.f;                          // l.240-242
}

static inline int dg_grid(int64_t n) {
    int64_t b = ceil_div(n, 256);
    const int64_t cap = (int64_t)kNumSMs * 8;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace mfn

using namespace mfn;

extern "C" int mfn_grid_cell_positions(const int32_t* cell_indices, int random_cells, int64_t n, int cascade, float scale, int grid_size, uint64_t seed,
                                       int32_t* indices_out, float* xyzs, void* stream) {
    if (n < 0 || cascade < 0 || grid_size < 2 || grid_size > 1024) { set_error("mfn_grid_cell_positions: bad argument"); return MFN_ERR_ARG; }
    if (n == 0) return MFN_OK;
    if (!xyzs || (random_cells && !indices_out)) { set_error("mfn_grid_cell_positions: null pointer"); return MFN_ERR_ARG; }
    const float s = fminf(exp2f((float)(cascade - 1)), scale);      // networks.py:253
    grid_positions_kernel<<<dg_grid(n), 256, 0, (cudaStream_t)stream>>>(cell_indices, random_cells, n, s, grid_size, seed, indices_out, xyzs);
    return check_launch("mfn_grid_cell_positions", (cudaStream_t)stream);
}

extern "C" int mfn_grid_draw_occupied(const int32_t* occupied_prefix_count, int64_t n_cells, int64_t n, uint64_t seed, int32_t* indices_out, void* stream) {
    if (n_cells < 1 || n < 0) { set_error("mfn_grid_draw_occupied: bad argument"); return MFN_ERR_ARG; }
    if (n == 0) return MFN_OK;
    if (!occupied_prefix_count || !indices_out) { set_error("mfn_grid_draw_occupied: null pointer"); return MFN_ERR_ARG; }
    grid_draw_occupied_kernel<<<dg_grid(n), 256, 0, (cudaStream_t)stream>>>(occupied_prefix_count, n_cells, n, seed, indices_out);
    return check_launch("mfn_grid_draw_occupied", (cudaStream_t)stream);
}

extern "C" int mfn_grid_update(float* density_grid_cascade, const int32_t* cell_indices, const float* sigmas, int64_t n_cells, int64_t n, float decay,
                               void* stream) {
    if (n_cells < 0 || n < 0) { set_error("mfn_grid_update: bad argument"); return MFN_ERR_ARG; }
    if (n_cells == 0) return MFN_OK;
    if (!density_grid_cascade || (n > 0 && !sigmas)) { set_error("mfn_grid_update: null pointer"); return MFN_ERR_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    if (!cell_indices) {
        if (n != n_cells) { set_error("mfn_grid_update: without indices every cell needs a density"); return MFN_ERR_ARG; }
        grid_update_all_kernel<<<dg_grid(n_cells), 256, 0, st>>>(density_grid_cascade, sigmas, n_cells, decay);
    } else {
        grid_decay_kernel<<<dg_grid(n_cells), 256, 0, st>>>(density_grid_cascade, n_cells, decay);
        note_launch(1);
        if (n > 0) grid_scatter_max_kernel<<<dg_grid(n), 256, 0, st>>>(density_grid_cascade, cell_indices, sigmas, n);
    }
    return check_launch("mfn_grid_update", st);
}

extern "C" int mfn_grid_mean_positive(const float* density_grid, int64_t n, float* scratch16, float* mean_out, void* stream) {
    if (n < 0) { set_error("mfn_grid_mean_positive: bad argument"); return MFN_ERR_ARG; }
    if (!density_grid || !scratch16 || !mean_out) { set_error("mfn_grid_mean_positive: null pointer"); return MFN_ERR_ARG; }
    grid_mean_positive_kernel<<<dg_grid(n > 0 ? n : 1), 256, 0, (cudaStream_t)stream>>>(density_grid, n, scratch16, reinterpret_cast<unsigned int*>(scratch16 + 2), mean_out);
    return check_launch("mfn_grid_mean_positive", (cudaStream_t)stream);
}

extern "C" int mfn_grid_mark_invisible(const float* K, const float* poses, int32_t n_cams, int32_t img_w, int32_t img_h, int32_t cascades, float scale,
                                       int32_t grid_size, float near_distance, float* density_grid, float* count_grid, void* stream) {
    if (n_cams < 1 || img_w < 1 || img_h < 1 || cascades < 1 || grid_size < 2 || grid_size > 1024) { set_error("mfn_grid_mark_invisible: bad argument"); return MFN_ERR_ARG; }
    if (!K || !poses || !density_grid) { set_error("mfn_grid_mark_invisible: null pointer"); return MFN_ERR_ARG; }
    const int64_t n_cells = (int64_t)grid_size * grid_size * grid_size;
    for (int c = 0; c < cascades; ++c) {
        const double s = fmin(exp2((double)(c - 1)), (double)scale);       // python floats: s = min(2**(c-1), scale); half = s / G  (l.222-223)
        const float span = (float)(s - s / (double)grid_size);
        grid_mark_invisible_kernel<<<(unsigned)ceil_div(n_cells, 256), 256, 0, (cudaStream_t)stream>>>(
            K, poses, n_cams, (float)img_w, (float)img_h, span, grid_size, near_distance, n_cells, density_grid + (size_t)c * n_cells,
            count_grid ? count_grid + (size_t)c * n_cells : nullptr);
        if (c) note_launch(1);
    }
    return check_launch("mfn_grid_mark_invisible", (cudaStream_t)stream);
}
