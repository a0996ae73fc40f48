// Microbenchmark of hash-grid backward (gradient scatter) variants on ray-coherent sample positions.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo tools/encbwd_bench.cu -o tools/bin/encbwd_bench \
//        -L mf-nerf_b200/lib -lmfnerf_b200 -Xlinker -rpath -Xlinker $PWD/mf-nerf_b200/lib
// Run:   tools/bin/encbwd_bench tools/data/samples_lego.bin [alive_fraction]
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>
#include "../include/mfnerf_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int L = 16;
struct Meta { uint32_t off[L + 1]; uint32_t res[L]; float scale[L]; uint32_t hashed; };

__device__ __forceinline__ uint32_t gidx(uint32_t x, uint32_t y, uint32_t z, uint32_t res, uint32_t size, bool hashed) {
    uint32_t i = hashed ? (x ^ (y * 2654435761u) ^ (z * 805459861u)) : (x + y * res + z * res * res);
    return i % size;
}

struct Cell { uint32_t gx, gy, gz; float wx, wy, wz; };
__device__ __forceinline__ Cell locate(const float* __restrict__ x01, int64_t i, float s) {
    Cell c;
    const float px = fmaf(x01[3 * i], s, 0.5f), py = fmaf(x01[3 * i + 1], s, 0.5f), pz = fmaf(x01[3 * i + 2], s, 0.5f);
    const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
    c.wx = px - fx; c.wy = py - fy; c.wz = pz - fz;
    c.gx = (uint32_t)(int)fx; c.gy = (uint32_t)(int)fy; c.gz = (uint32_t)(int)fz;
    return c;
}
__device__ __forceinline__ float cw(const Cell& c, int k) {
    return ((k & 1) ? c.wx : 1.f - c.wx) * (((k >> 1) & 1) ? c.wy : 1.f - c.wy) * ((k >> 2) ? c.wz : 1.f - c.wz);
}

// V1: level-major, dL_dout transposed to [L][N] half2 (coalesced), float2 atomics
__global__ void __launch_bounds__(256) v1_kernel(const float* __restrict__ x01, const __half2* __restrict__ dT, const __grid_constant__ Meta m, int64_t n,
                                                  float* __restrict__ dgrid) {
    const int l = blockIdx.y;
    const float s = m.scale[l];
    const uint32_t res = m.res[l], size = m.off[l + 1] - m.off[l];
    const bool hashed = (m.hashed >> l) & 1u;
    float2* lvl = reinterpret_cast<float2*>(dgrid) + m.off[l];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float2 g = __half22float2(dT[(size_t)l * n + i]);
        if (g.x == 0.f && g.y == 0.f) continue;
        const Cell c = locate(x01, i, s);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float w = cw(c, k);
            atomicAdd(lvl + gidx(c.gx + (k & 1), c.gy + ((k >> 1) & 1), c.gz + (k >> 2), res, size, hashed), make_float2(w * g.x, w * g.y));
        }
    }
}

// V2: as V1 with half2 atomics into an fp16 gradient table
__global__ void __launch_bounds__(256) v2_kernel(const float* __restrict__ x01, const __half2* __restrict__ dT, const __grid_constant__ Meta m, int64_t n,
                                                  __half2* __restrict__ dgrid) {
    const int l = blockIdx.y;
    const float s = m.scale[l];
    const uint32_t res = m.res[l], size = m.off[l + 1] - m.off[l];
    const bool hashed = (m.hashed >> l) & 1u;
    __half2* lvl = dgrid + m.off[l];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float2 g = __half22float2(dT[(size_t)l * n + i]);
        if (g.x == 0.f && g.y == 0.f) continue;
        const Cell c = locate(x01, i, s);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float w = cw(c, k);
            atomicAdd(lvl + gidx(c.gx + (k & 1), c.gy + ((k >> 1) & 1), c.gz + (k >> 2), res, size, hashed), __floats2half2_rn(w * g.x, w * g.y));
        }
    }
}

// V3: V1 + aggregation of runs of consecutive lanes that sit in the same cell (samples along one ray), adaptive per warp
template <bool HALF>
__global__ void __launch_bounds__(256) v3_kernel(const float* __restrict__ x01, const __half2* __restrict__ dT, const __grid_constant__ Meta m, int64_t n,
                                                  void* __restrict__ dgrid_, int min_heads_skip) {
    const int l = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const float s = m.scale[l];
    const uint32_t res = m.res[l], size = m.off[l + 1] - m.off[l];
    const bool hashed = (m.hashed >> l) & 1u;
    const int64_t n_pad = (n + 31) / 32 * 32;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += (int64_t)gridDim.x * blockDim.x) {
        float2 g = make_float2(0.f, 0.f);
        Cell c{};
        uint32_t key = 0xffffffffu;
        if (i < n) {
            g = __half22float2(dT[(size_t)l * n + i]);
            if (g.x != 0.f || g.y != 0.f) { c = locate(x01, i, s); key = c.gx | (c.gy << 11) | (c.gz << 22); }
        }
        const bool live = key != 0xffffffffu;
        const uint32_t live_mask = __ballot_sync(0xffffffffu, live);
        if (live_mask == 0) continue;
        const uint32_t prev = __shfl_up_sync(0xffffffffu, key, 1);
        const bool head = live && (lane == 0 || prev != key);
        const uint32_t heads = __ballot_sync(0xffffffffu, head);
        float v[16];
#pragma unroll
        for (int k = 0; k < 8; ++k) { const float w = live ? cw(c, k) : 0.f; v[2 * k] = w * g.x; v[2 * k + 1] = w * g.y; }
        const bool aggregate = __popc(heads) * min_heads_skip <= __popc(live_mask);   // warp-uniform
        if (aggregate) {
            // segmented reduction towards the run head: lane i adds lane i+d if that lane belongs to the same run
            const uint32_t run_id_mask = heads;   // run index of a lane = popc(heads & lanes <= lane)
            const int my_run = __popc(run_id_mask & (0xffffffffu >> (31 - lane)));
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int other_run = __shfl_down_sync(0xffffffffu, my_run, d);
                const bool other_live = (live_mask >> ((lane + d) & 31)) & 1u;
                const bool take = (lane + d < 32) && other_live && live && other_run == my_run;
#pragma unroll
                for (int k = 0; k < 16; ++k) { const float o = __shfl_down_sync(0xffffffffu, v[k], d); if (take) v[k] += o; }
            }
        }
        if (aggregate ? head : live) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t idx = gidx(c.gx + (k & 1), c.gy + ((k >> 1) & 1), c.gz + (k >> 2), res, size, hashed);
                if (HALF) atomicAdd(reinterpret_cast<__half2*>(dgrid_) + m.off[l] + idx, __floats2half2_rn(v[2 * k], v[2 * k + 1]));
                else atomicAdd(reinterpret_cast<float2*>(dgrid_) + m.off[l] + idx, make_float2(v[2 * k], v[2 * k + 1]));
            }
        }
    }
}

// V5: raw atomic-rate probes: `per_thread` float2 (or half2 / float) atomics per thread at random / coherent addresses in a `span`-entry table
template <int MODE>
__global__ void __launch_bounds__(256) v5_kernel(void* __restrict__ tab, uint32_t span, int64_t n_threads, int per_thread) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_threads; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t h = (uint32_t)i * 2654435761u + 12345u;
        for (int k = 0; k < per_thread; ++k) {
            h = h * 1664525u + 1013904223u;
            const uint32_t idx = (h >> 8) % span;
            if (MODE == 0) atomicAdd(reinterpret_cast<float2*>(tab) + idx, make_float2(1e-6f, 1e-6f));
            else if (MODE == 1) atomicAdd(reinterpret_cast<__half2*>(tab) + idx, __floats2half2_rn(1e-6f, 1e-6f));
            else atomicAdd(reinterpret_cast<float*>(tab) + idx, 1e-6f);
        }
    }
}

static float* d_flush = nullptr;
static void flush_l2() { CK(cudaMemsetAsync(d_flush, 1, 256u << 20, 0)); }

template <typename F>
static float time_it(F launch, void* grads, size_t grad_bytes, int reps = 7) {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    std::vector<float> t;
    for (int r = 0; r < reps; ++r) {
        CK(cudaMemsetAsync(grads, 0, grad_bytes, 0));
        flush_l2();
        CK(cudaMemsetAsync(grads, 0, grad_bytes, 0));   // leave the (zeroed) gradient table L2-resident like the optimiser does
        CK(cudaEventRecord(a, 0)); launch(); CK(cudaEventRecord(b, 0)); CK(cudaEventSynchronize(b));
        CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, a, b)); t.push_back(ms * 1e3f);
    }
    std::sort(t.begin(), t.end());
    return t[t.size() / 2];
}

int main(int argc, char** argv) {
    const char* path = argc > 1 ? argv[1] : "tools/data/samples_lego.bin";
    const float alive_frac = argc > 2 ? (float)atof(argv[2]) : 0.6f;
    FILE* f = fopen(path, "rb"); if (!f) { printf("cannot open %s\n", path); return 1; }
    int64_t n; if (fread(&n, 8, 1, f) != 1) return 1;
    std::vector<float> raw(n * 4); if (fread(raw.data(), 4, n * 4, f) != (size_t)n * 4) return 1; fclose(f);
    std::vector<float> x(n * 3); std::vector<__half> dy(n * 32), dyT(n * 32);
    uint32_t rng = 1;
    int64_t alive = 0;
    for (int64_t i = 0; i < n; ++i) {
        for (int k = 0; k < 3; ++k) x[3 * i + k] = raw[4 * i + k];
        const bool live = raw[4 * i + 3] < alive_frac;
        alive += live;
        for (int k = 0; k < 32; ++k) {
            rng = rng * 1664525u + 1013904223u;
            const float v = live ? ((float)(rng >> 8) / 8388608.f - 1.f) * 1e-2f : 0.f;
            dy[i * 32 + k] = __float2half(v);
            dyT[(size_t)(k / 2) * n * 2 + i * 2 + (k & 1)] = __float2half(v);
        }
    }
    mfn_grid_cfg cfg{16, 2, 19, 16, exp(log(2048 * 0.5 / 16) / 15), MFN_GRID_HASH, 1};
    Meta m{}; uint32_t off[L + 1], res[L]; float sc[L];
    const int64_t entries = mfn_grid_layout(&cfg, off, res, sc);
    for (int l = 0; l < L; ++l) { m.off[l] = off[l]; m.res[l] = res[l]; m.scale[l] = sc[l]; if ((uint64_t)res[l] * res[l] * res[l] > off[l + 1] - off[l]) m.hashed |= 1u << l; }
    m.off[L] = off[L];
    printf("samples %lld (alive %.1f%%), entries %lld, hashed mask %x\n", (long long)n, 100.0 * alive / n, (long long)entries, m.hashed);
    float *d_x, *d_g32, *d_ref; __half *d_dy, *d_dyT, *d_g16;
    CK(cudaMalloc(&d_x, n * 12)); CK(cudaMalloc(&d_dy, n * 64)); CK(cudaMalloc(&d_dyT, n * 64));
    CK(cudaMalloc(&d_g32, entries * 8)); CK(cudaMalloc(&d_ref, entries * 8)); CK(cudaMalloc(&d_g16, entries * 4)); CK(cudaMalloc(&d_flush, 256u << 20));
    CK(cudaMemcpy(d_x, x.data(), n * 12, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_dy, dy.data(), n * 64, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_dyT, dyT.data(), n * 64, cudaMemcpyHostToDevice));
    const dim3 grid((unsigned)std::min<int64_t>((n + 255) / 256, 148 * 8), L);

    auto check = [&](const char* name, bool half) {
        std::vector<float> a(entries * 2), b(entries * 2);
        CK(cudaMemcpy(a.data(), d_ref, entries * 8, cudaMemcpyDeviceToHost));
        if (half) {
            std::vector<__half> h(entries * 2); CK(cudaMemcpy(h.data(), d_g16, entries * 4, cudaMemcpyDeviceToHost));
            for (int64_t i = 0; i < entries * 2; ++i) b[i] = __half2float(h[i]);
        } else CK(cudaMemcpy(b.data(), d_g32, entries * 8, cudaMemcpyDeviceToHost));
        double mx = 0, err = 0; int64_t nz = 0;
        for (int64_t i = 0; i < entries * 2; ++i) { mx = std::max(mx, (double)fabsf(a[i])); err = std::max(err, (double)fabsf(a[i] - b[i])); nz += a[i] != 0.f; }
        printf("   %-28s max|ref| %.3e  max err %.3e (%.2e rel)  touched %.1f%% of the table\n", name, mx, err, err / mx, 100.0 * nz / (entries * 2));
    };

    float t0 = time_it([&] { mfn_grid_encode_bwd(d_x, d_dy, &cfg, n, d_ref, 0); }, d_ref, entries * 8);
    printf("V0 library (level-major, strided dL_dout, float2 red)      %8.1f us\n", t0);
    float t1 = time_it([&] { v1_kernel<<<grid, 256>>>(d_x, (const __half2*)d_dyT, m, n, d_g32); }, d_g32, entries * 8);
    printf("V1 transposed dL_dout, float2 red                          %8.1f us\n", t1); check("V1", false);
    float t2 = time_it([&] { v2_kernel<<<grid, 256>>>(d_x, (const __half2*)d_dyT, m, n, (__half2*)d_g16); }, d_g16, entries * 4);
    printf("V2 transposed dL_dout, half2 red                           %8.1f us\n", t2); check("V2", true);
    for (int mh : {1, 2, 3}) {
        float t3 = time_it([&] { v3_kernel<false><<<grid, 256>>>(d_x, (const __half2*)d_dyT, m, n, d_g32, mh); }, d_g32, entries * 8);
        printf("V3 run-aggregated (when live >= %d*heads), float2 red       %8.1f us\n", mh, t3); check("V3", false);
    }
    float t4 = time_it([&] { v3_kernel<true><<<grid, 256>>>(d_x, (const __half2*)d_dyT, m, n, d_g16, 2); }, d_g16, entries * 4);
    printf("V4 run-aggregated (2), half2 red                           %8.1f us\n", t4); check("V4", true);
    // raw atomic rates: same number of lane-atomics as the real problem (alive * 16 levels * 8 corners), random over one 2^19-entry level
    const int64_t n_thr = alive * 16;
    for (uint32_t span : {1u << 19, (uint32_t)entries}) {
        float a0 = time_it([&] { v5_kernel<0><<<148 * 8, 256>>>(d_g32, span, n_thr, 8); }, d_g32, entries * 8);
        float a1 = time_it([&] { v5_kernel<1><<<148 * 8, 256>>>(d_g16, span, n_thr, 8); }, d_g16, entries * 4);
        float a2 = time_it([&] { v5_kernel<2><<<148 * 8, 256>>>(d_g32, span, n_thr, 8); }, d_g32, entries * 8);
        printf("raw random red over %8u entries, %lld lane-atomics: float2 %8.1f us (%.1f G/s)  half2 %8.1f us (%.1f G/s)  float %8.1f us (%.1f G/s)\n", span,
               (long long)n_thr * 8, a0, n_thr * 8 / a0 * 1e-3, a1, n_thr * 8 / a1 * 1e-3, a2, n_thr * 8 / a2 * 1e-3);
    }
    return 0;
}
