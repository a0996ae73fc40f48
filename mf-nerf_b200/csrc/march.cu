// raymarching_train / raymarching_test  (ref: models/csrc/raymarching.cu:166-332, 335-454)
//
// Train marcher: no atomics, deterministic layout (row i of rays_a is ray i and start_idx is the exclusive
// prefix sum of N_samples -- a legal instance of the reference's atomic-arrival order, raymarching.cu:237-241):
//   1. march_count_warp_kernel : one warp per ray (march.cuh), stashes (t, dt) of every accepted sample
//   2. march_scan_write_kernel : every CTA sums the counts in front of its rays, writes rays_a / counter and expands the
//                                stash into xyzs / dirs / deltas / ts (coalesced, exactly N rows)       [mfn_raymarching_train]
//      or scan_rays_kernel + march_write_kernel when the caller wants the count first            [mfn_march_train_count/_write]
// Algorithmic bytes: 60 B/ray + 32 B/sample (+ 8 B/sample stash write+read, + C*G^3/8 B of bitfield).
#include "march.cuh"
#include "../../include/mfnerf_b200.h"
#include <stdlib.h>

namespace mfn {

constexpr int kMarchWarpsPerCta = 8;     // march_write: one warp per ray
constexpr int kCountThreads = 128;
constexpr int kWsHeader = 256;           // workspace: [header: u64 @8 total samples | u64 @16 call counter][counts][stash]

// one warp per ray (march.cuh: march_ray_warp_fast, and march_ray_warp for the few rays whose visit order cannot be proven)
template <bool ONE_CASCADE, bool CONST_DT>
__global__ void __launch_bounds__(kCountThreads)
march_count_warp_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float* __restrict__ hits_t,
                        const uint8_t* __restrict__ bitfield, int cascades, int grid_size, float scale, float esf,
                        const float* __restrict__ noise, int max_samples, int64_t n_rays, int32_t* __restrict__ counts, float2* __restrict__ stash,
                        int fast, unsigned long long* __restrict__ header) {
    __shared__ uint32_t lut[1024];
    morton_lut_fill(lut, grid_size, threadIdx.x, kCountThreads);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (kCountThreads / 32) + (threadIdx.x >> 5);
    if (r >= n_rays) return;
    MarchConst c = make_march_const(cascades, grid_size, scale, esf, max_samples, scale);
    c.lut = lut;
    const RayConst q = make_ray(rays_o, rays_d, r);
    float t1 = hits_t[2 * r];
    const float t2 = hits_t[2 * r + 1];
    if (t1 >= 0.0f) t1 = __fmaf_rn(march_dt(t1, c), noise[r], t1);  // only the first sample is jittered (l.195-198)
    float2* my = stash + r * (int64_t)max_samples;
    auto emit = [&](int rank, float t, float dt) { my[rank] = make_float2(t, dt); };
    int n = fast ? march_ray_warp_fast<ONE_CASCADE, CONST_DT, true>(t1, t2, max_samples, q, c, bitfield, lane, emit) : -1;
    if (n < 0) {      // exact replay of the reference's visit order (overwrites whatever the fast attempt stashed)
        if (fast && lane == 0) atomicAdd(header + 3, 1ull);      // statistics: rays re-marched
        n = march_ray_warp<ONE_CASCADE, CONST_DT>(t1, t2, max_samples, q, c, bitfield, lane, emit);
    }
    if (lane == 0) counts[r] = n;
}

constexpr int kScanThreads = 1024;
constexpr int kScanItems = 4;

__global__ void __launch_bounds__(kScanThreads)
scan_rays_kernel(const int32_t* __restrict__ counts, int64_t n_rays, int64_t* __restrict__ rays_a, int32_t* __restrict__ counter) {
    __shared__ int64_t warp_tot[kScanThreads / 32];
    __shared__ int64_t carry_s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < n_rays; base += (int64_t)kScanThreads * kScanItems) {
        const int64_t first = base + (int64_t)tid * kScanItems;
        int v[kScanItems];
        int local = 0;
#pragma unroll
        for (int i = 0; i < kScanItems; ++i) {
            v[i] = (first + i < n_rays) ? counts[first + i] : 0;
            local += v[i];
        }
        const int incl = warp_incl_scan_i(local, lane);
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        int64_t before = carry_s;
        for (int w = 0; w < wid; ++w) before += warp_tot[w];
        int64_t start = before + (incl - local);
#pragma unroll
        for (int i = 0; i < kScanItems; ++i) {
            if (first + i < n_rays) {
                int64_t* row = rays_a + 3 * (first + i);
                row[0] = first + i; row[1] = start; row[2] = v[i];
            }
            start += v[i];
        }
        __syncthreads();
        if (tid == kScanThreads - 1) carry_s = start;  // last thread holds the running total
        __syncthreads();
    }
    if (tid == 0) { counter[0] = (int32_t)carry_s; counter[1] = (int32_t)n_rays; }
}

__global__ void __launch_bounds__(kMarchWarpsPerCta * 32)
march_write_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, const int64_t* __restrict__ rays_a,
                   const float2* __restrict__ stash, int max_samples, int64_t n_rays, int64_t capacity,
                   float* __restrict__ xyzs, float* __restrict__ dirs, float* __restrict__ deltas, float* __restrict__ ts) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * kMarchWarpsPerCta + (threadIdx.x >> 5);
    if (r >= n_rays) return;
    const int64_t start = rays_a[3 * r + 1];
    const int n = (int)rays_a[3 * r + 2];
    if (n == 0) return;
    const float ox = rays_o[3 * r], oy = rays_o[3 * r + 1], oz = rays_o[3 * r + 2];
    const float dx = rays_d[3 * r], dy = rays_d[3 * r + 1], dz = rays_d[3 * r + 2];
    const float2* my = stash + r * (int64_t)max_samples;
    for (int s = lane; s < n; s += 32) {
        const int64_t o = start + s;
        if (o >= capacity) break;
        const float2 td = my[s];
        xyzs[3 * o] = __fmaf_rn(dx, td.x, ox); xyzs[3 * o + 1] = __fmaf_rn(dy, td.x, oy); xyzs[3 * o + 2] = __fmaf_rn(dz, td.x, oz);
        dirs[3 * o] = dx; dirs[3 * o + 1] = dy; dirs[3 * o + 2] = dz;
        ts[o] = td.x; deltas[o] = td.y;
    }
}

// scan + write in one kernel (mfn_raymarching_train): every CTA first sums the counts of all rays in front of its own 8 rays
// (<= 32 KiB of L2-resident int32 per CTA, read with 16-byte loads), which replaces the single-CTA prefix-sum kernel and its
// launch; then writes rays_a and expands the stash exactly like march_write_kernel.  The CTA owning the last ray publishes the
// total in counter[0].
__global__ void __launch_bounds__(kMarchWarpsPerCta * 32)
march_scan_write_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, const int32_t* __restrict__ counts,
                        const float2* __restrict__ stash, int max_samples, int64_t n_rays, int64_t capacity, int64_t* __restrict__ rays_a,
                        int32_t* __restrict__ counter, float* __restrict__ xyzs, float* __restrict__ dirs, float* __restrict__ deltas,
                        float* __restrict__ ts, unsigned long long* __restrict__ header) {
    __shared__ long long warp_sum[kMarchWarpsPerCta];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t first = (int64_t)blockIdx.x * kMarchWarpsPerCta;        // first ray of this CTA (a multiple of 8 -> 16-byte aligned counts)
    long long acc = 0;
    const int4* c4 = reinterpret_cast<const int4*>(counts);
    for (int64_t i = tid; i < first / 4; i += kMarchWarpsPerCta * 32) { const int4 v = __ldg(c4 + i); acc += (long long)v.x + v.y + v.z + v.w; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) warp_sum[warp] = acc;
    __syncthreads();
    long long start = 0;
#pragma unroll
    for (int w = 0; w < kMarchWarpsPerCta; ++w) start += warp_sum[w];
    const int64_t r = first + warp;
    for (int w = 0; w < warp; ++w) start += (first + w < n_rays) ? counts[first + w] : 0;
    if (r >= n_rays) return;
    int n = counts[r];
    // a caller that sized the sample arrays below the worst case (n_rays * max_samples, what the reference allocates) gets the
    // overflowing rays truncated CONSISTENTLY: rays_a and the counter never point past `capacity`
    if (start >= capacity) { start = capacity; n = 0; }
    else if (start + n > capacity) n = (int)(capacity - start);
    if (lane == 0) {
        int64_t* row = rays_a + 3 * r;
        row[0] = r; row[1] = start; row[2] = n;
        if (r == n_rays - 1) {
            counter[0] = (int32_t)(start + n); counter[1] = (int32_t)n_rays;
            header[1] += (unsigned long long)(start + n);     // running total of marched samples (statistics without an extra kernel)
            header[2] += 1ull;                                // call counter: seeds the next mfn_ray_setup's jitter
        }
    }
    if (n == 0) return;
    const float ox = rays_o[3 * r], oy = rays_o[3 * r + 1], oz = rays_o[3 * r + 2];
    const float dx = rays_d[3 * r], dy = rays_d[3 * r + 1], dz = rays_d[3 * r + 2];
    const float2* my = stash + r * (int64_t)max_samples;
    // four stash loads in flight per lane: the kernel's duration is its longest ray's chain of (load -> stores) rounds
    for (int s0 = lane; s0 < n; s0 += 128) {
        float2 td[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int s = s0 + 32 * u; td[u] = s < n ? my[s] : make_float2(0.f, 0.f); }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int s = s0 + 32 * u;
            const int64_t o = start + s;
            if (s < n && o < capacity) {
                xyzs[3 * o] = __fmaf_rn(dx, td[u].x, ox); xyzs[3 * o + 1] = __fmaf_rn(dy, td[u].x, oy); xyzs[3 * o + 2] = __fmaf_rn(dz, td[u].x, oz);
                dirs[3 * o] = dx; dirs[3 * o + 1] = dy; dirs[3 * o + 2] = dz;
                ts[o] = td[u].x; deltas[o] = td[u].y;
            }
        }
    }
}

// test-time marcher: one thread per alive ray, sequential control flow of the reference; writes the zero
// padding itself so the caller needs no memset.  (ref: raymarching.cu:353-403)
__global__ void march_test_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, float* __restrict__ hits_t,
                                  const int64_t* __restrict__ alive, const uint8_t* __restrict__ bitfield, int cascades,
                                  int grid_size, float scale, float esf, int max_samples, int n_samples, int64_t n_alive,
                                  float* __restrict__ xyzs, float* __restrict__ dirs, float* __restrict__ deltas,
                                  float* __restrict__ ts, int32_t* __restrict__ n_eff) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_alive) return;
    const int64_t r = alive[n];
    const MarchConst c = make_march_const(cascades, grid_size, scale, esf, max_samples, (float)cascades);
    const RayConst q = make_ray(rays_o, rays_d, r);
    const float t = hits_t[2 * r], t2 = hits_t[2 * r + 1];
    float* px = xyzs + n * (int64_t)n_samples * 3;
    float* pd = dirs + n * (int64_t)n_samples * 3;
    float* pdt = deltas + n * (int64_t)n_samples;
    float* pt = ts + n * (int64_t)n_samples;
    float t_after;
    const int s = march_ray_thread(t, t2, n_samples, q, c, bitfield,
                                   [&](int k, float tk, float dt, float x, float y, float z) {
                                       px[3 * k] = x; px[3 * k + 1] = y; px[3 * k + 2] = z;
                                       pd[3 * k] = q.dx; pd[3 * k + 1] = q.dy; pd[3 * k + 2] = q.dz;
                                       pt[k] = tk; pdt[k] = dt;
                                   }, &t_after);
    if (s > 0) hits_t[2 * r] = t_after;  // next call resumes after the last accepted sample (l.390)
    for (int k = s; k < n_samples; ++k) {
        px[3 * k] = 0.f; px[3 * k + 1] = 0.f; px[3 * k + 2] = 0.f;
        pd[3 * k] = 0.f; pd[3 * k + 1] = 0.f; pd[3 * k + 2] = 0.f;
        pt[k] = 0.f; pdt[k] = 0.f;
    }
    n_eff[n] = s;
}

}  // namespace mfn

using namespace mfn;

extern "C" int64_t mfn_march_train_workspace_bytes(int64_t n_rays, int max_samples) {
    if (n_rays < 0 || max_samples < 1) return -1;
    // [header 256 B: statistics][counts: n_rays int32, padded to 256 B][stash: n_rays * max_samples float2]
    const int64_t counts = ((n_rays * 4 + 255) / 256) * 256;
    return kWsHeader + counts + n_rays * (int64_t)max_samples * 8;
}

static bool march_args_ok(int cascades, int grid_size, int max_samples, const char* name) {
    if (cascades < 1 || grid_size < 1 || grid_size > 1024 || max_samples < 1 || (int64_t)cascades * grid_size * grid_size * grid_size > 0xffffffffLL) {
        set_error("%s: bad cascades/grid_size/max_samples (%d, %d, %d)", name, cascades, grid_size, max_samples);
        return false;
    }
    return true;
}

static int march_count_impl(const float* rays_o, const float* rays_d, const float* hits_t, const uint8_t* bitfield,
                            int cascades, float scale, float exp_step_factor, const float* noise, int grid_size,
                            int max_samples, int64_t n_rays, int64_t* rays_a, int32_t* counter, void* workspace,
                            int64_t workspace_bytes, void* stream, bool with_scan) {
    if (!march_args_ok(cascades, grid_size, max_samples, "mfn_march_train_count")) return MFN_ERR_ARG;
    if (n_rays < 0 || !counter) { set_error("mfn_march_train_count: bad argument"); return MFN_ERR_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    if (n_rays == 0) { cudaMemsetAsync(counter, 0, 8, st); return check_launch("mfn_march_train_count", st); }
    if (!rays_o || !rays_d || !hits_t || !bitfield || !noise || !rays_a || !workspace) { set_error("mfn_march_train_count: null pointer"); return MFN_ERR_ARG; }
    if (workspace_bytes < mfn_march_train_workspace_bytes(n_rays, max_samples)) { set_error("mfn_march_train_count: workspace too small"); return MFN_ERR_ARG; }
    int32_t* counts = (int32_t*)((char*)workspace + kWsHeader);
    float2* stash = (float2*)((char*)workspace + kWsHeader + ((n_rays * 4 + 255) / 256) * 256);
    {
        ProfScope ps("march_count", st);
        const unsigned wb = (unsigned)ceil_div(n_rays, kCountThreads / 32);
        const bool one = cascades == 1, cdt = exp_step_factor == 0.0f;
        const char* fenv = getenv("MFN_MARCH_FAST");      // "0": exact replay for every ray (A/B runs, tests); read per call so that a test can flip it
        const int fast = !(fenv && fenv[0] == '0');
#define MFN_MW(A, B) march_count_warp_kernel<A, B><<<wb, kCountThreads, 0, st>>>(rays_o, rays_d, hits_t, bitfield, cascades, grid_size, scale, exp_step_factor, \
                                                                                noise, max_samples, n_rays, counts, stash, fast, (unsigned long long*)workspace)
        if (one && cdt) MFN_MW(true, true); else if (one) MFN_MW(true, false); else if (cdt) MFN_MW(false, true); else MFN_MW(false, false);
#undef MFN_MW
    }
    if (with_scan) {
        note_launch(1);
        ProfScope ps("march_scan", st);
        scan_rays_kernel<<<1, kScanThreads, 0, st>>>(counts, n_rays, rays_a, counter);
    }
    return check_launch("mfn_march_train_count", st);
}

extern "C" int mfn_march_train_count(const float* rays_o, const float* rays_d, const float* hits_t, const uint8_t* bitfield,
                                     int cascades, float scale, float exp_step_factor, const float* noise, int grid_size,
                                     int max_samples, int64_t n_rays, int64_t* rays_a, int32_t* counter, void* workspace,
                                     int64_t workspace_bytes, void* stream) {
    return march_count_impl(rays_o, rays_d, hits_t, bitfield, cascades, scale, exp_step_factor, noise, grid_size, max_samples, n_rays, rays_a, counter,
                            workspace, workspace_bytes, stream, true);
}

extern "C" int mfn_march_train_write(const float* rays_o, const float* rays_d, const int64_t* rays_a, const void* workspace,
                                     int max_samples, int64_t n_rays, int64_t capacity, float* xyzs, float* dirs, float* deltas,
                                     float* ts, void* stream) {
    if (n_rays < 0 || capacity < 0 || max_samples < 1) { set_error("mfn_march_train_write: bad argument"); return MFN_ERR_ARG; }
    if (n_rays == 0 || capacity == 0) return MFN_OK;
    if (!rays_o || !rays_d || !rays_a || !workspace || !xyzs || !dirs || !deltas || !ts) { set_error("mfn_march_train_write: null pointer"); return MFN_ERR_ARG; }
    const float2* stash = (const float2*)((const char*)workspace + kWsHeader + ((n_rays * 4 + 255) / 256) * 256);
    const int blocks = (int)ceil_div(n_rays, kMarchWarpsPerCta);
    ProfScope ps("march_write", (cudaStream_t)stream);
    march_write_kernel<<<blocks, kMarchWarpsPerCta * 32, 0, (cudaStream_t)stream>>>(rays_o, rays_d, rays_a, stash, max_samples, n_rays,
                                                                                    capacity, xyzs, dirs, deltas, ts);
    return check_launch("mfn_march_train_write", (cudaStream_t)stream);
}

extern "C" int mfn_raymarching_train(const float* rays_o, const float* rays_d, const float* hits_t, const uint8_t* bitfield,
                                     int cascades, float scale, float exp_step_factor, const float* noise, int grid_size,
                                     int max_samples, int64_t n_rays, int64_t capacity, int64_t* rays_a, float* xyzs, float* dirs,
                                     float* deltas, float* ts, int32_t* counter, void* workspace, int64_t workspace_bytes, void* stream) {
    int rc = march_count_impl(rays_o, rays_d, hits_t, bitfield, cascades, scale, exp_step_factor, noise, grid_size, max_samples,
                              n_rays, rays_a, counter, workspace, workspace_bytes, stream, false);
    if (rc != MFN_OK || n_rays == 0) return rc;
    if (capacity < 0 || (capacity > 0 && (!xyzs || !dirs || !deltas || !ts))) { set_error("mfn_raymarching_train: null pointer"); return MFN_ERR_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    const int32_t* counts = (const int32_t*)((const char*)workspace + kWsHeader);
    const float2* stash = (const float2*)((const char*)workspace + kWsHeader + ((n_rays * 4 + 255) / 256) * 256);
    const int blocks = (int)ceil_div(n_rays, kMarchWarpsPerCta);
    ProfScope ps("march_write", st);
    march_scan_write_kernel<<<blocks, kMarchWarpsPerCta * 32, 0, st>>>(rays_o, rays_d, counts, stash, max_samples, n_rays, capacity, rays_a, counter,
                                                                       xyzs, dirs, deltas, ts, (unsigned long long*)workspace);
    return check_launch("mfn_raymarching_train", st);
}

extern "C" int mfn_raymarching_test(const float* rays_o, const float* rays_d, float* hits_t, const int64_t* alive_indices,
                                    const uint8_t* bitfield, int cascades, float scale, float exp_step_factor, int grid_size,
                                    int max_samples, int n_samples, int64_t n_alive, float* xyzs, float* dirs, float* deltas,
                                    float* ts, int32_t* n_eff_samples, void* stream) {
    if (!march_args_ok(cascades, grid_size, max_samples, "mfn_raymarching_test")) return MFN_ERR_ARG;
    if (n_alive < 0 || n_samples < 0) { set_error("mfn_raymarching_test: bad argument"); return MFN_ERR_ARG; }
    if (n_alive == 0) return MFN_OK;
    if (!rays_o || !rays_d || !hits_t || !alive_indices || !bitfield || !n_eff_samples || (n_samples > 0 && (!xyzs || !dirs || !deltas || !ts))) {
        set_error("mfn_raymarching_test: null pointer"); return MFN_ERR_ARG;
    }
    const int threads = 128;
    march_test_kernel<<<(int)ceil_div(n_alive, threads), threads, 0, (cudaStream_t)stream>>>(
        rays_o, rays_d, hits_t, alive_indices, bitfield, cascades, grid_size, scale, exp_step_factor, max_samples, n_samples, n_alive,
        xyzs, dirs, deltas, ts, n_eff_samples);
    return check_launch("mfn_raymarching_test", (cudaStream_t)stream);
}
