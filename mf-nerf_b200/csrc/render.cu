// Test-time rendering without host round trips: the reference's `__render_rays_test` while-loop (models/rendering.py:46-118)
// as a device-side wavefront.  The reference synchronises three times per iteration (alive count -> N_samples, valid-mask
// gather, alive[alive >= 0]); here the alive list, its length, the per-iteration sample count and the loop's sample budget all
// live in device memory:
//   begin      : AABB slab test + near clamp (rendering.py:27-29), alive = [0 .. N), outputs zeroed, plan initialised
//   iteration  : plan (N_samples = clamp(N_rays / N_alive, min_samples, 64), rendering.py:76-77; stop once the budget of
//                max_samples is spent, l.69) -> march (raymarching.cu:335-454 semantics, one thread per alive ray)
//                -> field (field_fused.cu, inference mode, sample count read from the plan)
//                -> composite (volumerendering.cu:205-285 semantics) + compaction of the still-alive rays into the other list
//                   (ballot + one atomicAdd per warp)
// The host enqueues iterations in batches and looks at the alive count only between batches.
// Results equal the reference loop's: every ray consumes its samples in the same order whatever the batching (the compositor
// checks T <= T_threshold after every sample), and the order of rays inside the alive list does not enter the arithmetic.
#include "field_internal.h"
#include "march.cuh"

namespace mfn {

struct RenderPlan {
    int32_t n_alive[2];      // length of alive list 0 / 1
    int32_t cur;             // which list is current
    int32_t n_samples;       // N_samples of this iteration
    int32_t n_rows;          // n_alive * n_samples: rows the field kernel evaluates
    int32_t spent;           // sum of N_samples so far (the reference's `samples`)
    int32_t iterations, rows_total_k;   // diagnostics: iterations run, field rows evaluated (in units of 1024)
    unsigned long long total_eff;   // sum of N_eff (the reference's total_samples)
    unsigned int ticket;            // blocks of the compositor that have finished this iteration
};

constexpr int kPlanBytes = 256;

__global__ void render_begin_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, float cx, float cy, float cz, float hx, float hy,
                                    float hz, float near_distance, int n_rays, float* __restrict__ hits_t, int32_t* __restrict__ alive0,
                                    float* __restrict__ opacity, float* __restrict__ depth, float* __restrict__ rgb, RenderPlan* __restrict__ plan) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r == 0) {
        plan->n_alive[0] = n_rays; plan->n_alive[1] = 0; plan->cur = 0; plan->n_samples = 0; plan->n_rows = 0; plan->spent = 0; plan->total_eff = 0ull; plan->iterations = 0; plan->rows_total_k = 0; plan->ticket = 0u;
    }
    if (r >= n_rays) return;
    // ray / AABB slab test exactly as intersect.cu (ref: intersection.cu:5-22, 48-54)
    const float ox = rays_o[3 * r], oy = rays_o[3 * r + 1], oz = rays_o[3 * r + 2];
    const float ix = __fdiv_rn(1.0f, rays_d[3 * r]), iy = __fdiv_rn(1.0f, rays_d[3 * r + 1]), iz = __fdiv_rn(1.0f, rays_d[3 * r + 2]);
    const float tminx = __fmul_rn(__fsub_rn(__fsub_rn(cx, hx), ox), ix), tmaxx = __fmul_rn(__fsub_rn(__fadd_rn(cx, hx), ox), ix);
    const float tminy = __fmul_rn(__fsub_rn(__fsub_rn(cy, hy), oy), iy), tmaxy = __fmul_rn(__fsub_rn(__fadd_rn(cy, hy), oy), iy);
    const float tminz = __fmul_rn(__fsub_rn(__fsub_rn(cz, hz), oz), iz), tmaxz = __fmul_rn(__fsub_rn(__fadd_rn(cz, hz), oz), iz);
    float t1 = fmaxf(fmaxf(fminf(tminx, tmaxx), fminf(tminy, tmaxy)), fminf(tminz, tmaxz));
    float t2 = fminf(fminf(fmaxf(tminx, tmaxx), fmaxf(tminy, tmaxy)), fmaxf(tminz, tmaxz));
    if (t1 > t2 || !(t2 > 0.f)) { t1 = -1.f; t2 = -1.f; }      // miss, or box behind the ray
    else {
        t1 = fmaxf(t1, 0.f);
        if (t1 >= 0.f && t1 < near_distance) t1 = near_distance;   // rendering.py:29
    }
    hits_t[2 * r] = t1; hits_t[2 * r + 1] = t2;
    alive0[r] = r;
    opacity[r] = 0.f; depth[r] = 0.f; rgb[3 * r] = 0.f; rgb[3 * r + 1] = 0.f; rgb[3 * r + 2] = 0.f;
}

// one thread: N_samples for the coming iteration, budget bookkeeping, reset of the list that its compositor is going to fill
__device__ void render_make_plan(RenderPlan* __restrict__ plan, int n_rays, int min_samples, int max_samples, int cap_rows) {
    const int cur = plan->cur;
    int na = plan->n_alive[cur];
    if (plan->spent >= max_samples) na = 0;                       // rendering.py:69  while samples < max_samples
    int ns = 0;
    if (na > 0) {
        ns = max(min(n_rays / na, 64), min_samples);              // rendering.py:76-77
        if ((long long)na * ns > cap_rows) ns = max(cap_rows / na, 1);
        plan->spent += ns;
    }
    plan->n_alive[cur] = na;
    plan->n_samples = ns;
    plan->n_rows = na * ns;
    plan->n_alive[cur ^ 1] = 0;
    if (na > 0) { plan->iterations += 1; plan->rows_total_k += (na * ns + 1023) / 1024; }
}
__global__ void render_plan_kernel(RenderPlan* __restrict__ plan, int n_rays, int min_samples, int max_samples, int cap_rows) {
    render_make_plan(plan, n_rays, min_samples, max_samples, cap_rows);
}

template <bool ONE_CASCADE>
__global__ void __launch_bounds__(128)
render_march_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, float* __restrict__ hits_t, const int32_t* __restrict__ alive_lists,
                    int list_stride, const RenderPlan* __restrict__ plan, const uint8_t* __restrict__ bitfield, int cascades, int grid_size, float scale,
                    float esf, int max_samples, float* __restrict__ deltas, float* __restrict__ ts, int32_t* __restrict__ n_eff) {
    const int cur = plan->cur, na = plan->n_alive[cur], ns = plan->n_samples;
    if ((int)(blockIdx.x * blockDim.x) >= na) return;
    __shared__ uint32_t lut[1024];
    morton_lut_fill(lut, grid_size, threadIdx.x, 128);
    __syncthreads();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= na) return;
    const int r = alive_lists[cur * list_stride + n];
    MarchConst c = make_march_const(cascades, grid_size, scale, esf, max_samples, (float)cascades);   // the test marcher's calc_dt quirk
    c.lut = lut;
    const RayConst q = make_ray(rays_o, rays_d, r);
    const float t = hits_t[2 * r], t2 = hits_t[2 * r + 1];
    // only t and dt of a sample are stored (8 B per row): the field kernel derives x = o + t d and the direction from the ray, and
    // neither it nor the compositor looks at rows past N_eff -- no padding is written
    float* pdt = deltas + (size_t)n * ns;
    float* pt = ts + (size_t)n * ns;
    float t_after;
    const int s = march_ray_thread<ONE_CASCADE, true>(t, t2, ns, q, c, bitfield, [&](int k, float tk, float dt, float, float, float) { pt[k] = tk; pdt[k] = dt; }, &t_after);
    if (s > 0) hits_t[2 * r] = t_after;
    n_eff[n] = s;
}

__device__ __forceinline__ float alpha_test(float sigma, float delta) { return __fadd_rn(1.0f, -__expf(-__fmul_rn(sigma, delta))); }

__global__ void __launch_bounds__(128)
render_composite_kernel(const float* __restrict__ sigmas, const float* __restrict__ rgbs, const float* __restrict__ deltas, const float* __restrict__ ts,
                        int32_t* __restrict__ alive_lists, int list_stride, RenderPlan* __restrict__ plan, float T_thr, const int32_t* __restrict__ n_eff,
                        float* __restrict__ opacity, float* __restrict__ depth, float* __restrict__ rgb, int n_rays, int min_samples, int max_samples,
                        int cap_rows) {
    __shared__ int s_cur, s_na, s_ns;
    if (threadIdx.x == 0) { s_cur = plan->cur; s_na = plan->n_alive[plan->cur]; s_ns = plan->n_samples; }   // read before the last block re-plans
    __syncthreads();
    const int cur = s_cur, na = s_na, ns = s_ns;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    bool keep = false;
    int r = 0, ne = 0;
    if (n < na) {
        ne = n_eff[n];
        r = alive_lists[cur * list_stride + n];
        if (ne > 0) {      // (N_eff == 0: the ray left the box -> dead, volumerendering.cu:223-226)
            float o = opacity[r], d = depth[r], cr = rgb[3 * r], cg = rgb[3 * r + 1], cb = rgb[3 * r + 2];
            float T = __fadd_rn(1.0f, -o);
            const size_t rowb = (size_t)n * ns;
            keep = true;
            for (int s = 0; s < ne; ++s) {
                const float a = alpha_test(sigmas[rowb + s], deltas[rowb + s]);
                const float w = __fmul_rn(a, T);
                cr = __fmaf_rn(w, rgbs[3 * (rowb + s)], cr);
                cg = __fmaf_rn(w, rgbs[3 * (rowb + s) + 1], cg);
                cb = __fmaf_rn(w, rgbs[3 * (rowb + s) + 2], cb);
                d = __fmaf_rn(w, ts[rowb + s], d);
                o = __fadd_rn(o, w);
                T = __fmul_rn(T, __fadd_rn(1.0f, -a));
                if (T <= T_thr) { keep = false; break; }
            }
            opacity[r] = o; depth[r] = d; rgb[3 * r] = cr; rgb[3 * r + 1] = cg; rgb[3 * r + 2] = cb;
        }
    }
    // compaction: survivors go to the other list
    const uint32_t km = __ballot_sync(0xffffffffu, keep);
    unsigned long long eff = (unsigned long long)ne;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) eff += __shfl_xor_sync(0xffffffffu, eff, o);
    int base = 0;
    if (lane == 0) {
        if (km) base = atomicAdd(&plan->n_alive[cur ^ 1], __popc(km));
        if (eff) atomicAdd(&plan->total_eff, eff);
    }
    base = __shfl_sync(0xffffffffu, base, 0);
    if (keep) alive_lists[(cur ^ 1) * list_stride + base + __popc(km & ((1u << lane) - 1u))] = r;
    // the last block to finish flips the lists and plans the next iteration (saves two single-thread launches per iteration)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&plan->ticket, 1u) == gridDim.x - 1) {
            __threadfence();
            plan->ticket = 0u;
            plan->cur = cur ^ 1;
            render_make_plan(plan, n_rays, min_samples, max_samples, cap_rows);
        }
    }
}

__global__ void render_finish_kernel(float* __restrict__ rgb, const float* __restrict__ opacity, float bg_r, float bg_g, float bg_b, int n_rays) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    const float k = 1.f - opacity[r];                              // rendering.py:112-116  rgb + bg * (1 - opacity)
    rgb[3 * r] += bg_r * k; rgb[3 * r + 1] += bg_g * k; rgb[3 * r + 2] += bg_b * k;
}

struct RenderWs { size_t plan, hits, alive, neff, deltas, ts, sigmas, rgbs, total; int64_t cap_rows; };
static RenderWs render_ws(int64_t n_rays, int min_samples) {
    RenderWs w{};
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    const int64_t cap = n_rays * (int64_t)(min_samples > 1 ? min_samples : 1);
    size_t o = 0;
    w.plan = o; o += kPlanBytes;
    w.hits = o; o += al(n_rays * 8);
    w.alive = o; o += al(n_rays * 4) * 2;
    w.neff = o; o += al(n_rays * 4);
    w.deltas = o; o += al(cap * 4);
    w.ts = o; o += al(cap * 4);
    w.sigmas = o; o += al(cap * 4);
    w.rgbs = o; o += al(cap * 12);
    w.total = o; w.cap_rows = cap;
    return w;
}

}  // namespace mfn

using namespace mfn;

extern "C" int64_t mfn_render_workspace_bytes(int64_t n_rays, int min_samples) {
    if (n_rays < 0 || n_rays > 0x3fffffff || min_samples < 1 || min_samples > 64) return -1;
    return (int64_t)render_ws(n_rays, min_samples).total;
}

extern "C" int mfn_render_begin(const float* rays_o, const float* rays_d, const float* center_host, const float* half_size_host, int64_t n_rays,
                                float near_distance, int min_samples, int max_samples, float* opacity, float* depth, float* rgb, void* workspace,
                                int64_t workspace_bytes, void* stream) {
    if (n_rays < 0 || n_rays > 0x3fffffff || min_samples < 1 || min_samples > 64 || max_samples < 1) { set_error("mfn_render_begin: bad argument"); return MFN_ERR_ARG; }
    if (n_rays == 0) return MFN_OK;
    const RenderWs w = render_ws(n_rays, min_samples);
    if (!rays_o || !rays_d || !center_host || !half_size_host || !opacity || !depth || !rgb || !workspace || (size_t)workspace_bytes < w.total) {
        set_error("mfn_render_begin: null pointer or workspace too small"); return MFN_ERR_ARG;
    }
    char* ws = (char*)workspace;
    cudaStream_t st = (cudaStream_t)stream;
    render_begin_kernel<<<(unsigned)ceil_div(n_rays, 256), 256, 0, st>>>(rays_o, rays_d, center_host[0], center_host[1], center_host[2], half_size_host[0],
                                                                         half_size_host[1], half_size_host[2], near_distance, (int)n_rays, (float*)(ws + w.hits),
                                                                         (int32_t*)(ws + w.alive), opacity, depth, rgb, (RenderPlan*)(ws + w.plan));
    render_plan_kernel<<<1, 1, 0, st>>>((RenderPlan*)(ws + w.plan), (int)n_rays, min_samples, max_samples, (int)w.cap_rows);   // plan of iteration 1
    note_launch(1);
    return check_launch("mfn_render_begin", st);
}

extern "C" int mfn_render_iterations(const mfn_field_cfg* cfg, const void* xyz_params_h, const void* rgb_params_h, const float* rays_o, const float* rays_d,
                                     int64_t n_rays, const uint8_t* density_bitfield, int cascades, float scale, float exp_step_factor, int grid_size,
                                     int max_samples, int min_samples, float T_threshold, int n_iterations, float* opacity, float* depth, float* rgb,
                                     void* workspace, int64_t workspace_bytes, void* stream) {
    if (!cfg || n_rays < 0 || n_rays > 0x3fffffff || min_samples < 1 || min_samples > 64 || n_iterations < 0 || cascades < 1 || grid_size < 1) {
        set_error("mfn_render_iterations: bad argument"); return MFN_ERR_ARG;
    }
    if (n_rays == 0 || n_iterations == 0) return MFN_OK;
    if (!fused_field_supported(cfg)) { set_error("mfn_render_iterations: field shape not covered by the fused kernels (L16 F2, 64-wide nets)"); return MFN_ERR_ARG; }
    const RenderWs w = render_ws(n_rays, min_samples);
    if (!xyz_params_h || !rgb_params_h || !rays_o || !rays_d || !density_bitfield || !opacity || !depth || !rgb || !workspace || (size_t)workspace_bytes < w.total) {
        set_error("mfn_render_iterations: null pointer or workspace too small"); return MFN_ERR_ARG;
    }
    GridMeta m;
    int rc = build_grid_meta(&cfg->grid, &m, "mfn_render_iterations");
    if (rc != MFN_OK) return rc;
    char* ws = (char*)workspace;
    cudaStream_t st = (cudaStream_t)stream;
    RenderPlan* plan = (RenderPlan*)(ws + w.plan);
    const int list_stride = (int)(((n_rays * 4 + 255) / 256 * 256) / 4);
    const unsigned ray_blocks = (unsigned)ceil_div(n_rays, 128);
    FusedArgs f{};
    f.n_max = w.cap_rows; f.n_dev = &plan->n_rows;
    f.ray_ts = (const float*)(ws + w.ts); f.rays_o = rays_o; f.rays_d = rays_d; f.ray_plan = (const int32_t*)plan; f.ray_alive_lists = (const int32_t*)(ws + w.alive);
    f.ray_n_eff = (const int32_t*)(ws + w.neff); f.ray_list_stride = list_stride;
    for (int k = 0; k < 3; ++k) { f.mn[k] = cfg->xyz_min[k]; f.mx[k] = cfg->xyz_max[k]; }
    f.w_sigma = (const __half*)xyz_params_h; f.table = f.w_sigma + 64 * 32 + 16 * 64; f.w_rgb = (const __half*)rgb_params_h; f.rgb_act = cfg->rgb_act;
    f.sigmas = (float*)(ws + w.sigmas); f.rgbs = (float*)(ws + w.rgbs);
    for (int it = 0; it < n_iterations; ++it) {
        (cascades == 1 ? render_march_kernel<true> : render_march_kernel<false>)<<<ray_blocks, 128, 0, st>>>(rays_o, rays_d, (float*)(ws + w.hits), (const int32_t*)(ws + w.alive), list_stride, plan, density_bitfield,
                                                        cascades, grid_size, scale, exp_step_factor, max_samples, (float*)(ws + w.deltas), (float*)(ws + w.ts),
                                                        (int32_t*)(ws + w.neff));
        note_launch(1);
        if ((rc = fused_field_forward(f, m, cfg->rgb_width, cfg->rgb_hidden, 4, st)) != MFN_OK) return rc;
        render_composite_kernel<<<ray_blocks, 128, 0, st>>>((const float*)(ws + w.sigmas), (const float*)(ws + w.rgbs), (const float*)(ws + w.deltas),
                                                            (const float*)(ws + w.ts), (int32_t*)(ws + w.alive), list_stride, plan, T_threshold,
                                                            (const int32_t*)(ws + w.neff), opacity, depth, rgb, (int)n_rays, min_samples, max_samples,
                                                            (int)w.cap_rows);
        note_launch(1);
    }
    return check_launch("mfn_render_iterations", st);
}

/* copies {alive rays, sum of N_samples spent, total effective samples (lo, hi)} to 4 int32 of HOST (pinned) memory, asynchronously */
extern "C" int mfn_render_status(const void* workspace, int32_t* status_host_pinned, void* stream) {
    if (!workspace || !status_host_pinned) { set_error("mfn_render_status: null pointer"); return MFN_ERR_ARG; }
    const RenderPlan* plan = (const RenderPlan*)workspace;
    // after the flip, `cur` names the list the last compositor filled; copy the whole plan head and let the host pick
    cudaError_t e = cudaMemcpyAsync(status_host_pinned, plan, 40, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("mfn_render_status: %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return MFN_ERR_CUDA; }
    return MFN_OK;
}

extern "C" int mfn_render_finish(float* rgb, const float* opacity, const float* bg_rgb_host, int64_t n_rays, void* stream) {
    if (n_rays < 0 || n_rays > 0x3fffffff) { set_error("mfn_render_finish: bad argument"); return MFN_ERR_ARG; }
    if (n_rays == 0) return MFN_OK;
    if (!rgb || !opacity || !bg_rgb_host) { set_error("mfn_render_finish: null pointer"); return MFN_ERR_ARG; }
    render_finish_kernel<<<(unsigned)ceil_div(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(rgb, opacity, bg_rgb_host[0], bg_rgb_host[1], bg_rgb_host[2], (int)n_rays);
    return check_launch("mfn_render_finish", (cudaStream_t)stream);
}
