// ray/AABB and ray/sphere intersection  (ref: models/csrc/intersection.cu:5-100, 103-197)
//
// The reference launches a (rays x voxels) grid, counts hits with atomics and then runs a torch sort +
// two gathers on the host side.  Here one thread owns one ray, walks the voxel list (it is 1 long on the
// render path: rendering.py:27-28 passes the scene box with max_hits=1), keeps the first `max_hits` hits
// in voxel order (a legal instance of the reference's atomic arrival order) and sorts them in place, so
// the op is a single launch and its output is deterministic.  All arithmetic is the reference's own
// expression tree in IEEE fp32 (no contraction is possible in (c-h-o)*inv_d), hence hits_t is bit-exact.
#include "common.cuh"
#include "../../include/mfnerf_b200.h"

namespace mfn {

__device__ __forceinline__ float2 slab_test(float ox, float oy, float oz, float ix, float iy, float iz,
                                            float cx, float cy, float cz, float hx, float hy, float hz) {
    // ref: intersection.cu:12-21
    const float tminx = __fmul_rn(__fsub_rn(__fsub_rn(cx, hx), ox), ix);
    const float tminy = __fmul_rn(__fsub_rn(__fsub_rn(cy, hy), oy), iy);
    const float tminz = __fmul_rn(__fsub_rn(__fsub_rn(cz, hz), oz), iz);
    const float tmaxx = __fmul_rn(__fsub_rn(__fadd_rn(cx, hx), ox), ix);
    const float tmaxy = __fmul_rn(__fsub_rn(__fadd_rn(cy, hy), oy), iy);
    const float tmaxz = __fmul_rn(__fsub_rn(__fadd_rn(cz, hz), oz), iz);
    const float t1 = fmaxf(fmaxf(fminf(tminx, tmaxx), fminf(tminy, tmaxy)), fminf(tminz, tmaxz));
    const float t2 = fminf(fminf(fmaxf(tminx, tmaxx), fmaxf(tminy, tmaxy)), fmaxf(tminz, tmaxz));
    if (t1 > t2) return make_float2(-1.f, -1.f);
    return make_float2(t1, t2);
}

__device__ __forceinline__ float2 sphere_test(float ox, float oy, float oz, float dx, float dy, float dz,
                                              float cx, float cy, float cz, float radius) {
    // ref: intersection.cu:109-120 (dot() from helper_math is a.x*b.x + a.y*b.y + a.z*b.z, contracted
    // left to right by nvcc: fma(z,z, fma(y,y, x*x)))
    const float px = ox - cx, py = oy - cy, pz = oz - cz;
    const float a = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    const float half_b = fmaf(dz, pz, fmaf(dy, py, dx * px));
    const float c = fmaf(pz, pz, fmaf(py, py, px * px)) - radius * radius;
    const float disc = half_b * half_b - a * c;
    if (disc < 0) return make_float2(-1.f, -1.f);
    const float s = sqrtf(disc);
    return make_float2((-half_b - s) / a, (-half_b + s) / a);
}

template <bool kSphere>
__global__ void intersect_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                 const float* __restrict__ centers, const float* __restrict__ extent,
                                 int64_t n_rays, int n_prims, int max_hits,
                                 int32_t* __restrict__ hit_cnt, float* __restrict__ hits_t,
                                 int64_t* __restrict__ hits_idx) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    const float ox = rays_o[3 * r], oy = rays_o[3 * r + 1], oz = rays_o[3 * r + 2];
    const float dx = rays_d[3 * r], dy = rays_d[3 * r + 1], dz = rays_d[3 * r + 2];
    const float ix = __fdiv_rn(1.0f, dx), iy = __fdiv_rn(1.0f, dy), iz = __fdiv_rn(1.0f, dz);
    float* ht = hits_t + r * (int64_t)max_hits * 2;
    int64_t* hi = hits_idx + r * (int64_t)max_hits;
    int cnt = 0;
    for (int v = 0; v < n_prims; ++v) {
        float2 t;
        if constexpr (kSphere) {
            t = sphere_test(ox, oy, oz, dx, dy, dz, centers[3 * v], centers[3 * v + 1], centers[3 * v + 2], extent[v]);
        } else {
            t = slab_test(ox, oy, oz, ix, iy, iz, centers[3 * v], centers[3 * v + 1], centers[3 * v + 2],
                          extent[3 * v], extent[3 * v + 1], extent[3 * v + 2]);
        }
        if (t.y > 0) {  // ref: intersection.cu:48-55
            if (cnt < max_hits) {
                // insertion keeps the kept hits ordered by t1 (near -> far), stable in voxel order
                const float t1 = fmaxf(t.x, 0.0f);
                int k = cnt;
                while (k > 0 && ht[2 * (k - 1)] > t1) {
                    ht[2 * k] = ht[2 * (k - 1)]; ht[2 * k + 1] = ht[2 * (k - 1) + 1]; hi[k] = hi[k - 1];
                    --k;
                }
                ht[2 * k] = t1; ht[2 * k + 1] = t.y; hi[k] = v;
            }
            ++cnt;
        }
    }
    hit_cnt[r] = cnt;
    // the reference sorts the whole (-1 padded) row ascending by t1, so the -1 padding ends up in FRONT of
    // the hits when a ray has fewer than max_hits of them (intersection.cu:95-97).
    const int kept = cnt < max_hits ? cnt : max_hits;
    const int pad = max_hits - kept;
    if (pad > 0) {
        for (int k = kept - 1; k >= 0; --k) {
            ht[2 * (k + pad)] = ht[2 * k]; ht[2 * (k + pad) + 1] = ht[2 * k + 1]; hi[k + pad] = hi[k];
        }
        for (int k = 0; k < pad; ++k) { ht[2 * k] = -1.f; ht[2 * k + 1] = -1.f; hi[k] = -1; }
    }
}

}  // namespace mfn

using namespace mfn;

static int launch_intersect(bool sphere, const float* o, const float* d, const float* c, const float* e, int64_t n_rays,
                            int64_t n_prims, int max_hits, int32_t* cnt, float* ht, int64_t* hi, void* stream, const char* name) {
    if (n_rays < 0 || n_prims < 0 || max_hits < 1 || n_prims > 0x7fffffff) { set_error("%s: bad sizes", name); return MFN_ERR_ARG; }
    if (n_rays == 0) return MFN_OK;
    if (!o || !d || !cnt || !ht || !hi || (n_prims > 0 && (!c || !e))) { set_error("%s: null pointer", name); return MFN_ERR_ARG; }
    const int threads = 128;
    const int blocks = (int)ceil_div(n_rays, threads);
    if (sphere) intersect_kernel<true><<<blocks, threads, 0, (cudaStream_t)stream>>>(o, d, c, e, n_rays, (int)n_prims, max_hits, cnt, ht, hi);
    else intersect_kernel<false><<<blocks, threads, 0, (cudaStream_t)stream>>>(o, d, c, e, n_rays, (int)n_prims, max_hits, cnt, ht, hi);
    return check_launch(name, (cudaStream_t)stream);
}

namespace mfn {
// Engine front end of a training step in ONE launch: ray / scene-box slab test with max_hits = 1 (rendering.py:27-28 ->
// intersection.cu:5-22, 48-54), the near-plane clamp of rendering.py:29, and the per-ray jitter the reference draws with
// torch.rand_like inside RayMarcher.forward (custom_functions.py:83).  The jitter is counter-based (splitmix64 of a device-side
// call counter and the ray index), so a replayed CUDA graph draws fresh noise every step without a host-side generator.
__global__ void ray_setup_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, float cx, float cy, float cz, float hx, float hy, float hz,
                                 float near_distance, int64_t n_rays, unsigned long long* __restrict__ call_counter, float* __restrict__ hits_t,
                                 float* __restrict__ noise) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long call = call_counter ? *call_counter : 0ull;   // only read here; bumped by mfn_raymarching_train
    if (r < n_rays) {
        const float ox = rays_o[3 * r], oy = rays_o[3 * r + 1], oz = rays_o[3 * r + 2];
        const float ix = __fdiv_rn(1.0f, rays_d[3 * r]), iy = __fdiv_rn(1.0f, rays_d[3 * r + 1]), iz = __fdiv_rn(1.0f, rays_d[3 * r + 2]);
        const float2 t = slab_test(ox, oy, oz, ix, iy, iz, cx, cy, cz, hx, hy, hz);
        float t1 = -1.f, t2 = -1.f;
        if (t.y > 0.f) {                       // (a miss is (-1, -1): fails this test too)
            t1 = fmaxf(t.x, 0.f); t2 = t.y;
            if (t1 >= 0.f && t1 < near_distance) t1 = near_distance;
        }
        hits_t[2 * r] = t1; hits_t[2 * r + 1] = t2;
        if (noise && call_counter) {
            unsigned long long x = (call + 1ull) * 0x9E3779B97F4A7C15ull ^ (unsigned long long)r * 0xD1342543DE82EF95ull;
            x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; x ^= x >> 31;
            noise[r] = (float)(x >> 40) * (1.0f / 16777216.0f);       // [0, 1), 24 random bits like torch.rand
        }
    }
}
}  // namespace mfn

/* see include/mfnerf_b200.h */
extern "C" int mfn_ray_setup(const float* rays_o, const float* rays_d, const float* center_host, const float* half_size_host, int64_t n_rays,
                             float near_distance, void* call_counter, float* hits_t, float* noise, void* stream) {
    if (n_rays < 0) { set_error("mfn_ray_setup: bad argument"); return MFN_ERR_ARG; }
    if (n_rays == 0) return MFN_OK;
    if (!rays_o || !rays_d || !center_host || !half_size_host || !hits_t) { set_error("mfn_ray_setup: null pointer"); return MFN_ERR_ARG; }
    ray_setup_kernel<<<(unsigned)ceil_div(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, center_host[0], center_host[1], center_host[2],
                                                                                  half_size_host[0], half_size_host[1], half_size_host[2], near_distance,
                                                                                  n_rays, (unsigned long long*)call_counter, hits_t, noise);
    return check_launch("mfn_ray_setup", (cudaStream_t)stream);
}

extern "C" int mfn_ray_aabb_intersect(const float* rays_o, const float* rays_d, const float* centers, const float* half_sizes,
                                      int64_t n_rays, int64_t n_voxels, int max_hits, int32_t* hit_cnt, float* hits_t,
                                      int64_t* hits_voxel_idx, void* stream) {
    return launch_intersect(false, rays_o, rays_d, centers, half_sizes, n_rays, n_voxels, max_hits, hit_cnt, hits_t, hits_voxel_idx,
                            stream, "mfn_ray_aabb_intersect");
}

extern "C" int mfn_ray_sphere_intersect(const float* rays_o, const float* rays_d, const float* centers, const float* radii,
                                        int64_t n_rays, int64_t n_spheres, int max_hits, int32_t* hit_cnt, float* hits_t,
                                        int64_t* hits_sphere_idx, void* stream) {
    return launch_intersect(true, rays_o, rays_d, centers, radii, n_rays, n_spheres, max_hits, hit_cnt, hits_t, hits_sphere_idx,
                            stream, "mfn_ray_sphere_intersect");
}
