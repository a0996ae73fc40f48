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
