// Ray generation and batch sampling on the device.  Replaces, for a data set resident in HBM, the reference's
//   DataLoader item      datasets/base.py:22-34   img_idxs / pix_idxs = np.random.choice(...), rgb = self.rays[img_idxs, pix_idxs]
//   NeRFSystem.forward   train.py:83-96           poses[img_idxs], directions[pix_idxs] (train) or one pose + all directions (test)
//   get_ray_directions   datasets/ray_utils.py:23-35   ((u - cx + 0.5) / fx, (v - cy + 0.5) / fy, 1), u = pix % W, v = pix / W
//   get_rays             datasets/ray_utils.py:60-68   rays_d = directions @ c2w[:, :3]^T, rays_o = c2w[:, 3]
// in ONE launch per batch: no host-side index draw, no worker processes, no host-to-device copy of rays.  HBM-bound by construction:
// 8 or 0 B of indices + 48 B pose (L2-resident: a few hundred poses) + 12 B pixel gather in, 36 B out per ray.
//
// Random draws are counter-based (splitmix64 of seed, call counter and ray index); the reference uses numpy's global generator, so
// the streams differ -- the distribution is the reference's (independent uniform image and pixel per ray, with replacement).
#include "common.cuh"
#include "../../include/mfnerf_b200.h"

namespace mfn {

__device__ __forceinline__ uint64_t rg_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

struct RayBatchArgs {
    float fx, fy, cx, cy;
    int32_t width, n_pix, n_images, channels;
    const float* directions;      // (n_pix, 3) or nullptr: computed from the intrinsics
    const float* poses;           // (n_images, 3, 4) camera-to-world
    const float* pixels;          // (n_images, n_pix, channels) or nullptr
    const int64_t* img_idxs;      // (n) or nullptr
    const int64_t* pix_idxs;      // (n) or nullptr
    int32_t draw, image;          // draw 0: given indices (missing img_idxs -> `image`, missing pix_idxs -> ray i is pixel i)
    uint64_t seed;
    const unsigned long long* call_counter;
    int64_t n;
    float* rays_o; float* rays_d; float* rgb;
    int64_t* img_out; int64_t* pix_out;
};

__global__ void __launch_bounds__(256) ray_batch_kernel(const __grid_constant__ RayBatchArgs a) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const uint64_t call = a.call_counter ? (uint64_t)*a.call_counter : 0ull;
    const uint64_t key = rg_splitmix64(a.seed ^ (call * 0xD6E8FEB86659FD93ull));
    // draw == 2 (ray_sampling_strategy 'same_image', base.py:26-27): one image for the whole batch
    const int64_t shared_img = (int64_t)uniform_below(rg_splitmix64(key ^ 0x5851F42D4C957F2Dull), (uint32_t)a.n_images);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
        int64_t img, pix;
        if (a.draw) {
            const uint64_t r0 = rg_splitmix64(key ^ ((uint64_t)i * 0xD1342543DE82EF95ull)), r1 = rg_splitmix64(r0);
            img = a.draw == 2 ? shared_img : (int64_t)uniform_below(r0, (uint32_t)a.n_images);      // uniform in [0, n_images)
            pix = (int64_t)uniform_below(r1, (uint32_t)a.n_pix);                                        // uniform in [0, n_pix)
        } else {
            img = a.img_idxs ? a.img_idxs[i] : (int64_t)a.image;
            pix = a.pix_idxs ? a.pix_idxs[i] : i;
        }
        if (a.img_out) a.img_out[i] = img;
        if (a.pix_out) a.pix_out[i] = pix;
        if (img < 0 || img >= a.n_images || pix < 0 || pix >= a.n_pix) {      // (the reference's indexing would raise) -> a ray that hits nothing
            for (int k = 0; k < 3; ++k) { a.rays_o[3 * i + k] = 0.f; a.rays_d[3 * i + k] = 0.f; if (a.rgb) a.rgb[3 * i + k] = 0.f; }
            continue;
        }
        float dx, dy, dz;
        if (a.directions) {
            dx = __ldg(a.directions + 3 * pix); dy = __ldg(a.directions + 3 * pix + 1); dz = __ldg(a.directions + 3 * pix + 2);
        } else {
            const float u = (float)(int)(pix % a.width), v = (float)(int)(pix / a.width);          // create_meshgrid(H, W, False): x = column
            dx = __fdiv_rn(__fadd_rn(__fsub_rn(u, a.cx), 0.5f), a.fx);                             // ray_utils.py:34-35, left to right
            dy = __fdiv_rn(__fadd_rn(__fsub_rn(v, a.cy), 0.5f), a.fy);
            dz = 1.0f;
        }
        const float4* P = reinterpret_cast<const float4*>(a.poses + 12 * img);                     // three rows [R_j0 R_j1 R_j2 t_j]
        const float4 p0 = __ldg(P), p1 = __ldg(P + 1), p2 = __ldg(P + 2);
        // rays_d[j] = sum_k directions[k] * c2w[j][k]  (ray_utils.py:60-64), accumulated k = 0, 1, 2
        a.rays_d[3 * i] = fmaf(dz, p0.z, fmaf(dy, p0.y, dx * p0.x));
        a.rays_d[3 * i + 1] = fmaf(dz, p1.z, fmaf(dy, p1.y, dx * p1.x));
        a.rays_d[3 * i + 2] = fmaf(dz, p2.z, fmaf(dy, p2.y, dx * p2.x));
        a.rays_o[3 * i] = p0.w; a.rays_o[3 * i + 1] = p1.w; a.rays_o[3 * i + 2] = p2.w;            // ray_utils.py:66
        if (a.rgb) {
            const float* px = a.pixels + ((size_t)img * a.n_pix + (size_t)pix) * a.channels;       // self.rays[img_idxs, pix_idxs][:, :3] (base.py:30-32)
            a.rgb[3 * i] = __ldg(px); a.rgb[3 * i + 1] = __ldg(px + 1); a.rgb[3 * i + 2] = __ldg(px + 2);
        }
    }
}

}  // namespace mfn

using namespace mfn;

extern "C" int mfn_ray_batch(const mfn_camera* cam, const float* directions, const float* poses, int32_t n_images, const float* pixels,
                             int32_t pixel_channels, const int64_t* img_idxs, const int64_t* pix_idxs, int32_t image, int32_t draw, uint64_t seed,
                             const void* call_counter, int64_t n_rays, float* rays_o, float* rays_d, float* rgb, int64_t* img_idxs_out,
                             int64_t* pix_idxs_out, void* stream) {
    if (!cam || cam->width < 1 || cam->height < 1 || n_images < 1 || n_rays < 0 || draw < 0 || draw > 2 || !(cam->fx != 0.f) || !(cam->fy != 0.f)) {
        set_error("mfn_ray_batch: bad argument"); return MFN_ERR_ARG;
    }
    if ((int64_t)cam->width * cam->height > 0x7fffffff) { set_error("mfn_ray_batch: more than 2^31 pixels per image"); return MFN_ERR_ARG; }
    if (n_rays == 0) return MFN_OK;
    if (!poses || !rays_o || !rays_d) { set_error("mfn_ray_batch: null pointer"); return MFN_ERR_ARG; }
    if (rgb && (!pixels || pixel_channels < 3)) { set_error("mfn_ray_batch: rgb output needs the pixel array with >= 3 channels"); return MFN_ERR_ARG; }
    if (!draw && !img_idxs && (image < 0 || image >= n_images)) { set_error("mfn_ray_batch: image index out of range"); return MFN_ERR_ARG; }
    if (!draw && !pix_idxs && n_rays > (int64_t)cam->width * cam->height) { set_error("mfn_ray_batch: more rays than pixels"); return MFN_ERR_ARG; }
    if (((uintptr_t)poses & 15u) != 0) { set_error("mfn_ray_batch: poses must be 16-byte aligned"); return MFN_ERR_ARG; }
    RayBatchArgs a{};
    a.fx = cam->fx; a.fy = cam->fy; a.cx = cam->cx; a.cy = cam->cy; a.width = cam->width; a.n_pix = cam->width * cam->height;
    a.n_images = n_images; a.channels = pixel_channels; a.directions = directions; a.poses = poses; a.pixels = pixels;
    a.img_idxs = img_idxs; a.pix_idxs = pix_idxs; a.draw = draw; a.image = image; a.seed = seed;
    a.call_counter = (const unsigned long long*)call_counter; a.n = n_rays; a.rays_o = rays_o; a.rays_d = rays_d; a.rgb = rgb;
    a.img_out = img_idxs_out; a.pix_out = pix_idxs_out;
    int64_t blocks = ceil_div(n_rays, 256);
    const int64_t cap = (int64_t)kNumSMs * 8;
    ray_batch_kernel<<<(unsigned)(blocks > cap ? cap : blocks), 256, 0, (cudaStream_t)stream>>>(a);
    return check_launch("mfn_ray_batch", (cudaStream_t)stream);
}
