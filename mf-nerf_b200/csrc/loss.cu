// Per-ray NeRF loss and its gradient w.r.t. the compositor outputs, in one launch.
// Restates losses.py:47-60 (NeRFLoss: squared rgb error + opacity entropy [+ lambda * distortion]) and train.py:178
// (loss = sum of the per-term means), together with the background blend of rendering.py:153-161
// (rgb_final = rgb + bg * (1 - opacity)).  Replaces ~15 elementwise ATen launches and their autograd graph.
#include "common.cuh"
#include "../../include/mfnerf_b200.h"

namespace mfn {

// loss_out[0] += sum of the rgb term, [1] += opacity term, [2] += distortion term (already divided by their counts)
__global__ void __launch_bounds__(256)
nerf_loss_kernel(const float* __restrict__ rgb, const float* __restrict__ opacity, const float* __restrict__ target, const float* __restrict__ distortion,
                 int64_t n_rays, float bg_r, float bg_g, float bg_b, float lambda_opacity, float lambda_distortion, float out_scale,
                 float* __restrict__ dL_drgb, float* __restrict__ dL_dopacity, float* __restrict__ dL_ddistortion, float* __restrict__ rgb_final,
                 float* __restrict__ loss_out) {
    const float inv_r = 1.0f / (float)n_rays, inv_3r = 1.0f / (3.0f * (float)n_rays);
    float l_rgb = 0.f, l_op = 0.f, l_dist = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rays; i += (int64_t)gridDim.x * blockDim.x) {
        const float o = opacity[i];
        const float t = 1.0f - o;
        const float fr = rgb[3 * i] + bg_r * t, fg = rgb[3 * i + 1] + bg_g * t, fb = rgb[3 * i + 2] + bg_b * t;
        const float er = fr - target[3 * i], eg = fg - target[3 * i + 1], eb = fb - target[3 * i + 2];
        l_rgb += (er * er + eg * eg + eb * eb) * inv_3r;
        const float gr = 2.0f * er * inv_3r, gg = 2.0f * eg * inv_3r, gb = 2.0f * eb * inv_3r;
        const float oe = o + 1e-10f;                         // losses.py:51
        const float lg = logf(oe);
        l_op += lambda_opacity * (-oe * lg) * inv_r;         // losses.py:53
        float g_o = -(gr * bg_r + gg * bg_g + gb * bg_b) + lambda_opacity * (-(lg + 1.0f)) * inv_r;
        if (rgb_final) { rgb_final[3 * i] = fr; rgb_final[3 * i + 1] = fg; rgb_final[3 * i + 2] = fb; }
        dL_drgb[3 * i] = gr * out_scale; dL_drgb[3 * i + 1] = gg * out_scale; dL_drgb[3 * i + 2] = gb * out_scale;
        dL_dopacity[i] = g_o * out_scale;
        if (distortion) {
            l_dist += lambda_distortion * distortion[i] * inv_r;
            dL_ddistortion[i] = lambda_distortion * inv_r * out_scale;
        }
    }
    l_rgb = warp_sum(l_rgb); l_op = warp_sum(l_op); l_dist = warp_sum(l_dist);
    if ((threadIdx.x & 31) == 0 && loss_out) {
        atomicAdd(loss_out + 0, l_rgb); atomicAdd(loss_out + 1, l_op);
        if (distortion) atomicAdd(loss_out + 2, l_dist);
    }
}

// rendering.py:29 -- hits_t[(t1 >= 0) & (t1 < NEAR), 0] = NEAR
__global__ void clamp_near_kernel(float* __restrict__ hits_t, int64_t n, float near_t) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float t1 = hits_t[2 * i];
    if (t1 >= 0.f && t1 < near_t) hits_t[2 * i] = near_t;
}

}  // namespace mfn

using namespace mfn;

extern "C" int mfn_nerf_loss_fwbw(const float* rgb, const float* opacity, const float* target, const float* distortion, int64_t n_rays,
                                  const float* bg_rgb_host, float lambda_opacity, float lambda_distortion, float grad_scale, float* dL_drgb,
                                  float* dL_dopacity, float* dL_ddistortion, float* rgb_final, float* loss_out, void* stream) {
    if (n_rays < 0) { set_error("mfn_nerf_loss_fwbw: bad n_rays"); return MFN_ERR_ARG; }
    if (n_rays == 0) return MFN_OK;
    if (!rgb || !opacity || !target || !dL_drgb || !dL_dopacity || !bg_rgb_host || (distortion && !dL_ddistortion)) {
        set_error("mfn_nerf_loss_fwbw: null pointer"); return MFN_ERR_ARG;
    }
    int64_t blocks = ceil_div(n_rays, 256);
    if (blocks > kNumSMs * 4) blocks = kNumSMs * 4;
    ProfScope ps("nerf_loss", (cudaStream_t)stream);
    nerf_loss_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(rgb, opacity, target, distortion, n_rays, bg_rgb_host[0], bg_rgb_host[1],
                                                                       bg_rgb_host[2], lambda_opacity, lambda_distortion, grad_scale, dL_drgb,
                                                                       dL_dopacity, dL_ddistortion, rgb_final, loss_out);
    return check_launch("mfn_nerf_loss_fwbw", (cudaStream_t)stream);
}

extern "C" int mfn_clamp_near(float* hits_t, int64_t n_rays, float near_distance, void* stream) {
    if (n_rays < 0 || (n_rays > 0 && !hits_t)) { set_error("mfn_clamp_near: bad argument"); return MFN_ERR_ARG; }
    if (n_rays == 0) return MFN_OK;
    clamp_near_kernel<<<(unsigned)ceil_div(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(hits_t, n_rays, near_distance);
    return check_launch("mfn_clamp_near", (cudaStream_t)stream);
}

// ---- segmented sum (replaces torch_scatter.segment_csr(src, indptr) as used by RayMarcher.backward,
// custom_functions.py:102-112: per-ray sums of dL/dxyz and dL/dxyz * t + dL/ddir).  One warp per segment, rows of `width`
// floats; segment s covers rows [indptr[s], indptr[s+1]).
namespace mfn {
__global__ void __launch_bounds__(256)
segment_sum_kernel(const float* __restrict__ src, const int64_t* __restrict__ indptr, int64_t n_seg, int width, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t s = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (s >= n_seg) return;
    const int64_t b = indptr[s], e = indptr[s + 1];
    for (int c = 0; c < width; ++c) {
        float acc = 0.f;
        for (int64_t i = b + lane; i < e; i += 32) acc += src[i * width + c];
        acc = warp_sum(acc);
        if (lane == 0) out[s * width + c] = acc;
    }
}
}  // namespace mfn

extern "C" int mfn_segment_sum(const float* src, const int64_t* indptr, int64_t n_segments, int width, float* out, void* stream) {
    if (n_segments < 0 || width < 1 || (n_segments > 0 && (!src || !indptr || !out))) { set_error("mfn_segment_sum: bad argument"); return MFN_ERR_ARG; }
    if (n_segments == 0) return MFN_OK;
    mfn::segment_sum_kernel<<<(unsigned)ceil_div(n_segments, 8), 256, 0, (cudaStream_t)stream>>>(src, indptr, n_segments, width, out);
    return check_launch("mfn_segment_sum", (cudaStream_t)stream);
}
