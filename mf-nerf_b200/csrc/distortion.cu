// Mip-NeRF-360 distortion loss in the DVGO-v2 prefix-sum form  (ref: models/csrc/losses.cu:9-175)
//
// One warp per rays_a row, samples strided over lanes, warp-shuffle scans with a running carry.  The
// reference materialises wts, four scans and _loss as N-sized temporaries with ~10 ATen launches; here
// the forward is one launch (20 B/sample + 28 B/ray) and the backward one launch (24 B/sample + 28 B/ray).
#include "common.cuh"
#include "../../include/mfnerf_b200.h"

namespace mfn {

constexpr int kDistWarps = 8;

__global__ void __launch_bounds__(kDistWarps * 32)
distortion_fw_kernel(const float* __restrict__ ws, const float* __restrict__ deltas, const float* __restrict__ ts,
                     const int64_t* __restrict__ rays_a, int64_t n_rows, float* __restrict__ loss,
                     float* __restrict__ ws_incl, float* __restrict__ wts_incl) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * kDistWarps + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const int64_t ray = rays_a[3 * row], start = rays_a[3 * row + 1];
    const int n = (int)rays_a[3 * row + 2];
    float w_run = 0.f, wt_run = 0.f, acc = 0.f;
    for (int base = 0; base < n; base += 32) {
        const int s = base + lane;
        const bool valid = s < n;
        const int64_t g = start + s;
        float w = 0.f, wt = 0.f, dl = 0.f;
        if (valid) { w = ws[g]; wt = __fmul_rn(w, ts[g]); dl = deltas[g]; }
        const float wi = w_run + warp_incl_scan(w, lane);     // inclusive scans (losses.cu:26-33)
        const float wti = wt_run + warp_incl_scan(wt, lane);
        float we = __shfl_up_sync(0xffffffffu, wi, 1);        // exclusive = previous inclusive (l.35-42)
        float wte = __shfl_up_sync(0xffffffffu, wti, 1);
        if (lane == 0) { we = w_run; wte = wt_run; }
        if (valid) {
            ws_incl[g] = wi; wts_incl[g] = wti;
            // _loss = 2*(wts_incl*ws_excl - ws_incl*wts_excl) + 1/3*ws*ws*deltas   (l.94-95)
            acc += 2.0f * (wti * we - wi * wte) + 0.33333334f * w * w * dl;
        }
        w_run = __shfl_sync(0xffffffffu, wi, 31);
        wt_run = __shfl_sync(0xffffffffu, wti, 31);
    }
    acc = warp_sum(acc);
    if (lane == 0) loss[ray] = acc;
}

__global__ void __launch_bounds__(kDistWarps * 32)
distortion_bw_kernel(const float* __restrict__ dL_dloss, const float* __restrict__ ws_incl, const float* __restrict__ wts_incl,
                     const float* __restrict__ ws, const float* __restrict__ deltas, const float* __restrict__ ts,
                     const int64_t* __restrict__ rays_a, int64_t n_rows, float* __restrict__ dL_dws) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * kDistWarps + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const int64_t ray = rays_a[3 * row], start = rays_a[3 * row + 1];
    const int n = (int)rays_a[3 * row + 2];
    if (n == 0) return;  // the reference reads index start-1 here and discards it (losses.cu:127-130)
    const float gl = dL_dloss[ray];
    const float w_sum = ws_incl[start + n - 1], wt_sum = wts_incl[start + n - 1];
    for (int s = lane; s < n; s += 32) {
        const int64_t g = start + s;
        const float t = ts[g];
        const float before = (s == 0) ? 0.f : (t * ws_incl[g - 1] - wts_incl[g - 1]);
        const float after = wt_sum - wts_incl[g] - t * (w_sum - ws_incl[g]);
        float v = gl * 2.0f * (before + after);                 // l.132-139
        v += (gl * 2.0f) / 3.0f * ws[g] * deltas[g];            // l.140 parses as ((g*2)/3)*w*delta
        dL_dws[g] = v;
    }
}

}  // namespace mfn

using namespace mfn;

extern "C" int mfn_distortion_loss_fw(const float* ws, const float* deltas, const float* ts, const int64_t* rays_a, int64_t n_rays,
                                      int64_t n_samples, float* loss, float* ws_inclusive_scan, float* wts_inclusive_scan, void* stream) {
    (void)n_samples;
    if (n_rays < 0) { set_error("mfn_distortion_loss_fw: bad argument"); return MFN_ERR_ARG; }
    if (n_rays == 0) return MFN_OK;
    if (!rays_a || !loss) { set_error("mfn_distortion_loss_fw: null pointer"); return MFN_ERR_ARG; }
    ProfScope ps("distortion_fw", (cudaStream_t)stream);
    distortion_fw_kernel<<<(int)ceil_div(n_rays, kDistWarps), kDistWarps * 32, 0, (cudaStream_t)stream>>>(
        ws, deltas, ts, rays_a, n_rays, loss, ws_inclusive_scan, wts_inclusive_scan);
    return check_launch("mfn_distortion_loss_fw", (cudaStream_t)stream);
}

extern "C" int mfn_distortion_loss_bw(const float* dL_dloss, const float* ws_inclusive_scan, const float* wts_inclusive_scan,
                                      const float* ws, const float* deltas, const float* ts, const int64_t* rays_a, int64_t n_rays,
                                      int64_t n_samples, float* dL_dws, void* stream) {
    (void)n_samples;
    if (n_rays < 0) { set_error("mfn_distortion_loss_bw: bad argument"); return MFN_ERR_ARG; }
    if (n_rays == 0) return MFN_OK;
    if (!rays_a || !dL_dloss) { set_error("mfn_distortion_loss_bw: null pointer"); return MFN_ERR_ARG; }
    ProfScope ps("distortion_bw", (cudaStream_t)stream);
    distortion_bw_kernel<<<(int)ceil_div(n_rays, kDistWarps), kDistWarps * 32, 0, (cudaStream_t)stream>>>(
        dL_dloss, ws_inclusive_scan, wts_inclusive_scan, ws, deltas, ts, rays_a, n_rays, dL_dws);
    return check_launch("mfn_distortion_loss_bw", (cudaStream_t)stream);
}
