// Transmittance compositing, forward / backward / test-time  (ref: models/csrc/volumerendering.cu)
//
// Train kernels: one warp per rays_a row; lanes stride over the ray's samples so every global access is
// coalesced (the reference walks each ray with a single thread).  The transmittance recurrence
// T_{s+1} = T_s * (1 - a_s) is replayed in the reference's exact sequential order (each lane runs the same
// 32-step chain from shuffled alphas) so ws = a*T and the early-termination index are bit-identical given
// identical inputs; the per-ray sums (rgb, depth, opacity) use a warp tree and therefore agree with the
// reference's sequential fp32 accumulation only to rounding (tolerance stated in tests/).
// Algorithmic bytes: fw 28 B/sample + 52 B/ray, bw 48 B/sample + 64 B/ray.
#include "common.cuh"
#include <stdlib.h>
#include "../../include/mfnerf_b200.h"

namespace mfn {

constexpr int kCompWarps = 8;
constexpr int kCompLossWarps = 4;      // default CTA size of composite_loss_train_kernel, in rays

// alpha exactly as the reference computes it: 1 - __expf(-sigma*delta)  (volumerendering.cu:30)
__device__ __forceinline__ float alpha_of(float sigma, float delta) { return __fadd_rn(1.0f, -__expf(-__fmul_rn(sigma, delta))); }

// Sequential transmittance over one 32-sample chunk.  Returns T before this lane's sample; *t_end is T after the whole chunk
// (identical in all lanes).  The 32 factors (1 - a_j) travel through shared memory: every lane reads them back as 8 independent
// 16-byte loads and then runs the 32-step product without waiting on a shuffle per step (the shuffle version spent a third of the
// kernel on short-scoreboard stalls; the kernel's duration is the longest ray's chain).
__device__ __forceinline__ float chunk_transmittance(float a, float T_in, int lane, float* s_om, float* t_end) {
    s_om[lane] = __fadd_rn(1.0f, -a);      // (1 - a_j) is rounded once per sample, exactly as in the sequential loop
    __syncwarp();
    float4 v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = reinterpret_cast<const float4*>(s_om)[q];
    __syncwarp();                          // the next chunk may overwrite s_om
    float Tj = T_in, mine = T_in;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const float o[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (4 * q + k == lane) mine = Tj;
            Tj = __fmul_rn(Tj, o[k]);
        }
    }
    *t_end = Tj;
    return mine;
}

__global__ void __launch_bounds__(kCompWarps * 32)
composite_train_fw_kernel(const float* __restrict__ sigmas, const float* __restrict__ rgbs, const float* __restrict__ deltas,
                          const float* __restrict__ ts, const int64_t* __restrict__ rays_a, float T_thr, int64_t n_rows,
                          int64_t* __restrict__ total_samples, float* __restrict__ opacity, float* __restrict__ depth,
                          float* __restrict__ rgb, float* __restrict__ ws) {
    __shared__ __align__(16) float s_om[kCompWarps][32];
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * kCompWarps + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const int64_t ray = rays_a[3 * row], start = rays_a[3 * row + 1];
    const int n = (int)rays_a[3 * row + 2];
    float T = 1.0f, acc_r = 0.f, acc_g = 0.f, acc_b = 0.f, acc_d = 0.f, acc_o = 0.f;
    int samples = n;
    int base = 0;
    // the chunk being composited was loaded one iteration earlier: the loads of chunk k+1 are in flight while the 32-step
    // transmittance chain of chunk k runs (the kernel's duration is the longest ray's chain, so latency per chunk is what counts)
    float n_sig = 0.f, n_dl = 0.f, n_t = 0.f, n_r = 0.f, n_g = 0.f, n_b = 0.f;
    if (lane < n) { const int64_t g = start + lane; n_sig = sigmas[g]; n_dl = deltas[g]; n_t = ts[g]; n_r = rgbs[3 * g]; n_g = rgbs[3 * g + 1]; n_b = rgbs[3 * g + 2]; }
    for (; base < n; base += 32) {
        const int s = base + lane;
        const bool valid = s < n;
        const int64_t g = start + s;
        const float sig = n_sig, dl = n_dl, t = n_t, cr = n_r, cg = n_g, cb = n_b;
        if (s + 32 < n) { const int64_t g2 = g + 32; n_sig = sigmas[g2]; n_dl = deltas[g2]; n_t = ts[g2]; n_r = rgbs[3 * g2]; n_g = rgbs[3 * g2 + 1]; n_b = rgbs[3 * g2 + 2]; }
        else { n_sig = 0.f; n_dl = 0.f; n_t = 0.f; n_r = 0.f; n_g = 0.f; n_b = 0.f; }
        const float a = valid ? alpha_of(sig, dl) : 0.f;
        float T_end;
        const float T_mine = chunk_transmittance(a, T, lane, s_om[threadIdx.x >> 5], &T_end);
        const float T_after = __fmul_rn(T_mine, __fadd_rn(1.0f, -a));
        const uint32_t stop = __ballot_sync(0xffffffffu, valid && T_after <= T_thr);  // l.41
        const int last = stop ? (__ffs(stop) - 1) : 31;
        const bool in = valid && lane <= last;
        const float w = in ? __fmul_rn(a, T_mine) : 0.f;
        if (valid) ws[g] = w;
        acc_r += w * cr; acc_g += w * cg; acc_b += w * cb; acc_d += w * t; acc_o += w;
        T = T_end;
        if (stop) { samples = base + last; base += 32; break; }
    }
    for (int s = base + lane; s < n; s += 32) ws[start + s] = 0.f;  // samples behind an early stop keep w = 0
    acc_r = warp_sum(acc_r); acc_g = warp_sum(acc_g); acc_b = warp_sum(acc_b); acc_d = warp_sum(acc_d); acc_o = warp_sum(acc_o);
    if (lane == 0) {
        total_samples[ray] = samples;
        opacity[ray] = acc_o; depth[ray] = acc_d;
        rgb[3 * ray] = acc_r; rgb[3 * ray + 1] = acc_g; rgb[3 * ray + 2] = acc_b;
    }
}

__global__ void __launch_bounds__(kCompWarps * 32)
composite_train_bw_kernel(const float* __restrict__ dL_dopacity, const float* __restrict__ dL_ddepth, const float* __restrict__ dL_drgb,
                          const float* __restrict__ dL_dws, const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                          const float* __restrict__ ws, const float* __restrict__ deltas, const float* __restrict__ ts,
                          const int64_t* __restrict__ rays_a, const float* __restrict__ opacity, const float* __restrict__ depth,
                          const float* __restrict__ rgb, float T_thr, int64_t n_rows, float* __restrict__ dL_dsigmas,
                          float* __restrict__ dL_drgbs) {
    __shared__ __align__(16) float s_om[kCompWarps][32];
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * kCompWarps + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const int64_t ray = rays_a[3 * row], start = rays_a[3 * row + 1];
    const int n = (int)rays_a[3 * row + 2];
    if (n == 0) return;
    // sum over the ray of dL/dw * w  (the reference builds it with an in-place inclusive scan, l.119-123)
    float wsum = 0.f;
    for (int s = lane; s < n; s += 32) wsum += dL_dws[start + s] * ws[start + s];
    wsum = warp_sum(wsum);
    const float R = rgb[3 * ray], G = rgb[3 * ray + 1], B = rgb[3 * ray + 2], O = opacity[ray], D = depth[ray];
    const float gR = dL_drgb[3 * ray], gG = dL_drgb[3 * ray + 1], gB = dL_drgb[3 * ray + 2];
    const float gO = dL_dopacity[ray], gD = dL_ddepth[ray];
    float T = 1.0f, cr_run = 0.f, cg_run = 0.f, cb_run = 0.f, d_run = 0.f, p_run = 0.f;
    int base = 0;
    float n_sig = 0.f, n_dl = 0.f, n_t = 0.f, n_r = 0.f, n_g = 0.f, n_b = 0.f, n_gw = 0.f, n_ws = 0.f;   // next chunk, prefetched (see the forward kernel)
    if (lane < n) {
        const int64_t g = start + lane;
        n_sig = sigmas[g]; n_dl = deltas[g]; n_t = ts[g]; n_r = rgbs[3 * g]; n_g = rgbs[3 * g + 1]; n_b = rgbs[3 * g + 2]; n_gw = dL_dws[g]; n_ws = ws[g];
    }
    for (; base < n; base += 32) {
        const int s = base + lane;
        const bool valid = s < n;
        const int64_t g = start + s;
        const float sig = n_sig, dl = n_dl, t = n_t, cr = n_r, cg = n_g, cb = n_b, gw = n_gw, wsv = n_ws;
        if (s + 32 < n) {
            const int64_t g2 = g + 32;
            n_sig = sigmas[g2]; n_dl = deltas[g2]; n_t = ts[g2]; n_r = rgbs[3 * g2]; n_g = rgbs[3 * g2 + 1]; n_b = rgbs[3 * g2 + 2]; n_gw = dL_dws[g2]; n_ws = ws[g2];
        } else { n_sig = 0.f; n_dl = 0.f; n_t = 0.f; n_r = 0.f; n_g = 0.f; n_b = 0.f; n_gw = 0.f; n_ws = 0.f; }
        const float a = valid ? alpha_of(sig, dl) : 0.f;
        float T_end;
        const float T_mine = chunk_transmittance(a, T, lane, s_om[threadIdx.x >> 5], &T_end);
        const float T_after = __fmul_rn(T_mine, __fadd_rn(1.0f, -a));
        const uint32_t stop = __ballot_sync(0xffffffffu, valid && T_after <= T_thr);  // l.148
        const int last = stop ? (__ffs(stop) - 1) : 31;
        const bool in = valid && lane <= last;
        const float w = valid ? __fmul_rn(a, T_mine) : 0.f;
        // inclusive running sums along the ray (r, g, b, d of l.131-132 and the dL_dws*ws prefix of l.119)
        const float r_in = cr_run + warp_incl_scan(w * cr, lane);
        const float g_in = cg_run + warp_incl_scan(w * cg, lane);
        const float b_in = cb_run + warp_incl_scan(w * cb, lane);
        const float d_in = d_run + warp_incl_scan(w * t, lane);
        const float p_in = p_run + warp_incl_scan(gw * wsv, lane);
        if (valid) {
            float ds = 0.f, dr = 0.f, dg = 0.f, db = 0.f;
            if (in) {
                dr = gR * w; dg = gG * w; db = gB * w;
                ds = dl * (gR * (cr * T_after - (R - r_in)) + gG * (cg * T_after - (G - g_in)) + gB * (cb * T_after - (B - b_in)) +
                           gO * (1.0f - O) + gD * (t * T_after - (D - d_in)) + T_after * gw - (wsum - p_in));
            }
            dL_dsigmas[g] = ds;
            dL_drgbs[3 * g] = dr; dL_drgbs[3 * g + 1] = dg; dL_drgbs[3 * g + 2] = db;
        }
        cr_run = __shfl_sync(0xffffffffu, r_in, 31); cg_run = __shfl_sync(0xffffffffu, g_in, 31);
        cb_run = __shfl_sync(0xffffffffu, b_in, 31); d_run = __shfl_sync(0xffffffffu, d_in, 31);
        p_run = __shfl_sync(0xffffffffu, p_in, 31);
        T = T_end;
        if (stop) { base += 32; break; }
    }
    for (int s = base + lane; s < n; s += 32) {
        const int64_t g = start + s;
        dL_dsigmas[g] = 0.f; dL_drgbs[3 * g] = 0.f; dL_drgbs[3 * g + 1] = 0.f; dL_drgbs[3 * g + 2] = 0.f;
    }
}

// Training step without distortion loss: compositing forward, the per-ray NeRF loss with its gradient (loss.cu, losses.py:47-60 +
// rendering.py:153-161) and compositing backward in ONE launch.  A ray's loss gradient depends on that ray's outputs only, so the warp
// that composited the ray goes straight on to its backward pass (second sweep over the samples, now L1/L2-hot) -- two launches, the
// loss kernel and the re-read of the compositor outputs disappear from the step's critical path.  Same device functions and the same
// order of operations as the three separate kernels; dL/ddepth and dL/dws are zero for this loss and their terms are dropped.
// The three loss sums are WRITTEN (not accumulated) by the last block to finish (ticket in `scratch`), which also clears *clear_flag.
__global__ void __launch_bounds__(kCompWarps * 32)
composite_loss_train_kernel(const float* __restrict__ sigmas, const float* __restrict__ rgbs, const float* __restrict__ deltas,
                            const float* __restrict__ ts, const int64_t* __restrict__ rays_a, const float* __restrict__ target, float T_thr,
                            int64_t n_rows, float bg_r, float bg_g, float bg_b, float lambda_opacity, float out_scale,
                            int64_t* __restrict__ total_samples, float* __restrict__ opacity, float* __restrict__ depth, float* __restrict__ rgb,
                            float* __restrict__ ws, float* __restrict__ rgb_final, float* __restrict__ dL_drgb, float* __restrict__ dL_dopacity,
                            float* __restrict__ dL_dsigmas, float* __restrict__ dL_drgbs, float* __restrict__ loss_out,
                            float* __restrict__ scratch, int32_t* __restrict__ clear_flag) {
    __shared__ __align__(16) float s_om[kCompWarps][32];
    __shared__ float s_loss[kCompWarps][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_warps = (int)(blockDim.x >> 5);      // rays per CTA: a CTA keeps its SM slots until its LONGEST ray is done (barrier below)
    const int64_t row = (int64_t)blockIdx.x * n_warps + warp;
    float l_rgb = 0.f, l_op = 0.f;
    if (row < n_rows) {
        const int64_t ray = rays_a[3 * row], start = rays_a[3 * row + 1];
        const int n = (int)rays_a[3 * row + 2];
        // ---- forward (composite_train_fw_kernel)
        float T = 1.0f, acc_r = 0.f, acc_g = 0.f, acc_b = 0.f, acc_d = 0.f, acc_o = 0.f;
        int samples = n;
        int base = 0;
        float n_sig = 0.f, n_dl = 0.f, n_t = 0.f, n_r = 0.f, n_g = 0.f, n_b = 0.f;
        if (lane < n) { const int64_t g = start + lane; n_sig = sigmas[g]; n_dl = deltas[g]; n_t = ts[g]; n_r = rgbs[3 * g]; n_g = rgbs[3 * g + 1]; n_b = rgbs[3 * g + 2]; }
        for (; base < n; base += 32) {
            const int s = base + lane;
            const bool valid = s < n;
            const int64_t g = start + s;
            const float sig = n_sig, dl = n_dl, t = n_t, cr = n_r, cg = n_g, cb = n_b;
            if (s + 32 < n) { const int64_t g2 = g + 32; n_sig = sigmas[g2]; n_dl = deltas[g2]; n_t = ts[g2]; n_r = rgbs[3 * g2]; n_g = rgbs[3 * g2 + 1]; n_b = rgbs[3 * g2 + 2]; }
            else { n_sig = 0.f; n_dl = 0.f; n_t = 0.f; n_r = 0.f; n_g = 0.f; n_b = 0.f; }
            const float a = valid ? alpha_of(sig, dl) : 0.f;
            float T_end;
            const float T_mine = chunk_transmittance(a, T, lane, s_om[warp], &T_end);
            const float T_after = __fmul_rn(T_mine, __fadd_rn(1.0f, -a));
            const uint32_t stop = __ballot_sync(0xffffffffu, valid && T_after <= T_thr);
            const int last = stop ? (__ffs(stop) - 1) : 31;
            const bool in = valid && lane <= last;
            const float w = in ? __fmul_rn(a, T_mine) : 0.f;
            if (valid) ws[g] = w;
            acc_r += w * cr; acc_g += w * cg; acc_b += w * cb; acc_d += w * t; acc_o += w;
            T = T_end;
            if (stop) { samples = base + last; base += 32; break; }
        }
        for (int s = base + lane; s < n; s += 32) ws[start + s] = 0.f;
        // (xor butterfly: every lane ends up with the same bits)
        const float R = warp_sum(acc_r), G = warp_sum(acc_g), B = warp_sum(acc_b), D = warp_sum(acc_d), O = warp_sum(acc_o);
        // ---- loss and its gradient w.r.t. this ray's outputs (nerf_loss_kernel)
        const float inv_r = 1.0f / (float)n_rows, inv_3r = 1.0f / (3.0f * (float)n_rows);
        const float tt = 1.0f - O;
        const float fr = R + bg_r * tt, fg = G + bg_g * tt, fb = B + bg_b * tt;
        const float er = fr - target[3 * ray], eg = fg - target[3 * ray + 1], eb = fb - target[3 * ray + 2];
        l_rgb = (er * er + eg * eg + eb * eb) * inv_3r;
        const float gr = 2.0f * er * inv_3r, gg = 2.0f * eg * inv_3r, gb = 2.0f * eb * inv_3r;
        const float oe = O + 1e-10f;
        const float lg = logf(oe);
        l_op = lambda_opacity * (-oe * lg) * inv_r;
        const float gR = gr * out_scale, gG = gg * out_scale, gB = gb * out_scale;
        const float gO = (-(gr * bg_r + gg * bg_g + gb * bg_b) + lambda_opacity * (-(lg + 1.0f)) * inv_r) * out_scale;
        if (lane == 0) {
            total_samples[ray] = samples;
            opacity[ray] = O; depth[ray] = D;
            rgb[3 * ray] = R; rgb[3 * ray + 1] = G; rgb[3 * ray + 2] = B;
            if (rgb_final) { rgb_final[3 * ray] = fr; rgb_final[3 * ray + 1] = fg; rgb_final[3 * ray + 2] = fb; }
            if (dL_drgb) { dL_drgb[3 * ray] = gR; dL_drgb[3 * ray + 1] = gG; dL_drgb[3 * ray + 2] = gB; }
            if (dL_dopacity) dL_dopacity[ray] = gO;
        }
        // ---- backward (composite_train_bw_kernel with dL/ddepth = 0, dL/dws = 0)
        T = 1.0f;
        float cr_run = 0.f, cg_run = 0.f, cb_run = 0.f;
        base = 0;
        n_sig = 0.f; n_dl = 0.f; n_r = 0.f; n_g = 0.f; n_b = 0.f;
        if (lane < n) { const int64_t g = start + lane; n_sig = sigmas[g]; n_dl = deltas[g]; n_r = rgbs[3 * g]; n_g = rgbs[3 * g + 1]; n_b = rgbs[3 * g + 2]; }
        for (; base < n; base += 32) {
            const int s = base + lane;
            const bool valid = s < n;
            const int64_t g = start + s;
            const float sig = n_sig, dl = n_dl, cr = n_r, cg = n_g, cb = n_b;
            if (s + 32 < n) { const int64_t g2 = g + 32; n_sig = sigmas[g2]; n_dl = deltas[g2]; n_r = rgbs[3 * g2]; n_g = rgbs[3 * g2 + 1]; n_b = rgbs[3 * g2 + 2]; }
            else { n_sig = 0.f; n_dl = 0.f; n_r = 0.f; n_g = 0.f; n_b = 0.f; }
            const float a = valid ? alpha_of(sig, dl) : 0.f;
            float T_end;
            const float T_mine = chunk_transmittance(a, T, lane, s_om[warp], &T_end);
            const float T_after = __fmul_rn(T_mine, __fadd_rn(1.0f, -a));
            const uint32_t stop = __ballot_sync(0xffffffffu, valid && T_after <= T_thr);
            const int last = stop ? (__ffs(stop) - 1) : 31;
            const bool in = valid && lane <= last;
            const float w = valid ? __fmul_rn(a, T_mine) : 0.f;
            const float r_in = cr_run + warp_incl_scan(w * cr, lane);
            const float g_in = cg_run + warp_incl_scan(w * cg, lane);
            const float b_in = cb_run + warp_incl_scan(w * cb, lane);
            if (valid) {
                float ds = 0.f, dr = 0.f, dg = 0.f, db = 0.f;
                if (in) {
                    dr = gR * w; dg = gG * w; db = gB * w;
                    ds = dl * (gR * (cr * T_after - (R - r_in)) + gG * (cg * T_after - (G - g_in)) + gB * (cb * T_after - (B - b_in)) + gO * (1.0f - O));
                }
                dL_dsigmas[g] = ds;
                dL_drgbs[3 * g] = dr; dL_drgbs[3 * g + 1] = dg; dL_drgbs[3 * g + 2] = db;
            }
            cr_run = __shfl_sync(0xffffffffu, r_in, 31); cg_run = __shfl_sync(0xffffffffu, g_in, 31); cb_run = __shfl_sync(0xffffffffu, b_in, 31);
            T = T_end;
            if (stop) { base += 32; break; }
        }
        for (int s = base + lane; s < n; s += 32) {
            const int64_t g = start + s;
            dL_dsigmas[g] = 0.f; dL_drgbs[3 * g] = 0.f; dL_drgbs[3 * g + 1] = 0.f; dL_drgbs[3 * g + 2] = 0.f;
        }
    }
    // ---- loss totals: block partial -> two atomics per block; the last block publishes the sums and leaves the scratch zeroed
    if (lane == 0) { s_loss[warp][0] = l_rgb; s_loss[warp][1] = l_op; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int k = 0; k < n_warps; ++k) { a0 += s_loss[k][0]; a1 += s_loss[k][1]; }
        atomicAdd(scratch, a0); atomicAdd(scratch + 1, a1);
        __threadfence();
        unsigned int* ticket = reinterpret_cast<unsigned int*>(scratch + 2);
        if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
            __threadfence();
            const float S0 = atomicAdd(scratch, 0.f), S1 = atomicAdd(scratch + 1, 0.f);
            if (loss_out) { loss_out[0] = S0; loss_out[1] = S1; loss_out[2] = 0.f; }
            scratch[0] = 0.f; scratch[1] = 0.f; *ticket = 0u;
            if (clear_flag) *clear_flag = 0;
        }
    }
}

// test-time compositor: one thread per alive ray, the reference's sequential order (so opacity / depth /
// rgb are bit-identical to the reference given identical inputs).  (ref: volumerendering.cu:219-248)
__global__ void composite_test_fw_kernel(const float* __restrict__ sigmas, const float* __restrict__ rgbs, const float* __restrict__ deltas,
                                         const float* __restrict__ ts, int64_t* __restrict__ alive, float T_thr,
                                         const int32_t* __restrict__ n_eff, int n_samples, int64_t n_alive,
                                         float* __restrict__ opacity, float* __restrict__ depth, float* __restrict__ rgb) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_alive) return;
    const int ne = n_eff[n];
    if (ne == 0) { alive[n] = -1; return; }
    const int64_t r = alive[n];
    float o = opacity[r], d = depth[r], cr = rgb[3 * r], cg = rgb[3 * r + 1], cb = rgb[3 * r + 2];
    float T = __fadd_rn(1.0f, -o);
    const int64_t rowb = n * (int64_t)n_samples;
    for (int s = 0; s < ne; ++s) {
        const float a = alpha_of(sigmas[rowb + s], deltas[rowb + s]);
        const float w = __fmul_rn(a, T);
        cr = __fmaf_rn(w, rgbs[3 * (rowb + s)], cr);
        cg = __fmaf_rn(w, rgbs[3 * (rowb + s) + 1], cg);
        cb = __fmaf_rn(w, rgbs[3 * (rowb + s) + 2], cb);
        d = __fmaf_rn(w, ts[rowb + s], d);
        o = __fadd_rn(o, w);
        T = __fmul_rn(T, __fadd_rn(1.0f, -a));
        if (T <= T_thr) { alive[n] = -1; break; }
    }
    opacity[r] = o; depth[r] = d; rgb[3 * r] = cr; rgb[3 * r + 1] = cg; rgb[3 * r + 2] = cb;
}

}  // namespace mfn

using namespace mfn;

extern "C" int mfn_composite_train_fw(const float* sigmas, const float* rgbs, const float* deltas, const float* ts, const int64_t* rays_a,
                                      float T_threshold, int64_t n_rays, int64_t n_samples, int64_t* total_samples, float* opacity,
                                      float* depth, float* rgb, float* ws, void* stream) {
    (void)n_samples;
    if (n_rays < 0) { set_error("mfn_composite_train_fw: bad argument"); return MFN_ERR_ARG; }
    if (n_rays == 0) return MFN_OK;
    if (!rays_a || !total_samples || !opacity || !depth || !rgb) { set_error("mfn_composite_train_fw: null pointer"); return MFN_ERR_ARG; }
    ProfScope ps("composite_train_fw", (cudaStream_t)stream);
    composite_train_fw_kernel<<<(int)ceil_div(n_rays, kCompWarps), kCompWarps * 32, 0, (cudaStream_t)stream>>>(
        sigmas, rgbs, deltas, ts, rays_a, T_threshold, n_rays, total_samples, opacity, depth, rgb, ws);
    return check_launch("mfn_composite_train_fw", (cudaStream_t)stream);
}

extern "C" int mfn_composite_train_bw(const float* dL_dopacity, const float* dL_ddepth, const float* dL_drgb, const float* dL_dws,
                                      const float* sigmas, const float* rgbs, const float* ws, const float* deltas, const float* ts,
                                      const int64_t* rays_a, const float* opacity, const float* depth, const float* rgb,
                                      float T_threshold, int64_t n_rays, int64_t n_samples, float* dL_dsigmas, float* dL_drgbs,
                                      void* stream) {
    (void)n_samples;
    if (n_rays < 0) { set_error("mfn_composite_train_bw: bad argument"); return MFN_ERR_ARG; }
    if (n_rays == 0) return MFN_OK;
    if (!rays_a || !dL_dopacity || !dL_ddepth || !dL_drgb || !opacity || !depth || !rgb) { set_error("mfn_composite_train_bw: null pointer"); return MFN_ERR_ARG; }
    ProfScope ps("composite_train_bw", (cudaStream_t)stream);
    composite_train_bw_kernel<<<(int)ceil_div(n_rays, kCompWarps), kCompWarps * 32, 0, (cudaStream_t)stream>>>(
        dL_dopacity, dL_ddepth, dL_drgb, dL_dws, sigmas, rgbs, ws, deltas, ts, rays_a, opacity, depth, rgb, T_threshold, n_rays,
        dL_dsigmas, dL_drgbs);
    return check_launch("mfn_composite_train_bw", (cudaStream_t)stream);
}

extern "C" int mfn_composite_test_fw(const float* sigmas, const float* rgbs, const float* deltas, const float* ts,
                                     int64_t* alive_indices, float T_threshold, const int32_t* n_eff_samples, int n_samples,
                                     int64_t n_alive, float* opacity, float* depth, float* rgb, void* stream) {
    if (n_alive < 0 || n_samples < 0) { set_error("mfn_composite_test_fw: bad argument"); return MFN_ERR_ARG; }
    if (n_alive == 0) return MFN_OK;
    if (!alive_indices || !n_eff_samples || !opacity || !depth || !rgb) { set_error("mfn_composite_test_fw: null pointer"); return MFN_ERR_ARG; }
    const int threads = 128;
    composite_test_fw_kernel<<<(int)ceil_div(n_alive, threads), threads, 0, (cudaStream_t)stream>>>(
        sigmas, rgbs, deltas, ts, alive_indices, T_threshold, n_eff_samples, n_samples, n_alive, opacity, depth, rgb);
    return check_launch("mfn_composite_test_fw", (cudaStream_t)stream);
}

extern "C" int mfn_composite_loss_train(const float* sigmas, const float* rgbs, const float* deltas, const float* ts, const int64_t* rays_a,
                                        const float* target, float T_threshold, int64_t n_rays, int64_t n_samples, const float* bg_rgb_host,
                                        float lambda_opacity, float grad_scale, int64_t* total_samples, float* opacity, float* depth, float* rgb,
                                        float* ws, float* rgb_final, float* dL_drgb, float* dL_dopacity, float* dL_dsigmas, float* dL_drgbs,
                                        float* loss_out, float* scratch16, int32_t* clear_flag, void* stream) {
    (void)n_samples;
    if (n_rays < 0) { set_error("mfn_composite_loss_train: bad argument"); return MFN_ERR_ARG; }
    if (n_rays == 0) return MFN_OK;
    if (!rays_a || !target || !bg_rgb_host || !total_samples || !opacity || !depth || !rgb || !ws || !dL_dsigmas || !dL_drgbs || !scratch16) {
        set_error("mfn_composite_loss_train: null pointer"); return MFN_ERR_ARG;
    }
    ProfScope ps("composite_loss_train", (cudaStream_t)stream);
    static int cw = 0;      // rays (warps) per CTA of the fused kernel; measured on B200 (8192 rays): see DESIGN.md
    if (cw == 0) { const char* e = getenv("MFN_COMP_WARPS"); cw = e ? atoi(e) : kCompLossWarps; if (cw < 1 || cw > kCompWarps) cw = kCompLossWarps; }
    composite_loss_train_kernel<<<(int)ceil_div(n_rays, cw), cw * 32, 0, (cudaStream_t)stream>>>(
        sigmas, rgbs, deltas, ts, rays_a, target, T_threshold, n_rays, bg_rgb_host[0], bg_rgb_host[1], bg_rgb_host[2], lambda_opacity, grad_scale,
        total_samples, opacity, depth, rgb, ws, rgb_final, dL_drgb, dL_dopacity, dL_dsigmas, dL_drgbs, loss_out, scratch16, clear_flag);
    return check_launch("mfn_composite_loss_train", (cudaStream_t)stream);
}
