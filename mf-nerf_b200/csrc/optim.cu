// Fused Adam over the flat fp32 parameter vectors (train.py:136 apex FusedAdam(lr, eps=1e-15),
// betas (0.9, 0.999), no weight decay, bias correction on) with the AMP bookkeeping folded in: gradient unscale,
// skip-on-overflow (GradScaler semantics), refresh of the fp16 shadow parameters the field kernels read, and
// zeroing of the gradient buffer for the next step.  HBM-bound: 30 B/param (p,g,m,v read; p,m,v,p16 written; g zeroed).
#include "common.cuh"
#include "umma.cuh"
#include <stdlib.h>
#include "../../include/mfnerf_b200.h"

namespace mfn {

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, __half* __restrict__ p_h, int64_t n,
            float lr, float beta1, float beta2, float eps, float bc1, float bc2, float grad_scale, const int32_t* __restrict__ skip_flag, int zero_grad,
            const float* __restrict__ hyper, const float* __restrict__ amp, float inv_world) {
    if (hyper) { lr = hyper[0]; bc1 = hyper[1]; bc2 = hyper[2]; }      // per-step scalars from device memory: the launch can live in a CUDA graph
    const bool skip = skip_flag && *skip_flag != 0;
    if (amp) {
        // AMP state in device memory (mfn_amp_*): the loss scale is dynamic and the bias corrections count APPLIED steps only (a skipped
        // step does not advance Adam's t -- torch.optim / GradScaler semantics); mfn_amp_update leaves the corrections of the next
        // step in amp[4], amp[5] (evaluated in double by one thread), so nothing here is more than a load
        bc1 = amp[4]; bc2 = amp[5];
        grad_scale = inv_world / amp[0];
    }
    const uint64_t pol_keep = umma::policy_evict_last();   // the fp16 shadow (hash table + weights) is what the next step gathers from: keep it in L2
    const int64_t n4 = n / 4;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 gp = reinterpret_cast<float4*>(g)[i];
        if (!skip) {
            float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
            float* P = &pp.x; float* G = &gp.x; float* M = &mm.x; float* V = &vv.x;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float gr = G[k] * grad_scale;
                M[k] = beta1 * M[k] + (1.f - beta1) * gr;
                V[k] = beta2 * V[k] + (1.f - beta2) * gr * gr;
                P[k] -= lr * (M[k] / bc1) / (sqrtf(V[k] / bc2) + eps);
            }
            reinterpret_cast<float4*>(p)[i] = pp; reinterpret_cast<float4*>(m)[i] = mm; reinterpret_cast<float4*>(v)[i] = vv;
            if (p_h) {
                __half2 lo = __floats2half2_rn(pp.x, pp.y), hi = __floats2half2_rn(pp.z, pp.w);
                umma::st_hint_b64(reinterpret_cast<uint2*>(p_h) + i, *reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi), pol_keep);
            }
        }
        if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // tail (n % 4)
    for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (!skip) {
            const float gr = g[i] * grad_scale;
            const float mi = beta1 * m[i] + (1.f - beta1) * gr, vi = beta2 * v[i] + (1.f - beta2) * gr * gr;
            m[i] = mi; v[i] = vi;
            const float pi = p[i] - lr * (mi / bc1) / (sqrtf(vi / bc2) + eps);
            p[i] = pi;
            if (p_h) p_h[i] = __float2half_rn(pi);
        }
        if (zero_grad) g[i] = 0.f;
    }
}

// GradScaler bookkeeping after an optimiser step (torch.amp.GradScaler.update, the reference trains under Lightning precision=16):
// overflow -> the step was skipped: scale *= backoff, growth tracker reset; otherwise the tracker advances and every `interval` good
// steps the scale grows.  amp = {scale, growth tracker, skipped steps (total), applied steps (total)}.
__global__ void amp_update_kernel(float* __restrict__ amp, const int32_t* __restrict__ skip_flag, float backoff, float growth, float interval,
                                  float min_scale, float max_scale, float beta1, float beta2, int init_only) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (!init_only) {
        if (skip_flag && *skip_flag != 0) {
            amp[0] = fmaxf(amp[0] * backoff, min_scale); amp[1] = 0.f; amp[2] += 1.f;
        } else {
            amp[3] += 1.f;
            const float g = amp[1] + 1.f;
            if (g >= interval) { amp[0] = fminf(amp[0] * growth, max_scale); amp[1] = 0.f; } else amp[1] = g;
        }
    }
    const double t = (double)amp[3] + 1.0;      // bias corrections of the NEXT step, in double from the float betas (apex's arithmetic)
    amp[4] = (float)(1.0 - pow((double)beta1, t)); amp[5] = (float)(1.0 - pow((double)beta2, t));
}

__global__ void cast_f32_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = __float2half_rn(src[i]);
}

}  // namespace mfn

using namespace mfn;

static int adam_launch(float* params, float* grads, float* exp_avg, float* exp_avg_sq, void* params_h, int64_t n, float lr, float beta1, float beta2,
                       float eps, float bc1, float bc2, const float* hyper, float grad_scale, const int32_t* skip_flag, int zero_grad, void* stream,
                       const float* amp = nullptr, float inv_world = 1.f);

extern "C" int mfn_adam_hyper(float lr, float beta1, float beta2, int step, float* hyper_host) {
    if (step < 1 || !hyper_host) { set_error("mfn_adam_hyper: bad argument"); return MFN_ERR_ARG; }
    // bias corrections evaluated in double from the float betas, like apex's multi_tensor_adam (1 - std::pow(float beta, int step))
    hyper_host[0] = lr; hyper_host[1] = (float)(1.0 - pow((double)beta1, (double)step)); hyper_host[2] = (float)(1.0 - pow((double)beta2, (double)step));
    return MFN_OK;
}

extern "C" int mfn_adam_step_dev(float* params, float* grads, float* exp_avg, float* exp_avg_sq, void* params_h, int64_t n, const float* hyper_dev,
                                 float beta1, float beta2, float eps, float grad_scale, const int32_t* skip_flag, int zero_grad, void* stream) {
    if (!hyper_dev) { set_error("mfn_adam_step_dev: null pointer"); return MFN_ERR_ARG; }
    return adam_launch(params, grads, exp_avg, exp_avg_sq, params_h, n, 0.f, beta1, beta2, eps, 1.f, 1.f, hyper_dev, grad_scale, skip_flag, zero_grad, stream);
}

extern "C" int mfn_adam_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, void* params_h, int64_t n, float lr, float beta1,
                             float beta2, float eps, int step, float grad_scale, const int32_t* skip_flag, int zero_grad, void* stream) {
    if (step < 1) { set_error("mfn_adam_step: bad argument"); return MFN_ERR_ARG; }
    float h[3];
    mfn_adam_hyper(lr, beta1, beta2, step, h);
    return adam_launch(params, grads, exp_avg, exp_avg_sq, params_h, n, h[0], beta1, beta2, eps, h[1], h[2], nullptr, grad_scale, skip_flag, zero_grad, stream);
}

static int adam_launch(float* params, float* grads, float* exp_avg, float* exp_avg_sq, void* params_h, int64_t n, float lr, float beta1, float beta2,
                       float eps, float bc1, float bc2, const float* hyper, float grad_scale, const int32_t* skip_flag, int zero_grad, void* stream,
                       const float* amp, float inv_world) {
    if (n < 0) { set_error("mfn_adam_step: bad argument"); return MFN_ERR_ARG; }
    if (n == 0) return MFN_OK;
    if (!params || !grads || !exp_avg || !exp_avg_sq) { set_error("mfn_adam_step: null pointer"); return MFN_ERR_ARG; }
    if ((((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) || ((uintptr_t)params_h & 7)) {
        set_error("mfn_adam_step: buffers must be 16-byte aligned"); return MFN_ERR_ARG;
    }
    int64_t blocks = ceil_div(n / 4 + 1, 256);
    static int per_sm = 0;
    if (per_sm == 0) { const char* e = getenv("MFN_ADAM_BPS"); per_sm = e ? atoi(e) : 32; if (per_sm < 1) per_sm = 32; }      // measured: 8 -> 75 us, 16 -> 64 us, 32 -> 63 us (11.4 M params)
    const int64_t cap = (int64_t)kNumSMs * per_sm;
    if (blocks > cap) blocks = cap;
    ProfScope ps("adam", (cudaStream_t)stream);
    adam_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, (__half*)params_h, n, lr, beta1, beta2, eps, bc1,
                                                                    bc2, grad_scale, skip_flag, zero_grad, hyper, amp, inv_world);
    return check_launch("mfn_adam_step", (cudaStream_t)stream);
}

extern "C" int mfn_adam_step_amp(float* params, float* grads, float* exp_avg, float* exp_avg_sq, void* params_h, int64_t n, const float* lr_dev,
                                 float beta1, float beta2, float eps, float inv_world, const float* amp_state, const int32_t* skip_flag, int zero_grad,
                                 void* stream) {
    if (!lr_dev || !amp_state) { set_error("mfn_adam_step_amp: null pointer"); return MFN_ERR_ARG; }
    return adam_launch(params, grads, exp_avg, exp_avg_sq, params_h, n, 0.f, beta1, beta2, eps, 1.f, 1.f, lr_dev, 1.f, skip_flag, zero_grad, stream, amp_state,
                       inv_world);
}

extern "C" int mfn_amp_update(float* amp_state, const int32_t* skip_flag, float backoff, float growth, int growth_interval, float min_scale, float max_scale,
                              float beta1, float beta2, void* stream) {
    if (!amp_state || !(backoff > 0.f) || !(growth >= 1.f) || growth_interval < 1 || !(min_scale > 0.f) || max_scale < min_scale) {
        set_error("mfn_amp_update: bad argument"); return MFN_ERR_ARG;
    }
    amp_update_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(amp_state, skip_flag, backoff, growth, (float)growth_interval, min_scale, max_scale, beta1, beta2, 0);
    return check_launch("mfn_amp_update", (cudaStream_t)stream);
}

extern "C" int mfn_amp_init(float* amp_state, float loss_scale, int applied_steps, float beta1, float beta2, void* stream) {
    if (!amp_state || !(loss_scale > 0.f) || applied_steps < 0) { set_error("mfn_amp_init: bad argument"); return MFN_ERR_ARG; }
    const float h[8] = {loss_scale, 0.f, 0.f, (float)applied_steps, 0.f, 0.f, 0.f, 0.f};
    if (cudaMemcpyAsync(amp_state, h, sizeof(h), cudaMemcpyHostToDevice, (cudaStream_t)stream) != cudaSuccess) { set_error("mfn_amp_init: copy failed"); return MFN_ERR_CUDA; }
    amp_update_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(amp_state, nullptr, 1.f, 1.f, 1.f, 1.f, 1.f, beta1, beta2, 1);
    return check_launch("mfn_amp_init", (cudaStream_t)stream);
}

extern "C" int mfn_cast_f32_to_f16(const float* src, void* dst, int64_t n, void* stream) {
    if (n < 0 || (n > 0 && (!src || !dst))) { set_error("mfn_cast_f32_to_f16: bad argument"); return MFN_ERR_ARG; }
    if (n == 0) return MFN_OK;
    int64_t blocks = ceil_div(n, 256);
    const int64_t cap = (int64_t)kNumSMs * 8;
    if (blocks > cap) blocks = cap;
    cast_f32_f16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, (__half*)dst, n);
    return check_launch("mfn_cast_f32_to_f16", (cudaStream_t)stream);
}
