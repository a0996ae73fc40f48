// The NGP field as one C-ABI op pair: (xyz, dir) -> (sigma, rgb) and its backward.
// Replaces models/networks.py:96-155 (NGP.density / NGP.forward) + TruncExp (custom_functions.py:162-173) on top of
// tcnn:  x01 = (x - xyz_min)/(xyz_max - xyz_min) -> hash grid -> MLP 32->64->16 -> sigma = exp(h[0]);
//        d/|d| -> (d+1)/2 -> SH deg 4 -> MLP [SH(16) | h(16)] -> rgb_width x rgb_layers -> 3 (sigmoid).
// v1 pipeline: encoder / MLP kernels launched back to back on one stream with their intermediates in a caller-provided
// workspace; every kernel takes the sample count from device memory (n_dev) so a whole training step needs no host sync.
#include "field_internal.h"
#include "sh4.cuh"
#include <stdlib.h>

namespace mfn {

// MFN_FIELD_IMPL=v1 selects the unfused mma.sync pipeline below (kept for A/B measurements; every shape field_cfg_ok accepts is fused)
static bool use_fused(const mfn_field_cfg* c) {
    static int v1 = -1;
    if (v1 < 0) { const char* e = getenv("MFN_FIELD_IMPL"); v1 = (e && e[0] == 'v' && e[1] == '1') ? 1 : 0; }
    return !v1 && fused_field_supported(c);
}

// workspace of the fused path: [saved X tiles | rgb outputs (n,4) f16 | dfeats (n,32) f16 | weight-gradient partials | x01 (n,4) f32 | dirs (n,3) f32 | count]
struct FusedWs { size_t blobs, rgb, dfeats, partials, x01, dirs, count, total; };
static FusedWs fused_ws(const mfn_field_cfg* c, int64_t n, bool training) {
    FusedWs w{};
    size_t o = 0;
    if (training) {
        w.blobs = o; o += fused_blob_bytes(n);
        w.rgb = o; o += (size_t)(n * 8 + 255) / 256 * 256;
        w.dfeats = o; o += (size_t)((n + 63) / 64 * 64) * 64;
        w.partials = o; o += (fused_partial_bytes(c->rgb_width) + 255) / 256 * 256;
        w.x01 = o; o += (size_t)(n * 16 + 255) / 256 * 256;
        w.dirs = o; o += (size_t)(n * 12 + 255) / 256 * 256;
        w.count = o; o += 256;         // int32: the sample count of the forward pass (mfn_field_count_ptr)
    }
    w.total = o > 256 ? o : 256;
    return w;
}

struct FieldWs {
    size_t feats, acts1, cat, acts2, out2, dout2, dcat, dh, dfeats, out1, total;
};
static inline size_t al(size_t x) { return (x + 255) / 256 * 256; }
static FieldWs field_ws(const mfn_field_cfg* c, int64_t n, bool training) {
    FieldWs w{};
    const size_t n_enc = (size_t)c->grid.n_levels * c->grid.n_features;
    size_t o = 0;
    w.feats = o; o += al(n * n_enc * 2);
    w.acts1 = o; o += al((size_t)c->sigma_hidden * n * c->sigma_width * 2);
    w.cat = o; o += al(n * 32 * 2);
    w.acts2 = o; o += al((size_t)c->rgb_hidden * n * c->rgb_width * 2);
    w.out2 = o; o += al(n * 16 * 2);
    w.out1 = o; o += al(n * 16 * 2);
    if (training) {
        w.dout2 = o; o += al(n * 16 * 2);
        w.dcat = o; o += al(n * 32 * 2);
        w.dh = o; o += al(n * 16 * 2);
        w.dfeats = o; o += al(n * n_enc * 2);
    }
    w.total = o;
    return w;
}

static int field_cfg_ok(const mfn_field_cfg* c, const char* who) {
    if (!c) { set_error("%s: null config", who); return MFN_ERR_ARG; }
    const int n_enc = c->grid.n_levels * c->grid.n_features;
    if (n_enc != 32) { set_error("%s: n_levels*n_features_per_level must be 32 (got %d)", who, n_enc); return MFN_ERR_ARG; }
    if (c->sigma_width != 64 || c->sigma_hidden != 1) { set_error("%s: sigma net must be 64 wide with 1 hidden layer", who); return MFN_ERR_ARG; }
    if ((c->rgb_width != 64 && c->rgb_width != 128) || c->rgb_hidden < 1 || c->rgb_hidden > 2) { set_error("%s: rgb net must be 64/128 wide with 1-2 hidden layers", who); return MFN_ERR_ARG; }
    return MFN_OK;
}

// SH of the normalised direction into cat[:, 0:16] and sigma = exp(h0) with h0 = cat[:, 16]
// (networks.py:107 TruncExp forward = plain exp; :145-146 direction normalisation and (d+1)/2 mapping)

__global__ void sh_sigma_kernel(const float* __restrict__ dirs, int64_t n_max, const int32_t* __restrict__ n_dev, __half* __restrict__ cat,
                                float* __restrict__ sigmas) {
    const int64_t n = n_dev ? min((int64_t)*n_dev, n_max) : n_max;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float dx = dirs[3 * i], dy = dirs[3 * i + 1], dz = dirs[3 * i + 2];
        const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);
        const float ux = (dx / nrm + 1.0f) / 2.0f, uy = (dy / nrm + 1.0f) / 2.0f, uz = (dz / nrm + 1.0f) / 2.0f;
        float o[16];
        sh4_eval(fmaf(ux, 2.f, -1.f), fmaf(uy, 2.f, -1.f), fmaf(uz, 2.f, -1.f), o);
        __align__(16) __half h[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) h[k] = __float2half_rn(o[k]);
        uint4* dst = reinterpret_cast<uint4*>(cat + (size_t)i * 32);
        dst[0] = reinterpret_cast<const uint4*>(h)[0];
        dst[1] = reinterpret_cast<const uint4*>(h)[1];
        sigmas[i] = expf(__half2float(cat[(size_t)i * 32 + 16]));
    }
}

__global__ void exp_h0_kernel(const __half* __restrict__ out1, int64_t n_max, const int32_t* __restrict__ n_dev, float* __restrict__ sigmas) {
    const int64_t n = n_dev ? min((int64_t)*n_dev, n_max) : n_max;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        sigmas[i] = expf(__half2float(out1[(size_t)i * 16]));
}

// dL/d(rgb net output) in fp16, scaled: [dL_drgbs * loss_scale, 0 x 13]
__global__ void prep_dout2_kernel(const float* __restrict__ dL_drgbs, float loss_scale, int64_t n_max, const int32_t* __restrict__ n_dev,
                                  __half* __restrict__ dout2, int32_t* __restrict__ overflow) {
    const int64_t n = n_dev ? min((int64_t)*n_dev, n_max) : n_max;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        __align__(16) __half h[16];
        bool bad = false;
#pragma unroll
        for (int k = 0; k < 3; ++k) { h[k] = __float2half_rn(dL_drgbs[3 * i + k] * loss_scale); bad |= !isfinite(__half2float(h[k])); }
#pragma unroll
        for (int k = 3; k < 16; ++k) h[k] = __float2half_rn(0.f);
        uint4* dst = reinterpret_cast<uint4*>(dout2 + (size_t)i * 16);
        dst[0] = reinterpret_cast<const uint4*>(h)[0];
        dst[1] = reinterpret_cast<const uint4*>(h)[1];
        if (bad && overflow) *overflow = 1;
    }
}

// dL/dh = (rgb-net input gradient)[16:32] + e0 * dL/dsigma * exp(clamp(h0, -15, 15)) * loss_scale   (TruncExp backward,
// custom_functions.py:170-173)
__global__ void merge_dh_kernel(const __half* __restrict__ dcat, const __half* __restrict__ cat, const float* __restrict__ dL_dsigmas, float loss_scale,
                                int64_t n_max, const int32_t* __restrict__ n_dev, __half* __restrict__ dh, int32_t* __restrict__ overflow) {
    const int64_t n = n_dev ? min((int64_t)*n_dev, n_max) : n_max;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        __align__(16) __half h[16];
        const uint4* src = reinterpret_cast<const uint4*>(dcat + (size_t)i * 32 + 16);
        reinterpret_cast<uint4*>(h)[0] = src[0];
        reinterpret_cast<uint4*>(h)[1] = src[1];
        const float h0 = __half2float(cat[(size_t)i * 32 + 16]);
        const float g = dL_dsigmas[i] * expf(fminf(fmaxf(h0, -15.f), 15.f)) * loss_scale;
        h[0] = __float2half_rn(__half2float(h[0]) + g);
        if (!isfinite(__half2float(h[0])) && overflow) *overflow = 1;
        uint4* dst = reinterpret_cast<uint4*>(dh + (size_t)i * 16);
        dst[0] = reinterpret_cast<const uint4*>(h)[0];
        dst[1] = reinterpret_cast<const uint4*>(h)[1];
    }
}

static inline unsigned ew_grid(int64_t n_max) {
    int64_t b = ceil_div(n_max, 256);
    const int64_t cap = (int64_t)kNumSMs * 8;
    return (unsigned)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace mfn

using namespace mfn;

static size_t ws_need(const mfn_field_cfg* cfg, int64_t n_max, bool training) {
    return use_fused(cfg) ? fused_ws(cfg, n_max, training).total : field_ws(cfg, n_max, training).total;
}

extern "C" int mfn_field_is_fused(const mfn_field_cfg* cfg) {
    if (field_cfg_ok(cfg, "mfn_field_is_fused") != MFN_OK) return MFN_ERR_ARG;
    return use_fused(cfg) ? 1 : 0;
}

extern "C" void* mfn_field_count_ptr(const mfn_field_cfg* cfg, void* workspace, int64_t n_max) {
    if (field_cfg_ok(cfg, "mfn_field_count_ptr") != MFN_OK || !workspace || n_max < 0 || !use_fused(cfg)) return nullptr;
    return (char*)workspace + fused_ws(cfg, n_max, true).count;
}

extern "C" int64_t mfn_field_workspace_bytes(const mfn_field_cfg* cfg, int64_t n_max, int training) {
    if (field_cfg_ok(cfg, "mfn_field_workspace_bytes") != MFN_OK || n_max < 0) return -1;
    // the fused kernels keep ~170 B/sample in training and nothing at inference; the v1 pipeline (MFN_FIELD_IMPL=v1) every intermediate
    return (int64_t)ws_need(cfg, n_max, training != 0);
}

static void make_enc(EncArgs& e, const mfn_field_cfg* cfg, const float* xyzs, int64_t n_max, const int32_t* n_dev) {
    e.x = xyzs; e.normalize = true; e.n_max = n_max; e.n_dev = n_dev;
    for (int k = 0; k < 3; ++k) { e.mn[k] = cfg->xyz_min[k]; e.mx[k] = cfg->xyz_max[k]; }
}

static void make_fused(FusedArgs& f, const mfn_field_cfg* cfg, const void* xyz_params_h, const void* rgb_params_h, const float* xyzs, const float* dirs,
                       int64_t n_max, const int32_t* n_dev) {
    f = FusedArgs{};
    f.xyzs = xyzs; f.dirs = dirs; f.n_max = n_max; f.n_dev = n_dev;
    for (int k = 0; k < 3; ++k) { f.mn[k] = cfg->xyz_min[k]; f.mx[k] = cfg->xyz_max[k]; }
    f.w_sigma = (const __half*)xyz_params_h;
    f.table = f.w_sigma + 64 * 32 + 16 * 64;
    f.w_rgb = (const __half*)rgb_params_h;
    f.rgb_act = cfg->rgb_act;
}

extern "C" int mfn_field_fwd(const mfn_field_cfg* cfg, const void* xyz_params_h, const void* rgb_params_h, const float* xyzs, const float* dirs,
                             int64_t n_max, const int32_t* n_dev, float* sigmas, float* rgbs, void* workspace, int64_t workspace_bytes,
                             void* stream) {
    int rc = field_cfg_ok(cfg, "mfn_field_fwd");
    if (rc != MFN_OK) return rc;
    if (n_max < 0) { set_error("mfn_field_fwd: bad n_max"); return MFN_ERR_ARG; }
    if (n_max == 0) return MFN_OK;
    const FieldWs w = field_ws(cfg, n_max, false);
    if (!xyz_params_h || !rgb_params_h || !xyzs || !dirs || !sigmas || !rgbs || !workspace || (size_t)workspace_bytes < ws_need(cfg, n_max, false)) {
        set_error("mfn_field_fwd: null pointer or workspace too small"); return MFN_ERR_ARG;
    }
    GridMeta m;
    if ((rc = build_grid_meta(&cfg->grid, &m, "mfn_field_fwd")) != MFN_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)workspace;
    if (use_fused(cfg)) {
        // training layout when the workspace is large enough for it (the matching mfn_field_bwd reads it), inference otherwise
        const FusedWs fw = fused_ws(cfg, n_max, true);
        const bool train = (size_t)workspace_bytes >= fw.total;
        FusedArgs f; make_fused(f, cfg, xyz_params_h, rgb_params_h, xyzs, dirs, n_max, n_dev);
        f.sigmas = sigmas; f.rgbs = rgbs;
        if (train) { f.blobs = (unsigned char*)ws + fw.blobs; f.rgb_h = (uint2*)(ws + fw.rgb); f.dirs_copy = (float*)(ws + fw.dirs); f.x01 = (float4*)(ws + fw.x01); f.n_out = (int32_t*)(ws + fw.count); }
        { static const char* dbg_env = getenv("MFN_FWD_DBG"); if (dbg_env) f.dbg = (long long*)strtoull(dbg_env, nullptr, 0); }
        return fused_field_forward(f, m, cfg->rgb_width, cfg->rgb_hidden, train ? 1 : 0, st);
    }
    const int n_mlp1 = 64 * 32 + 16 * 64;
    const __half* p = (const __half*)xyz_params_h;
    EncArgs e; make_enc(e, cfg, xyzs, n_max, n_dev);
    if ((rc = grid_encode_forward(e, p + n_mlp1, m, cfg->grid.n_features, (__half*)(ws + w.feats), st)) != MFN_OK) return rc;
    MlpFwdArgs a{};
    a.in = (const __half*)(ws + w.feats); a.W = p; a.n_hidden = 1; a.out_act = MFN_ACT_NONE; a.n_max = n_max; a.n_dev = n_dev;
    a.out = (__half*)(ws + w.cat) + 16; a.out_stride = 32; a.acts = (__half*)(ws + w.acts1); a.tag = "mlp_sigma_fwd";
    if ((rc = mlp_forward(a, 32, 64, st)) != MFN_OK) return rc;
    { ProfScope ps("sh_sigma", st);
      sh_sigma_kernel<<<ew_grid(n_max), 256, 0, st>>>(dirs, n_max, n_dev, (__half*)(ws + w.cat), sigmas); }
    MlpFwdArgs b{};
    b.in = (const __half*)(ws + w.cat); b.W = (const __half*)rgb_params_h; b.n_hidden = cfg->rgb_hidden; b.out_act = cfg->rgb_act;
    b.n_max = n_max; b.n_dev = n_dev; b.out = (__half*)(ws + w.out2); b.out_rgb32 = rgbs; b.acts = (__half*)(ws + w.acts2); b.tag = "mlp_rgb_fwd";
    if ((rc = mlp_forward(b, 32, cfg->rgb_width, st)) != MFN_OK) return rc;
    return check_launch("mfn_field_fwd", st);
}

static int field_bwd_impl(const mfn_field_cfg* cfg, const void* xyz_params_h, const void* rgb_params_h, const float* xyzs, int64_t n_max,
                          const int32_t* n_dev, const float* dL_dsigmas, const float* dL_drgbs, float loss_scale, const float* loss_scale_dev,
                          float* d_xyz_params, float* d_rgb_params, int32_t* overflow_flag, void* workspace, int64_t workspace_bytes, void* stream);

extern "C" int mfn_field_bwd(const mfn_field_cfg* cfg, const void* xyz_params_h, const void* rgb_params_h, const float* xyzs, int64_t n_max,
                             const int32_t* n_dev, const float* dL_dsigmas, const float* dL_drgbs, float loss_scale, float* d_xyz_params,
                             float* d_rgb_params, int32_t* overflow_flag, void* workspace, int64_t workspace_bytes, void* stream) {
    return field_bwd_impl(cfg, xyz_params_h, rgb_params_h, xyzs, n_max, n_dev, dL_dsigmas, dL_drgbs, loss_scale, nullptr, d_xyz_params, d_rgb_params,
                          overflow_flag, workspace, workspace_bytes, stream);
}

extern "C" int mfn_field_bwd_amp(const mfn_field_cfg* cfg, const void* xyz_params_h, const void* rgb_params_h, const float* xyzs, int64_t n_max,
                                 const int32_t* n_dev, const float* dL_dsigmas, const float* dL_drgbs, const float* amp_state, float* d_xyz_params,
                                 float* d_rgb_params, int32_t* overflow_flag, void* workspace, int64_t workspace_bytes, void* stream) {
    if (!amp_state) { set_error("mfn_field_bwd_amp: null pointer"); return MFN_ERR_ARG; }
    if (field_cfg_ok(cfg, "mfn_field_bwd_amp") == MFN_OK && !use_fused(cfg)) {
        set_error("mfn_field_bwd_amp: the device-side loss scale is read by the fused kernels only (MFN_FIELD_IMPL=v1 takes mfn_field_bwd)"); return MFN_ERR_ARG;
    }
    return field_bwd_impl(cfg, xyz_params_h, rgb_params_h, xyzs, n_max, n_dev, dL_dsigmas, dL_drgbs, 0.f, amp_state, d_xyz_params, d_rgb_params,
                          overflow_flag, workspace, workspace_bytes, stream);
}

static int field_bwd_impl(const mfn_field_cfg* cfg, const void* xyz_params_h, const void* rgb_params_h, const float* xyzs, int64_t n_max,
                          const int32_t* n_dev, const float* dL_dsigmas, const float* dL_drgbs, float loss_scale, const float* loss_scale_dev,
                          float* d_xyz_params, float* d_rgb_params, int32_t* overflow_flag, void* workspace, int64_t workspace_bytes, void* stream) {
    int rc = field_cfg_ok(cfg, "mfn_field_bwd");
    if (rc != MFN_OK) return rc;
    if (n_max < 0) { set_error("mfn_field_bwd: bad n_max"); return MFN_ERR_ARG; }
    if (n_max == 0) return MFN_OK;
    const FieldWs w = field_ws(cfg, n_max, true);
    if (!xyz_params_h || !rgb_params_h || !xyzs || !dL_dsigmas || !dL_drgbs || !d_xyz_params || !d_rgb_params || !workspace ||
        (size_t)workspace_bytes < ws_need(cfg, n_max, true)) { set_error("mfn_field_bwd: null pointer or workspace too small"); return MFN_ERR_ARG; }
    GridMeta m;
    if ((rc = build_grid_meta(&cfg->grid, &m, "mfn_field_bwd")) != MFN_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)workspace;
    if (use_fused(cfg)) {
        const FusedWs fw = fused_ws(cfg, n_max, true);
        if ((size_t)workspace_bytes < fw.total) { set_error("mfn_field_bwd: workspace too small"); return MFN_ERR_ARG; }
        FusedArgs f; make_fused(f, cfg, xyz_params_h, rgb_params_h, xyzs, nullptr, n_max, n_dev);
        f.blobs = (unsigned char*)ws + fw.blobs; f.rgb_h = (uint2*)(ws + fw.rgb); f.dirs_copy = (float*)(ws + fw.dirs); f.dfeats = (__half*)(ws + fw.dfeats);
        f.dfeats_stride = (n_max + 63) / 64 * 64;
        f.partials = (float*)(ws + fw.partials); f.dL_dsigmas = dL_dsigmas; f.dL_drgbs = dL_drgbs; f.loss_scale = loss_scale; f.loss_scale_dev = loss_scale_dev; f.overflow = overflow_flag;
        { static const char* dbg_env = getenv("MFN_FWD_DBG"); if (dbg_env) f.dbg = (long long*)strtoull(dbg_env, nullptr, 0); }
        WgradReduce wr{};
        if ((rc = fused_field_backward(f, cfg->rgb_width, cfg->rgb_hidden, d_xyz_params, d_rgb_params, &wr, st)) != MFN_OK) return rc;
        return grid_scatter_level_major((const float4*)(ws + fw.x01), n_max, n_dev, f.dfeats, f.dfeats_stride, m, d_xyz_params + 64 * 32 + 16 * 64, wr, st);
    }
    const int n_mlp1 = 64 * 32 + 16 * 64;
    const __half* p = (const __half*)xyz_params_h;
    note_launch(2);  // prep_dout2 + merge_dh
    { ProfScope ps("prep_dout2", st);
      prep_dout2_kernel<<<ew_grid(n_max), 256, 0, st>>>(dL_drgbs, loss_scale, n_max, n_dev, (__half*)(ws + w.dout2), overflow_flag); }
    MlpBwdArgs b{};
    b.dOut = (const __half*)(ws + w.dout2); b.in = (const __half*)(ws + w.cat); b.acts = (const __half*)(ws + w.acts2);
    b.outv = (const __half*)(ws + w.out2); b.W = (const __half*)rgb_params_h; b.n_hidden = cfg->rgb_hidden; b.out_act = cfg->rgb_act;
    b.n_max = n_max; b.n_dev = n_dev; b.dIn = (__half*)(ws + w.dcat); b.dW = d_rgb_params; b.tag = "mlp_rgb_bwd";
    if ((rc = mlp_backward(b, 32, cfg->rgb_width, st)) != MFN_OK) return rc;
    { ProfScope ps("merge_dh", st);
      merge_dh_kernel<<<ew_grid(n_max), 256, 0, st>>>((const __half*)(ws + w.dcat), (const __half*)(ws + w.cat), dL_dsigmas, loss_scale, n_max, n_dev,
                                                     (__half*)(ws + w.dh), overflow_flag); }
    MlpBwdArgs a{};
    a.dOut = (const __half*)(ws + w.dh); a.in = (const __half*)(ws + w.feats); a.acts = (const __half*)(ws + w.acts1);
    a.outv = nullptr; a.W = p; a.n_hidden = 1; a.out_act = MFN_ACT_NONE; a.n_max = n_max; a.n_dev = n_dev;
    a.dIn = (__half*)(ws + w.dfeats); a.dW = d_xyz_params; a.tag = "mlp_sigma_bwd";
    if ((rc = mlp_backward(a, 32, 64, st)) != MFN_OK) return rc;
    EncArgs e; make_enc(e, cfg, xyzs, n_max, n_dev);
    return grid_encode_backward(e, (const __half*)(ws + w.dfeats), m, cfg->grid.n_features, d_xyz_params + n_mlp1, overflow_flag, st);
}

extern "C" int mfn_geo_fwd(const mfn_field_cfg* cfg, const void* xyz_params_h, const float* xyzs, int64_t n_max, const int32_t* n_dev,
                           void* h_out, void* stream) {
    int rc = field_cfg_ok(cfg, "mfn_geo_fwd");
    if (rc != MFN_OK) return rc;
    if (n_max < 0) { set_error("mfn_geo_fwd: bad n_max"); return MFN_ERR_ARG; }
    if (!use_fused(cfg)) { set_error("mfn_geo_fwd: only the fused shape (16 levels x 2 features, 64-wide one-hidden-layer network)"); return MFN_ERR_ARG; }
    if (n_max == 0) return MFN_OK;
    if (!xyz_params_h || !xyzs || !h_out) { set_error("mfn_geo_fwd: null pointer"); return MFN_ERR_ARG; }
    GridMeta m;
    if ((rc = build_grid_meta(&cfg->grid, &m, "mfn_geo_fwd")) != MFN_OK) return rc;
    FusedArgs f; make_fused(f, cfg, xyz_params_h, nullptr, xyzs, nullptr, n_max, n_dev);
    f.h_out = (__half*)h_out;
    return fused_field_forward(f, m, cfg->rgb_width, cfg->rgb_hidden, 3, (cudaStream_t)stream);
}

extern "C" int mfn_density_fwd(const mfn_field_cfg* cfg, const void* xyz_params_h, const float* xyzs, int64_t n_max, const int32_t* n_dev,
                               float* sigmas, void* workspace, int64_t workspace_bytes, void* stream) {
    int rc = field_cfg_ok(cfg, "mfn_density_fwd");
    if (rc != MFN_OK) return rc;
    if (n_max < 0) { set_error("mfn_density_fwd: bad n_max"); return MFN_ERR_ARG; }
    if (n_max == 0) return MFN_OK;
    const FieldWs w = field_ws(cfg, n_max, false);
    if (!xyz_params_h || !xyzs || !sigmas || !workspace || (size_t)workspace_bytes < ws_need(cfg, n_max, false)) {
        set_error("mfn_density_fwd: null pointer or workspace too small"); return MFN_ERR_ARG;
    }
    GridMeta m;
    if ((rc = build_grid_meta(&cfg->grid, &m, "mfn_density_fwd")) != MFN_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)workspace;
    if (use_fused(cfg)) {
        FusedArgs f; make_fused(f, cfg, xyz_params_h, nullptr, xyzs, nullptr, n_max, n_dev);
        f.sigmas = sigmas;
        return fused_field_forward(f, m, cfg->rgb_width, cfg->rgb_hidden, 2, st);
    }
    const int n_mlp1 = 64 * 32 + 16 * 64;
    const __half* p = (const __half*)xyz_params_h;
    EncArgs e; make_enc(e, cfg, xyzs, n_max, n_dev);
    if ((rc = grid_encode_forward(e, p + n_mlp1, m, cfg->grid.n_features, (__half*)(ws + w.feats), st)) != MFN_OK) return rc;
    MlpFwdArgs a{};
    a.in = (const __half*)(ws + w.feats); a.W = p; a.n_hidden = 1; a.out_act = MFN_ACT_NONE; a.n_max = n_max; a.n_dev = n_dev;
    a.out = (__half*)(ws + w.out1); a.tag = "mlp_sigma_fwd";
    if ((rc = mlp_forward(a, 32, 64, st)) != MFN_OK) return rc;
    exp_h0_kernel<<<ew_grid(n_max), 256, 0, st>>>((const __half*)(ws + w.out1), n_max, n_dev, sigmas);
    return check_launch("mfn_density_fwd", st);
}
