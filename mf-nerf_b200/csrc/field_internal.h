// Internal (C++-side) launch interfaces shared by mlp.cu / encoder.cu / field.cu.  Not part of the C ABI.
#pragma once
#include "common.cuh"
#include "../../include/mfnerf_b200.h"

namespace mfn {

constexpr int kMaxLevels = 32;

struct GridMeta {
    uint32_t offset[kMaxLevels + 1];  // in entries
    uint32_t res[kMaxLevels];
    float scale[kMaxLevels];
    uint32_t hashed;                  // bit l set: level l is hashed
    int n_levels;
    // MixedFeature grid (MFN_GRID_MIXED): several levels share one hash table; size[l] = entries of level l's table
    // (= offset[l+1] - offset[l] for the plain hash grid), canon[l] = scale of the table's canonical (finest) level / scale[l]
    uint32_t size[kMaxLevels];
    float canon[kMaxLevels];
    int mixed;
};
int build_grid_meta(const mfn_grid_cfg* cfg, GridMeta* m, const char* who);

// n = n_dev ? min(*n_dev, n_max) : n_max everywhere: lets a whole training step run without a host sync.
struct MlpFwdArgs {
    const __half* in; int in_stride;       // row stride in halfs, 0 = dense
    const __half* W; int n_hidden; int out_act;
    int64_t n_max; const int32_t* n_dev;
    __half* out; int out_stride;           // 0 = 16
    float* out_rgb32;                      // optional (n,3) fp32 copy of outputs 0..2
    __half* acts;                          // optional (n_hidden, n_max, width)
    const char* tag;                       // profiling name
};
struct MlpBwdArgs {
    const __half* dOut;                    // (n,16) dense
    const __half* in; int in_stride;
    const __half* acts;
    const __half* outv; int out_stride;
    const __half* W; int n_hidden; int out_act;
    int64_t n_max; const int32_t* n_dev;
    __half* dIn; int din_stride;           // optional
    float* dW;
    const char* tag;
};
int mlp_forward(const MlpFwdArgs& a, int in_dim, int width, cudaStream_t st);
int mlp_backward(const MlpBwdArgs& a, int in_dim, int width, cudaStream_t st);

struct EncArgs {
    const float* x; bool normalize; float mn[3], mx[3];   // normalize: x01 = (x - mn) / (mx - mn)
    int64_t n_max; const int32_t* n_dev;
};
int grid_encode_forward(const EncArgs& e, const __half* table, const GridMeta& m, int F, __half* out, cudaStream_t st);
// dT: level-major gradient [L][dT_stride] of half2 (F = 2), loss-scaled
// per-CTA MLP weight-gradient partials of the fused backward kernel, reduced by an extra grid row of the scatter kernel
struct WgradReduce {
    const float* partials; int n_parts; int stride;   // partials[c * stride + j]
    int n_sigma, n_rgb;                               // outputs: d_sigma[0 .. n_sigma), d_rgb[0 .. n_rgb)  (accumulated into)
    float* d_sigma; float* d_rgb;
};
// x01: normalised positions (n,4) f32 (w unused)
int grid_scatter_level_major(const float4* x01, int64_t n_max, const int32_t* n_dev, const __half* dT, int64_t dT_stride, const GridMeta& m, float* dgrid,
                             const WgradReduce& wr, cudaStream_t st);
int grid_encode_backward(const EncArgs& e, const __half* dL_dout, const GridMeta& m, int F, float* dgrid, int32_t* overflow_flag, cudaStream_t st);

// ---- fused tcgen05 field kernels (field_fused.cu) ----
struct FusedArgs {
    const float* xyzs; const float* dirs;
    float mn[3], mx[3];
    int64_t n_max; const int32_t* n_dev;
    const __half* table;
    const __half* w_sigma;   // [W1 64x32][W2 16x64]
    const __half* w_rgb;     // [W3 64x32]([W4 64x64])[W5 16x64]
    float* sigmas; float* rgbs;
    uint2* rgb_h;            // training: the rgb outputs as 3 fp16 (+ pad) per sample, kept in the workspace for the backward pass
    float* dirs_copy;        // training: (n,3) view directions, kept in the workspace (the backward pass re-encodes them; the caller's
                             // array may already hold the next step's samples)
    __half* h_out;           // mode 3: (n, 16) fp16 raw outputs of the sigma network
    int32_t* n_out;          // training: where the forward pass leaves min(*n_dev, n_max) for the backward pass (workspace)
    unsigned char* blobs;    // training: the saved X tiles, 8 KiB per 128 samples
    int rgb_act;
    // backward only
    const float* dL_dsigmas; const float* dL_drgbs; float loss_scale;
    const float* loss_scale_dev;   // when not NULL: the (dynamic) loss scale lives in device memory (mfn_amp_*), read by every CTA
    __half* dfeats;          // level-major [16][dfeats_stride] half2, loss-scaled
    int64_t dfeats_stride;
    float4* x01;             // training: normalised positions (n,4) f32, written by the forward kernel for the scatter kernel
    float* partials;         // [gridDim.x][10240 (rgb width 64) / 25600 (128)] per-CTA weight gradients
    int32_t* overflow;
    long long* dbg;          // optional phase timestamps (tools only)
    // render wavefront (forward MODE 4): rows are (alive slot, sample) pairs; position and direction are derived from the ray and the
    // sample's t instead of being read from per-row arrays, and rows past a ray's N_eff are skipped (the reference evaluates
    // xyzs[valid_mask] only, rendering.py:83-92)
    const float* ray_ts; const float* rays_o; const float* rays_d; const int32_t* ray_plan; const int32_t* ray_alive_lists; const int32_t* ray_n_eff;
    int ray_list_stride;
    unsigned int* sched;     // forward kernel: dynamic tile scheduler slot {next tile, leavers} (set by the launcher), nullptr = static stride
};
bool fused_field_supported(const mfn_field_cfg* c);
size_t fused_blob_bytes(int64_t n_max);
size_t fused_partial_bytes(int rgb_width);
int fused_field_forward(const FusedArgs& a, const GridMeta& m, int rgb_width, int rgb_hidden, int mode, cudaStream_t st);   // mode 0 inference, 1 training, 2 density, 3 raw sigma-net outputs
// launches the backward kernel; the reduction of its per-CTA weight-gradient partials is described in *wr for the scatter kernel to do
int fused_field_backward(const FusedArgs& a, int rgb_width, int rgb_hidden, float* d_sigma_params, float* d_rgb_params, WgradReduce* wr, cudaStream_t st);

}  // namespace mfn
