/* mfnerf_b200.h -- C ABI of libmfnerf_b200.so, the sm_100a (B200) implementation of MF-NeRF's per-ray /
 * per-sample hot path.  This is the drop-in boundary that replaces the reference's pybind11 module `vren`
 * (models/csrc/binding.cpp:234-251) and the tiny-cuda-nn calls of models/networks.py:36-94.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is DEVICE memory unless the name ends in `_host`;
 *   - the library never allocates user-visible memory: outputs and workspaces are passed in
 *     (query *_workspace_bytes first); nothing is zero-filled by the caller unless stated;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, no call synchronises;
 *   - return value: 0 on success, <0 on error (MFN_ERR_*), text via mfn_last_error() (thread-local);
 *   - set MFN_DEBUG_SYNC=1 to synchronise and check after every launch.
 * Layouts are the reference's: row-major contiguous, (N,3) float triples, rays_a = (R,3) int64
 * {ray_idx, start_idx, N_samples}.
 */
#ifndef MFNERF_B200_H_
#define MFNERF_B200_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MFN_VERSION 100

#define MFN_DTYPE_F32 0
#define MFN_DTYPE_F16 1
#define MFN_DTYPE_F64 2

/* output activations of the MLP ops */
#define MFN_ACT_NONE 0
#define MFN_ACT_SIGMOID 1
#define MFN_ACT_EXP 2

int mfn_version(void);
const char* mfn_last_error(void);
/* compute capability of the current device as major*10+minor (100 on B200), -1 without a device */
int mfn_device_arch(void);

/* ---- measurement hooks (bench.py): kernel launches issued so far by this library, and CUDA-event bracketing of named
 * kernels (names: march_count, march_write, grid_encode_fwd, grid_encode_bwd, mlp_sigma_fwd, mlp_rgb_fwd, mlp_sigma_bwd,
 * mlp_rgb_bwd, mlp_fwd, mlp_bwd, composite_train_fw, composite_train_bw, adam, ...).  mfn_profile_set registers one
 * (name, start, stop) entry per call (up to 32; an empty name clears the table).  The events are recorded on the launch stream
 * and also work inside stream capture, so a kernel can be timed inside a replayed CUDA graph. */
int64_t mfn_launch_count(void);
void* mfn_event_create(void);
int mfn_event_destroy(void* event);
int mfn_event_elapsed_ms(void* start_event, void* stop_event, float* ms_host);
int mfn_profile_set(const char* kernel_name, void* start_event, void* stop_event);

/* ---- grid utilities -------------------------------------------------------------------------------- */
/* replaces vren.morton3D  (binding.cpp:47-51 -> raymarching.cu:62-87): coords (n,3) int32 -> indices (n) int32 */
int mfn_morton3d(const int32_t* coords, int64_t n, int32_t* indices, void* stream);
/* replaces vren.morton3D_invert (binding.cpp:54-58 -> raymarching.cu:90-119): indices (n) -> coords (n,3) */
int mfn_morton3d_invert(const int32_t* indices, int64_t n, int32_t* coords, void* stream);
/* replaces vren.packbits (binding.cpp:35-44 -> raymarching.cu:122-161): bit i of byte k = grid[8k+i] > thr.
 * grid holds 8*n_bytes elements of `dtype` (MFN_DTYPE_*). */
int mfn_packbits(const void* density_grid, int dtype, int64_t n_bytes, float density_threshold, uint8_t* density_bitfield, void* stream);

/* same, fp32 grid, threshold = min(*threshold_dev, max_threshold) read on the device: lets update_density_grid
 * (networks.py:268-271: packbits(grid, min(mean_density, density_threshold))) run without the .item() host sync */
int mfn_packbits_dev_thr(const float* density_grid, int64_t n_bytes, float max_threshold, const float* threshold_dev,
                         uint8_t* density_bitfield, void* stream);

/* ---- occupancy-grid update (NGP.update_density_grid, networks.py:157-197, 242-271) without host round trips ------------------------
 * positions of grid cells with uniform jitter inside the cell, xyzs (n,3):  cell_indices != NULL -> those morton indices;
 * random_cells != 0 -> n uniformly random cells (their morton indices go to indices_out); else cell i for i < n (all cells).
 * cascade c covers [-s, s]^3 with s = min(2^(c-1), scale) (l.253-260).  Counter-based RNG keyed by `seed`. */
int mfn_grid_cell_positions(const int32_t* cell_indices, int random_cells, int64_t n, int cascade, float scale, int grid_size, uint64_t seed,
                            int32_t* indices_out, float* xyzs, void* stream);
/* n draws from the occupied cells of one cascade (l.186-193), given the inclusive prefix count (int32, n_cells) of `grid > thr`:
 * uniform k in [0, total) -> index of the k-th occupied cell.  With no occupied cell the draws are uniform over all cells. */
int mfn_grid_draw_occupied(const int32_t* occupied_prefix_count, int64_t n_cells, int64_t n, uint64_t seed, int32_t* indices_out, void* stream);
/* one cascade: grid = grid < 0 ? grid : max(grid * decay, sigma at the queried cells) (l.263-266).  cell_indices == NULL: all
 * n == n_cells cells were queried in order; else `sigmas[i]` belongs to cell `cell_indices[i]`. */
int mfn_grid_update(float* density_grid_cascade, const int32_t* cell_indices, const float* sigmas, int64_t n_cells, int64_t n, float decay,
                    void* stream);
/* *mean_out = mean of the cells > 0 (l.268), device to device; scratch16 = 16 zero-initialised bytes the call leaves zeroed */
int mfn_grid_mean_positive(const float* density_grid, int64_t n, float* scratch16, float* mean_out, void* stream);
/* NGP.mark_invisible_cells (networks.py:199-240; called once before training, train.py:159-162): density_grid (cascades, G^3)
 * f32 in morton order <- 0 where the cell centre is seen by at least one camera at depth >= near_distance inside the image and by
 * none closer than that, else -1; count_grid (same shape, or NULL) <- fraction of the cameras covering the cell (self.count_grid).
 * K (3,3) row-major intrinsics and poses (n_cams,3,4) camera-to-world are DEVICE arrays, as in the reference's call. */
int mfn_grid_mark_invisible(const float* K, const float* poses, int32_t n_cams, int32_t img_w, int32_t img_h, int32_t cascades, float scale,
                            int32_t grid_size, float near_distance, float* density_grid, float* count_grid, void* stream);

/* ---- ray generation + batch sampling on the device ---------------------------------------------------
 * One launch replaces the DataLoader item (datasets/base.py:22-34), the pose / direction gathers of NeRFSystem.forward
 * (train.py:83-96) and get_ray_directions + get_rays (datasets/ray_utils.py:23-35, 60-68) for a data set resident in device memory:
 *   ray i: image m = img_idxs[i], pixel p = pix_idxs[i]  (u = p % width, v = p / width)
 *   direction = directions[p] if given, else ((u - cx + 0.5) / fx, (v - cy + 0.5) / fy, 1)
 *   rays_d[i] = direction @ poses[m][:, :3]^T,  rays_o[i] = poses[m][:, 3],  rgb[i] = pixels[m][p][0:3]
 * draw = 0: indices are read (img_idxs NULL -> every ray uses `image`; pix_idxs NULL -> ray i is pixel i: a whole test view);
 * draw = 1: 'all_images' sampling, image and pixel drawn per ray; draw = 2: 'same_image', one image per call (base.py:24-29).
 * Draws are counter-based on (seed, *call_counter, i); call_counter = uint64 in device memory or NULL, read only -- pass
 * (char*)march_workspace + 16 to get a fresh batch on every step without host involvement.  img_idxs_out / pix_idxs_out (int64, or
 * NULL) receive the indices used.  rgb / pixels may be NULL (no targets).  poses must be 16-byte aligned. */
typedef struct mfn_camera {
    float fx, fy, cx, cy;      /* K[0][0], K[1][1], K[0][2], K[1][2] */
    int32_t width, height;
} mfn_camera;
int mfn_ray_batch(const mfn_camera* camera_host, const float* directions, const float* poses, int32_t n_images, const float* pixels,
                  int32_t pixel_channels, const int64_t* img_idxs, const int64_t* pix_idxs, int32_t image, int32_t draw, uint64_t seed,
                  const void* call_counter, int64_t n_rays, float* rays_o, float* rays_d, float* rgb, int64_t* img_idxs_out,
                  int64_t* pix_idxs_out, void* stream);

/* ---- intersection ---------------------------------------------------------------------------------- */
/* replaces vren.ray_aabb_intersect (binding.cpp:4-16 -> intersection.cu:59-100).
 * hit_cnt (n_rays) int32, hits_t (n_rays,max_hits,2) f32 (-1 = no hit), hits_voxel_idx (n_rays,max_hits) int64.
 * Rows are sorted by t1 ascending exactly like the reference's sort (so -1 padding comes first). */
int mfn_ray_aabb_intersect(const float* rays_o, const float* rays_d, const float* centers, const float* half_sizes,
                           int64_t n_rays, int64_t n_voxels, int max_hits, int32_t* hit_cnt, float* hits_t,
                           int64_t* hits_voxel_idx, void* stream);
/* engine front end of a training step in one launch: scene-box slab test with max_hits = 1 (rendering.py:27-28), near clamp
 * (rendering.py:29) and the marcher's per-ray jitter (custom_functions.py:83 torch.rand_like), counter-based.  hits_t (n_rays,2);
 * noise (n_rays) or NULL; call_counter: uint64 in device memory read (not written) by this call -- pass
 * (char*)march_workspace + 16, which mfn_raymarching_train increments once per call, to get fresh jitter on every graph replay. */
int mfn_ray_setup(const float* rays_o, const float* rays_d, const float* center_host, const float* half_size_host, int64_t n_rays,
                  float near_distance, void* call_counter, float* hits_t, float* noise, void* stream);
/* replaces vren.ray_sphere_intersect (binding.cpp:19-32 -> intersection.cu:156-197) */
int mfn_ray_sphere_intersect(const float* rays_o, const float* rays_d, const float* centers, const float* radii,
                             int64_t n_rays, int64_t n_spheres, int max_hits, int32_t* hit_cnt, float* hits_t,
                             int64_t* hits_sphere_idx, void* stream);

/* ---- ray marching ---------------------------------------------------------------------------------- */
/* replaces vren.raymarching_train (binding.cpp:60-81 -> raymarching.cu:166-332), split so the caller can size
 * the sample arrays exactly: _count fills rays_a (n_rays,3) int64 and counter (2) int32 = {total samples,
 * n_rays}; _write then emits exactly the samples (rows < capacity).  mfn_raymarching_train = both in two launches; when
 * `capacity` is below the total it truncates the overflowing rays consistently (rays_a, counter[0] <= capacity).
 * Workspace header (first 256 bytes, zero it once): u64 @8 += samples marched by each call, u64 @16 += 1 per call. */
int64_t mfn_march_train_workspace_bytes(int64_t n_rays, int max_samples);
int mfn_march_train_count(const float* rays_o, const float* rays_d, const float* hits_t, const uint8_t* density_bitfield,
                          int cascades, float scale, float exp_step_factor, const float* noise, int grid_size, int max_samples,
                          int64_t n_rays, int64_t* rays_a, int32_t* counter, void* workspace, int64_t workspace_bytes, void* stream);
int mfn_march_train_write(const float* rays_o, const float* rays_d, const int64_t* rays_a, const void* workspace, int max_samples,
                          int64_t n_rays, int64_t capacity, float* xyzs, float* dirs, float* deltas, float* ts, void* stream);
int mfn_raymarching_train(const float* rays_o, const float* rays_d, const float* hits_t, const uint8_t* density_bitfield,
                          int cascades, float scale, float exp_step_factor, const float* noise, int grid_size, int max_samples,
                          int64_t n_rays, int64_t capacity, int64_t* rays_a, float* xyzs, float* dirs, float* deltas, float* ts,
                          int32_t* counter, void* workspace, int64_t workspace_bytes, void* stream);
/* replaces vren.raymarching_test (binding.cpp:84-107 -> raymarching.cu:335-454).  hits_t (n_rays_total,2) is
 * advanced in place; xyzs/dirs (n_alive,n_samples,3), deltas/ts (n_alive,n_samples) are fully written
 * (zero padded); n_eff_samples (n_alive) int32. */
int mfn_raymarching_test(const float* rays_o, const float* rays_d, float* hits_t, const int64_t* alive_indices,
                         const uint8_t* density_bitfield, int cascades, float scale, float exp_step_factor, int grid_size,
                         int max_samples, int n_samples, int64_t n_alive, float* xyzs, float* dirs, float* deltas, float* ts,
                         int32_t* n_eff_samples, void* stream);

/* ---- compositing ----------------------------------------------------------------------------------- */
/* replaces vren.composite_train_fw (binding.cpp:110-127 -> volumerendering.cu:6-84).  All outputs fully written. */
int mfn_composite_train_fw(const float* sigmas, const float* rgbs, const float* deltas, const float* ts, const int64_t* rays_a,
                           float T_threshold, int64_t n_rays, int64_t n_samples, int64_t* total_samples, float* opacity,
                           float* depth, float* rgb, float* ws, void* stream);
/* replaces vren.composite_train_bw (binding.cpp:130-167 -> volumerendering.cu:87-202) */
int mfn_composite_train_bw(const float* dL_dopacity, const float* dL_ddepth, const float* dL_drgb, const float* dL_dws,
                           const float* sigmas, const float* rgbs, const float* ws, const float* deltas, const float* ts,
                           const int64_t* rays_a, const float* opacity, const float* depth, const float* rgb, float T_threshold,
                           int64_t n_rays, int64_t n_samples, float* dL_dsigmas, float* dL_drgbs, void* stream);
/* replaces vren.composite_test_fw (binding.cpp:170-198 -> volumerendering.cu:205-285).  opacity/depth/rgb and
 * alive_indices are updated in place; sigmas/deltas/ts (n_alive,n_samples), rgbs (n_alive,n_samples,3). */
int mfn_composite_test_fw(const float* sigmas, const float* rgbs, const float* deltas, const float* ts, int64_t* alive_indices,
                          float T_threshold, const int32_t* n_eff_samples, int n_samples, int64_t n_alive, float* opacity,
                          float* depth, float* rgb, void* stream);

/* ---- distortion loss ------------------------------------------------------------------------------- */
/* replaces vren.distortion_loss_fw (binding.cpp:201-213 -> losses.cu:64-109) */
int mfn_distortion_loss_fw(const float* ws, const float* deltas, const float* ts, const int64_t* rays_a, int64_t n_rays,
                           int64_t n_samples, float* loss, float* ws_inclusive_scan, float* wts_inclusive_scan, void* stream);
/* replaces vren.distortion_loss_bw (binding.cpp:216-231 -> losses.cu:145-175) */
int mfn_distortion_loss_bw(const float* dL_dloss, const float* ws_inclusive_scan, const float* wts_inclusive_scan, const float* ws,
                           const float* deltas, const float* ts, const int64_t* rays_a, int64_t n_rays, int64_t n_samples,
                           float* dL_dws, void* stream);

/* ---- field: encodings (replace tcnn's HashGrid / SphericalHarmonics encodings, networks.py:36-47,60-67) ------- */
#define MFN_GRID_HASH 0
/* MixedFeature grid of the MF-NeRF fork (networks.py:40-46 `--grid MixedFeature --N_tables K`).  The fork's tiny-cuda-nn is not part
 * of the reference tree, so the semantics are DEFINED here (parity unpinned): level l lives in table k = l*K/L; the K tables have 2^T
 * entries each and are always hashed; a vertex is hashed by its coordinates in the table's canonical grid (the finest level of the
 * table): canonical vertex = round((v - 0.5) * scale_canonical / scale_l + 0.5).  Runs on the unfused field kernels. */
#define MFN_GRID_MIXED 1
typedef struct mfn_grid_cfg {
    int32_t n_levels;            /* L */
    int32_t n_features;          /* F: 1, 2, 4 or 8 */
    int32_t log2_hashmap_size;   /* T */
    int32_t base_resolution;     /* N_min */
    double per_level_scale;      /* b */
    int32_t grid_type;           /* MFN_GRID_* */
    int32_t n_tables;            /* K of the MixedFeature grid (1 <= K <= L); ignored by the plain hash grid */
} mfn_grid_cfg;
/* HOST helper: level table.  offsets_host[L+1] (entries), resolutions_host[L], scales_host[L] may be NULL.
 * Returns the total number of table entries (parameters = entries * F), <0 on error. */
int64_t mfn_grid_layout(const mfn_grid_cfg* cfg_host, uint32_t* offsets_host, uint32_t* resolutions_host, float* scales_host);
/* x01 (n,3) f32 in [0,1]; table fp16 (entries*F); out (n, L*F) fp16, feature order [level][feature] */
int mfn_grid_encode_fwd(const float* x01, const void* table, const mfn_grid_cfg* cfg_host, int64_t n, void* out, void* stream);
/* dL_dout (n, L*F) fp16; dgrid fp32 (entries*F), ACCUMULATED into with atomics (caller zeroes it) */
int mfn_grid_encode_bwd(const float* x01, const void* dL_dout, const mfn_grid_cfg* cfg_host, int64_t n, float* dgrid, void* stream);
/* gradient w.r.t. the sample position (the reference's --optimize_ext path, custom_functions.py:102-112 / train.py:91,122,138; tcnn
 * kernel_grid_backward_input): dL_dout (n, L*F) fp16 -> dx01 (n,3) f32 = sum_l dL_dout_l . d feat_l / d x01 (same scale as dL_dout) */
int mfn_grid_encode_bwd_input(const float* x01, const void* table, const void* dL_dout, const mfn_grid_cfg* cfg_host, int64_t n, float* dx01,
                              void* stream);
/* gradient of the SH encoding w.r.t. dirs01: dL_dout (n,16) fp16 -> ddirs01 (n,3) f32 */
int mfn_sh4_bwd(const float* dirs01, const void* dL_dout, int64_t n, float* ddirs01, void* stream);
/* degree-4 spherical harmonics of 2*d01-1: (n,3) f32 -> 16 fp16 values written at out[i*out_stride + out_offset ...] */
int mfn_sh4_fwd(const float* dirs01, int64_t n, void* out, int out_stride, int out_offset, void* stream);

/* ---- field: fully fused MLPs (replaces tcnn.Network / the MLP half of tcnn.NetworkWithInputEncoding,
 * networks.py:36-57,69-79).  fp16 activations (n,in_dim) -> (n,16) (output padded to 16 like tcnn), fp16 weights laid out
 * first->last as row-major (out x in) matrices: [width x in_dim][(n_hidden-1) x width x width][16 x width], no biases,
 * ReLU hidden activations, `out_act` = MFN_ACT_*.  `acts` (n_hidden, n, width) fp16 receives the hidden activations
 * when non-NULL (needed by mfn_mlp_bwd).  in_dim in {16,32,64}, width in {64,128}, n_hidden in {1,2}. */
int64_t mfn_mlp_param_count(int in_dim, int width, int n_hidden);
int mfn_mlp_fwd(const void* in, const void* weights, int in_dim, int width, int n_hidden, int out_act, int64_t n, void* out,
                void* acts, void* stream);
/* dL_dout (n,16) fp16 (already multiplied by the caller's loss scale); dL_din (n,in_dim) fp16 or NULL; dW fp32, same layout as
 * `weights`, ACCUMULATED into (caller zeroes it). */
int mfn_mlp_bwd(const void* dL_dout, const void* in, const void* acts, const void* out, const void* weights, int in_dim, int width,
                int n_hidden, int out_act, int64_t n, void* dL_din, float* dW, void* stream);

/* ---- field: the whole NGP field as one op pair (replaces NGP.density / NGP.forward, networks.py:96-155, and
 * TruncExp, custom_functions.py:162-173).  xyz_params_h = fp16 copy of `xyz_encoder.params` ([3072 MLP | grid]),
 * rgb_params_h = fp16 copy of `rgb_net.params`.  The sample count is min(*n_dev, n_max) when n_dev != NULL (device
 * memory), else n_max: no host synchronisation is needed between the marcher and the field. */
typedef struct mfn_field_cfg {
    mfn_grid_cfg grid;
    int32_t sigma_width, sigma_hidden;   /* 64, 1 (networks.py:48-57) */
    int32_t rgb_width, rgb_hidden;       /* hparams.rgb_channels, hparams.rgb_layers */
    int32_t rgb_act;                     /* MFN_ACT_SIGMOID, or MFN_ACT_NONE with --use_exposure */
    float xyz_min[3], xyz_max[3];        /* scene box: x01 = (x - xyz_min) / (xyz_max - xyz_min) */
} mfn_field_cfg;
int64_t mfn_field_workspace_bytes(const mfn_field_cfg* cfg_host, int64_t n_max, int training);
/* 1 when this shape runs on the fused tcgen05 kernels (field_fused.cu), 0 when it runs on the unfused pipeline, <0 on a bad config.
 * The fused backward reads nothing but the workspace, dL_dsigmas and dL_drgbs: `xyzs` may already be overwritten when it runs. */
int mfn_field_is_fused(const mfn_field_cfg* cfg_host);
/* Fused shapes, training workspace: address (inside `workspace`) of the int32 where mfn_field_fwd leaves the sample count it used,
 * min(*n_dev, n_max).  Pass it as `n_dev` to mfn_field_bwd when the caller's own counter may be overwritten in between (the engine
 * starts marching the next batch while the backward pass of this one is still running).  NULL for shapes on the unfused pipeline. */
void* mfn_field_count_ptr(const mfn_field_cfg* cfg_host, void* workspace, int64_t n_max);
/* xyzs, dirs (n,3) f32 -> sigmas (n) f32, rgbs (n,3) f32 (values rounded to fp16 like tcnn's output) */
int mfn_field_fwd(const mfn_field_cfg* cfg_host, const void* xyz_params_h, const void* rgb_params_h, const float* xyzs, const float* dirs,
                  int64_t n_max, const int32_t* n_dev, float* sigmas, float* rgbs, void* workspace, int64_t workspace_bytes, void* stream);
/* needs the workspace of the matching mfn_field_fwd call (training=1 size).  d_*_params: fp32, same layout as the
 * parameters, ACCUMULATED into and left multiplied by loss_scale; *overflow_flag is set to 1 if an fp16 gradient overflowed. */
int mfn_field_bwd(const mfn_field_cfg* cfg_host, const void* xyz_params_h, const void* rgb_params_h, const float* xyzs, int64_t n_max,
                  const int32_t* n_dev, const float* dL_dsigmas, const float* dL_drgbs, float loss_scale, float* d_xyz_params,
                  float* d_rgb_params, int32_t* overflow_flag, void* workspace, int64_t workspace_bytes, void* stream);
/* tcnn.NetworkWithInputEncoding.forward for the fused shape (networks.py:36-57, called at :107): xyzs (n,3) f32 ->
 * h_out (n,16) fp16, the raw (un-activated) outputs of grid encoding + 32->64->16 network, with x01 = (x - xyz_min) / (xyz_max -
 * xyz_min) as in the other field calls (pass 0 / 1 for positions that are already in [0,1]).  Inference only: nothing is saved for a
 * backward pass.  Returns an error for shapes mfn_field_is_fused() reports 0 for. */
int mfn_geo_fwd(const mfn_field_cfg* cfg_host, const void* xyz_params_h, const float* xyzs, int64_t n_max, const int32_t* n_dev,
                void* h_out, void* stream);
/* density only (NGP.density, used by update_density_grid, networks.py:258) */
int mfn_density_fwd(const mfn_field_cfg* cfg_host, const void* xyz_params_h, const float* xyzs, int64_t n_max, const int32_t* n_dev,
                    float* sigmas, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- test-time rendering without host round trips (replaces the while-loop of __render_rays_test, rendering.py:46-118) ----------
 * The alive list, its length, N_samples = clamp(N_rays / N_alive, min_samples, 64) and the max_samples budget live in the
 * workspace; one iteration = plan -> march (raymarching_test semantics) -> field -> composite (composite_test_fw semantics)
 * + compaction of the surviving rays.  Call _begin once (AABB test + near clamp, rendering.py:27-29, outputs zeroed), then
 * _iterations in batches until mfn_render_status reports no alive rays, then _finish (rgb += bg * (1 - opacity), l.112-116).
 * center_host / half_size_host / bg_rgb_host: 3 floats in HOST memory.  Needs the field shape the fused kernels cover. */
int64_t mfn_render_workspace_bytes(int64_t n_rays, int min_samples);
int mfn_render_begin(const float* rays_o, const float* rays_d, const float* center_host, const float* half_size_host, int64_t n_rays,
                     float near_distance, int min_samples, int max_samples, float* opacity, float* depth, float* rgb, void* workspace,
                     int64_t workspace_bytes, void* stream);
int mfn_render_iterations(const mfn_field_cfg* cfg_host, const void* xyz_params_h, const void* rgb_params_h, const float* rays_o, const float* rays_d,
                          int64_t n_rays, const uint8_t* density_bitfield, int cascades, float scale, float exp_step_factor, int grid_size,
                          int max_samples, int min_samples, float T_threshold, int n_iterations, float* opacity, float* depth, float* rgb,
                          void* workspace, int64_t workspace_bytes, void* stream);
/* asynchronously copies 10 int32 {n_alive[0], n_alive[1], current list, N_samples, rows, samples spent, iterations, field rows / 1024, total N_eff lo, hi}
 * to HOST (pinned) memory */
int mfn_render_status(const void* workspace, int32_t* status_host_pinned, void* stream);
int mfn_render_finish(float* rgb, const float* opacity, const float* bg_rgb_host, int64_t n_rays, void* stream);

/* ---- per-ray loss (losses.py:47-60 NeRFLoss + train.py:178 + background blend rendering.py:153-161) --------------- */
/* loss = mean((rgb + bg*(1-opacity) - target)^2) + mean(lambda_o * -(o+1e-10)*log(o+1e-10)) [+ mean(lambda_d * distortion)].
 * Writes the gradients w.r.t. rgb (n,3), opacity (n) and, if distortion != NULL, the per-ray distortion loss (n), each
 * multiplied by grad_scale; loss_out (3 floats, device, ACCUMULATED) receives the three terms; rgb_final may be NULL.
 * bg_rgb_host: 3 floats in HOST memory. */
int mfn_nerf_loss_fwbw(const float* rgb, const float* opacity, const float* target, const float* distortion, int64_t n_rays,
                       const float* bg_rgb_host, float lambda_opacity, float lambda_distortion, float grad_scale, float* dL_drgb,
                       float* dL_dopacity, float* dL_ddistortion, float* rgb_final, float* loss_out, void* stream);
/* mfn_composite_train_fw + mfn_nerf_loss_fwbw (without distortion term) + mfn_composite_train_bw in ONE launch: the engine's training step
 * when distortion_loss_w == 0 (VolumeRenderer.forward custom_functions.py:137-146 -> NeRFLoss losses.py:47-60 with the background blend of
 * rendering.py:153-161 -> VolumeRenderer.backward :148-159).  Outputs as in the three calls (dL_drgb (n,3), dL_dopacity (n), rgb_final may be
 * NULL); loss_out (3 floats, or NULL) is WRITTEN, not accumulated: {rgb term, opacity term, 0}.  scratch16: 16 zero-initialised bytes the
 * call leaves zeroed.  clear_flag (or NULL): an int32 set to 0 at the end of the call -- the backward pass's overflow flag, cleared here so
 * that the step needs no separate memset. */
int mfn_composite_loss_train(const float* sigmas, const float* rgbs, const float* deltas, const float* ts, const int64_t* rays_a,
                             const float* target, float T_threshold, int64_t n_rays, int64_t n_samples, const float* bg_rgb_host,
                             float lambda_opacity, float grad_scale, int64_t* total_samples, float* opacity, float* depth, float* rgb,
                             float* ws, float* rgb_final, float* dL_drgb, float* dL_dopacity, float* dL_dsigmas, float* dL_drgbs,
                             float* loss_out, float* scratch16, int32_t* clear_flag, void* stream);
/* replaces torch_scatter.segment_csr(src, indptr) (sum) in RayMarcher.backward (custom_functions.py:102-112):
 * src (rows, width) f32, indptr (n_segments+1) int64 -> out (n_segments, width) f32 */
int mfn_segment_sum(const float* src, const int64_t* indptr, int64_t n_segments, int width, float* out, void* stream);
/* rendering.py:29: hits_t (n,2), t1 in [0, near) -> near */
int mfn_clamp_near(float* hits_t, int64_t n_rays, float near_distance, void* stream);

/* ---- optimiser (train.py:136 FusedAdam(eps=1e-15)) ---------------------------------------- */
/* p -= lr * mhat / (sqrt(vhat) + eps) with g = grads * grad_scale; skipped when *skip_flag != 0; params_h (fp16 shadow,
 * may be NULL) refreshed; grads zeroed afterwards when zero_grad != 0.  `step` is 1-based. */
int mfn_adam_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, void* params_h, int64_t n, float lr, float beta1,
                  float beta2, float eps, int step, float grad_scale, const int32_t* skip_flag, int zero_grad, void* stream);
/* the same step with its per-step scalars in DEVICE memory, hyper_dev = {lr, 1 - beta1^step, 1 - beta2^step} (fill a host copy with
 * mfn_adam_hyper and upload it ahead of time): nothing in the launch changes from step to step, so it can be part of a CUDA graph. */
int mfn_adam_hyper(float lr, float beta1, float beta2, int step, float* hyper_host);
/* ---- AMP bookkeeping (the reference trains under Lightning precision=16: autocast + torch.amp.GradScaler, train.py:284-295) -------
 * amp_state = 8 floats in DEVICE memory {loss scale, growth tracker, skipped steps (total), applied steps (total), 1 - beta1^t and
 * 1 - beta2^t of the next step, 2 spare}, initialised by mfn_amp_init (pageable host copy of 32 bytes + one tiny launch).  mfn_field_bwd_amp scales the gradients by amp_state[0]; mfn_adam_step_amp unscales by
 * inv_world / amp_state[0] (inv_world = 1 / number of data-parallel ranks whose gradients were summed), takes lr from lr_dev[0] and
 * the bias corrections of t = amp_state[3] + 1 from amp_state[4..5] -- a skipped step does not advance Adam's t; mfn_amp_update then applies GradScaler's
 * update rule from the overflow flag (skip -> scale *= backoff; `growth_interval` consecutive applied steps -> scale *= growth).
 * Nothing in these launches changes from step to step, so all three can sit in a CUDA graph. */
int mfn_field_bwd_amp(const mfn_field_cfg* cfg_host, const void* xyz_params_h, const void* rgb_params_h, const float* xyzs, int64_t n_max,
                      const int32_t* n_dev, const float* dL_dsigmas, const float* dL_drgbs, const float* amp_state, float* d_xyz_params,
                      float* d_rgb_params, int32_t* overflow_flag, void* workspace, int64_t workspace_bytes, void* stream);
int mfn_adam_step_amp(float* params, float* grads, float* exp_avg, float* exp_avg_sq, void* params_h, int64_t n, const float* lr_dev,
                      float beta1, float beta2, float eps, float inv_world, const float* amp_state, const int32_t* skip_flag, int zero_grad,
                      void* stream);
int mfn_amp_init(float* amp_state, float loss_scale, int applied_steps, float beta1, float beta2, void* stream);
int mfn_amp_update(float* amp_state, const int32_t* skip_flag, float backoff, float growth, int growth_interval, float min_scale,
                   float max_scale, float beta1, float beta2, void* stream);
int mfn_adam_step_dev(float* params, float* grads, float* exp_avg, float* exp_avg_sq, void* params_h, int64_t n, const float* hyper_dev,
                      float beta1, float beta2, float eps, float grad_scale, const int32_t* skip_flag, int zero_grad, void* stream);
/* ---- data-parallel gradient exchange + optimiser in ONE kernel over NVLink / NVSwitch peer memory (replaces DDP's bucketed NCCL
 * all-reduce + replicated optimiser, train.py:284-285; csrc/dp_exchange.cu).  Every rank owns the shard [shard_begin, shard_begin + n)
 * of the flat parameter vector: its kernel sums that shard of ALL ranks' gradient buffers (multimem.ld_reduce through the NVSwitch
 * when grads_mc != 0, else `world` peer loads), applies Adam (AMP state as in mfn_adam_step_amp, inv_world =
 * 1 / world) to its fp32 master shard and stores the fp16 shadow into EVERY rank's copy (multimem.st / peer stores).
 * *_ptrs_host: HOST arrays of `world` peer-mapped device addresses (index = rank) of each rank's gradient buffer (f32), fp16 shadow and
 * int32 overflow flag; *_mc: multicast addresses of the same two buffers or 0; skip_out (device, may be NULL) receives the OR of all
 * ranks' flags (set: nothing is updated).  The caller must barrier all ranks before (gradients complete) and after (shadows written)
 * the launch, and clears its own gradient buffer after the second barrier. */
int mfn_dp_exchange_adam(int world, const uint64_t* grads_ptrs_host, const uint64_t* shadow_ptrs_host, const uint64_t* flag_ptrs_host,
                         uint64_t grads_mc, uint64_t shadow_mc, float* params_shard, float* exp_avg_shard, float* exp_avg_sq_shard,
                         int64_t shard_begin, int64_t n, const float* lr_dev, float beta1, float beta2, float eps, const float* amp_state,
                         int32_t* skip_out, void* stream);
int mfn_cast_f32_to_f16(const float* src, void* dst, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MFNERF_B200_H_ */
