/* vren_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, CPU restatement of the reference's `vren` CUDA kernels (lly00412/MF-NeRF, models/csrc/ *.cu files).
 * It exists to CHECK the sm_100a kernels in mf-nerf_b200/csrc; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product never does.
 *
 * Parity pinning: the reference ships no tests or golden vectors, so this restatement is
 * pinned against outputs of the reference's OWN kernels -- oracle/_ref/vren_ref*.so, built from
 * /root/reference/models/csrc by oracle/build_ref_vren.sh -- run on a B200 on seeded inputs and committed
 * as tests/golden/vren_ref_*.npz by tests/golden/make_golden_vren.py.
 *
 * Floating point: compile with -ffp-contract=off.  Where the reference's SASS (nvcc -O2, -fmad=true) fuses a
 * multiply-add the code below says fmaf() explicitly; everything else is a separately rounded IEEE op, so the
 * marcher reproduces the GPU's N_samples / ts / deltas / xyzs bit for bit.  The compositor uses libm expf()
 * where the GPU uses the ex2.approx-based __expf(), so those outputs agree to ~1e-6 relative, not bitwise.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SQRT3 1.73205080757f

/* ---- morton (ref: raymarching.cu:35-60) ---- */
static inline uint32_t expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
static inline uint32_t morton3D(uint32_t x, uint32_t y, uint32_t z) { return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2); }
static inline uint32_t morton3D_invert(uint32_t x) {
    x = x & 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}

/* ref: raymarching.cu:62-70 */
void orc_morton3d(const int32_t* coords, int64_t n, int32_t* out) {
    for (int64_t i = 0; i < n; ++i) out[i] = (int32_t)morton3D((uint32_t)coords[3 * i], (uint32_t)coords[3 * i + 1], (uint32_t)coords[3 * i + 2]);
}
/* ref: raymarching.cu:90-101 (signed shift of the int index, then uint conversion) */
void orc_morton3d_invert(const int32_t* idx, int64_t n, int32_t* coords) {
    for (int64_t i = 0; i < n; ++i) {
        const int32_t v = idx[i];
        coords[3 * i + 0] = (int32_t)morton3D_invert((uint32_t)(v >> 0));
        coords[3 * i + 1] = (int32_t)morton3D_invert((uint32_t)(v >> 1));
        coords[3 * i + 2] = (int32_t)morton3D_invert((uint32_t)(v >> 2));
    }
}
/* ref: raymarching.cu:122-141 */
void orc_packbits_f32(const float* grid, int64_t n_bytes, float thr, uint8_t* bits) {
    for (int64_t n = 0; n < n_bytes; ++n) {
        uint8_t b = 0;
        for (int i = 0; i < 8; ++i) b |= (grid[8 * n + i] > thr) ? (uint8_t)(1u << i) : 0;
        bits[n] = b;
    }
}

/* ---- ray / AABB (ref: intersection.cu:5-22, 25-56, 59-100) ----
 * The reference keeps the first max_hits hits in atomic arrival order and then sorts each (-1 padded) row by
 * t1 ascending.  Arrival order is not defined; voxel order is used here (and by the CUDA kernel under test). */
void orc_ray_aabb_intersect(const float* o, const float* d, const float* centers, const float* half, int64_t n_rays, int64_t n_vox,
                            int max_hits, int32_t* hit_cnt, float* hits_t, int64_t* hits_idx) {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n_rays; ++r) {
        const float ix = 1.0f / d[3 * r], iy = 1.0f / d[3 * r + 1], iz = 1.0f / d[3 * r + 2];
        float* ht = hits_t + r * (int64_t)max_hits * 2;
        int64_t* hi = hits_idx + r * (int64_t)max_hits;
        for (int k = 0; k < max_hits; ++k) { ht[2 * k] = -1.f; ht[2 * k + 1] = -1.f; hi[k] = -1; }
        int cnt = 0;
        for (int64_t v = 0; v < n_vox; ++v) {
            const float tminx = ((centers[3 * v] - half[3 * v]) - o[3 * r]) * ix, tmaxx = ((centers[3 * v] + half[3 * v]) - o[3 * r]) * ix;
            const float tminy = ((centers[3 * v + 1] - half[3 * v + 1]) - o[3 * r + 1]) * iy, tmaxy = ((centers[3 * v + 1] + half[3 * v + 1]) - o[3 * r + 1]) * iy;
            const float tminz = ((centers[3 * v + 2] - half[3 * v + 2]) - o[3 * r + 2]) * iz, tmaxz = ((centers[3 * v + 2] + half[3 * v + 2]) - o[3 * r + 2]) * iz;
            float t1 = fmaxf(fmaxf(fminf(tminx, tmaxx), fminf(tminy, tmaxy)), fminf(tminz, tmaxz));
            float t2 = fminf(fminf(fmaxf(tminx, tmaxx), fmaxf(tminy, tmaxy)), fmaxf(tminz, tmaxz));
            if (t1 > t2) { t1 = -1.f; t2 = -1.f; }
            if (t2 > 0) {
                if (cnt < max_hits) { ht[2 * cnt] = fmaxf(t1, 0.0f); ht[2 * cnt + 1] = t2; hi[cnt] = v; }
                ++cnt;
            }
        }
        hit_cnt[r] = cnt;
        /* stable ascending sort of the whole row by t1 (insertion sort; rows are tiny) */
        for (int a = 1; a < max_hits; ++a) {
            const float k0 = ht[2 * a], k1 = ht[2 * a + 1]; const int64_t ki = hi[a];
            int b = a - 1;
            while (b >= 0 && ht[2 * b] > k0) { ht[2 * (b + 1)] = ht[2 * b]; ht[2 * (b + 1) + 1] = ht[2 * b + 1]; hi[b + 1] = hi[b]; --b; }
            ht[2 * (b + 1)] = k0; ht[2 * (b + 1) + 1] = k1; hi[b + 1] = ki;
        }
    }
}

/* ---- marching helpers (ref: raymarching.cu:7-32) ---- */
typedef struct { float dt_min, dt_max, esf, gs, gs_inv, gsm1, scale; int cascades; uint32_t g3; } march_const;

static march_const make_const(int cascades, int grid_size, float scale, float esf, int max_samples, float dt_scale) {
    march_const c;
    c.esf = esf;
    c.dt_min = SQRT3 / (float)max_samples;                 /* SQRT3/max_samples                       */
    c.dt_max = (dt_scale * 3.4641015529632568359f) / (float)grid_size; /* SQRT3*2*scale/grid_size, SQRT3*2 folded exactly */
    c.gs = (float)grid_size; c.gs_inv = 1.0f / c.gs; c.gsm1 = c.gs + -1.0f;
    c.scale = scale; c.cascades = cascades; c.g3 = (uint32_t)grid_size * grid_size * grid_size;
    return c;
}
/* clamp(t*esf, lo, hi) = fmaxf(lo, fminf(f, hi)) (helper_math.h clamp) */
static inline float calc_dt(float t, const march_const* c) { return fmaxf(fminf(t * c->esf, c->dt_max), c->dt_min); }
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

typedef struct { float x, y, z, dt, t_target; int occ; } probe_t;

/* one loop body of raymarching.cu:205-228 */
static inline probe_t probe(float t, const float* o, const float* d, const float* inv, const march_const* c, const uint8_t* bits) {
    probe_t p;
    p.x = fmaf(d[0], t, o[0]); p.y = fmaf(d[1], t, o[1]); p.z = fmaf(d[2], t, o[2]);
    p.dt = calc_dt(t, c);
    int e_pos, e_dt;
    (void)frexpf(fmaxf(fabsf(p.x), fmaxf(fabsf(p.y), fabsf(p.z))), &e_pos);   /* mip_from_pos, l.19-23 */
    (void)frexpf(p.dt * c->gs, &e_dt);                                          /* mip_from_dt,  l.29-32 */
    const int mip = imax(imin(c->cascades - 1, imax(0, e_pos + 1)), imin(c->cascades - 1, imax(0, e_dt)));
    const float bound = fminf(scalbnf(1.0f, mip - 1), c->scale);
    const float bound_inv = 1.0f / bound;
    const int nx = (int)fmaxf(0.0f, fminf((fmaf(p.x, bound_inv, 1.0f) * 0.5f) * c->gs, c->gsm1));
    const int ny = (int)fmaxf(0.0f, fminf((fmaf(p.y, bound_inv, 1.0f) * 0.5f) * c->gs, c->gsm1));
    const int nz = (int)fmaxf(0.0f, fminf((fmaf(p.z, bound_inv, 1.0f) * 0.5f) * c->gs, c->gsm1));
    const uint32_t idx = (uint32_t)mip * c->g3 + morton3D((uint32_t)nx, (uint32_t)ny, (uint32_t)nz);
    p.occ = (bits[idx / 8] >> (idx % 8)) & 1;
    const float sx = copysignf(1.0f, d[0]), sy = copysignf(1.0f, d[1]), sz = copysignf(1.0f, d[2]);
    const float tx = fmaf(bound, fmaf((fmaf(sx, 0.5f, (float)nx + 0.5f)) * c->gs_inv, 2.0f, -1.0f), -p.x) * inv[0];
    const float ty = fmaf(bound, fmaf((fmaf(sy, 0.5f, (float)ny + 0.5f)) * c->gs_inv, 2.0f, -1.0f), -p.y) * inv[1];
    const float tz = fmaf(bound, fmaf((fmaf(sz, 0.5f, (float)nz + 0.5f)) * c->gs_inv, 2.0f, -1.0f), -p.z) * inv[2];
    p.t_target = t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
    return p;
}

/* ref: raymarching.cu:166-280.  Pass 1 counts, an exclusive prefix sum replaces the atomic allocation
 * (row i of rays_a = ray i), pass 2 writes.  Outputs xyzs/dirs/deltas/ts must hold `capacity` rows. */
int64_t orc_raymarching_train(const float* rays_o, const float* rays_d, const float* hits_t, const uint8_t* bits, int cascades, float scale,
                              float esf, const float* noise, int grid_size, int max_samples, int64_t n_rays, int64_t capacity,
                              int64_t* rays_a, float* xyzs, float* dirs, float* deltas, float* ts) {
    const march_const c = make_const(cascades, grid_size, scale, esf, max_samples, scale);
    int32_t* counts = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_rays > 0 ? n_rays : 1));
    float* t_first = (float*)malloc(sizeof(float) * (size_t)(n_rays > 0 ? n_rays : 1));
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t r = 0; r < n_rays; ++r) {
        const float* o = rays_o + 3 * r; const float* d = rays_d + 3 * r;
        const float inv[3] = {1.0f / d[0], 1.0f / d[1], 1.0f / d[2]};
        float t1 = hits_t[2 * r]; const float t2 = hits_t[2 * r + 1];
        if (t1 >= 0) t1 = fmaf(calc_dt(t1, &c), noise[r], t1);   /* l.195-198 */
        t_first[r] = t1;
        float t = t1; int n = 0;
        while (0 <= t && t < t2 && n < max_samples) {             /* l.204 */
            const probe_t p = probe(t, o, d, inv, &c, bits);
            if (p.occ) { t = t + p.dt; ++n; }
            else { do { t = calc_dt(t, &c) + t; } while (t < p.t_target); }
        }
        counts[r] = n;
    }
    int64_t total = 0;
    for (int64_t r = 0; r < n_rays; ++r) { rays_a[3 * r] = r; rays_a[3 * r + 1] = total; rays_a[3 * r + 2] = counts[r]; total += counts[r]; }
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t r = 0; r < n_rays; ++r) {
        const float* o = rays_o + 3 * r; const float* d = rays_d + 3 * r;
        const float inv[3] = {1.0f / d[0], 1.0f / d[1], 1.0f / d[2]};
        const float t2 = hits_t[2 * r + 1];
        const int64_t start = rays_a[3 * r + 1];
        float t = t_first[r]; int s = 0; const int n = counts[r];
        while (t < t2 && s < n) {                                   /* l.245 */
            const probe_t p = probe(t, o, d, inv, &c, bits);
            if (p.occ) {
                const int64_t k = start + s;
                if (k < capacity) {
                    xyzs[3 * k] = p.x; xyzs[3 * k + 1] = p.y; xyzs[3 * k + 2] = p.z;
                    dirs[3 * k] = d[0]; dirs[3 * k + 1] = d[1]; dirs[3 * k + 2] = d[2];
                    ts[k] = t; deltas[k] = p.dt;
                }
                t = t + p.dt; ++s;
            } else { do { t = calc_dt(t, &c) + t; } while (t < p.t_target); }
        }
    }
    free(counts); free(t_first);
    return total;
}

/* ref: raymarching.cu:335-404.  NOTE calc_dt gets (float)cascades where the train marcher passes scale. */
void orc_raymarching_test(const float* rays_o, const float* rays_d, float* hits_t, const int64_t* alive, const uint8_t* bits, int cascades,
                          float scale, float esf, int grid_size, int max_samples, int n_samples, int64_t n_alive, float* xyzs, float* dirs,
                          float* deltas, float* ts, int32_t* n_eff) {
    const march_const c = make_const(cascades, grid_size, scale, esf, max_samples, (float)cascades);
    memset(xyzs, 0, sizeof(float) * 3 * (size_t)n_alive * n_samples); memset(dirs, 0, sizeof(float) * 3 * (size_t)n_alive * n_samples);
    memset(deltas, 0, sizeof(float) * (size_t)n_alive * n_samples); memset(ts, 0, sizeof(float) * (size_t)n_alive * n_samples);
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t n = 0; n < n_alive; ++n) {
        const int64_t r = alive[n];
        const float* o = rays_o + 3 * r; const float* d = rays_d + 3 * r;
        const float inv[3] = {1.0f / d[0], 1.0f / d[1], 1.0f / d[2]};
        float t = hits_t[2 * r]; const float t2 = hits_t[2 * r + 1];
        int s = 0;
        while (t < t2 && s < n_samples) {
            const probe_t p = probe(t, o, d, inv, &c, bits);
            if (p.occ) {
                const int64_t k = n * (int64_t)n_samples + s;
                xyzs[3 * k] = p.x; xyzs[3 * k + 1] = p.y; xyzs[3 * k + 2] = p.z;
                dirs[3 * k] = d[0]; dirs[3 * k + 1] = d[1]; dirs[3 * k + 2] = d[2];
                ts[k] = t; deltas[k] = p.dt;
                t = t + p.dt;
                hits_t[2 * r] = t;                                  /* l.390 */
                ++s;
            } else { do { t = calc_dt(t, &c) + t; } while (t < p.t_target); }
        }
        n_eff[n] = s;
    }
}

/* ---- compositing (ref: volumerendering.cu:6-45) ---- */
void orc_composite_train_fw(const float* sigmas, const float* rgbs, const float* deltas, const float* ts, const int64_t* rays_a, float T_thr,
                            int64_t n_rays, int64_t* total_samples, float* opacity, float* depth, float* rgb, float* ws) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t n = 0; n < n_rays; ++n) {
        const int64_t ray = rays_a[3 * n], start = rays_a[3 * n + 1]; const int N = (int)rays_a[3 * n + 2];
        int samples = 0; float T = 1.0f, r = 0.f, g = 0.f, b = 0.f, d = 0.f, o = 0.f;
        for (int k = 0; k < N; ++k) ws[start + k] = 0.f;
        while (samples < N) {
            const int64_t s = start + samples;
            const float a = 1.0f - expf(-(sigmas[s] * deltas[s]));
            const float w = a * T;
            r = fmaf(w, rgbs[3 * s], r); g = fmaf(w, rgbs[3 * s + 1], g); b = fmaf(w, rgbs[3 * s + 2], b);
            d = fmaf(w, ts[s], d); o = o + w; ws[s] = w;
            T = T * (1.0f - a);
            if (T <= T_thr) break;
            ++samples;
        }
        total_samples[ray] = samples; opacity[ray] = o; depth[ray] = d; rgb[3 * ray] = r; rgb[3 * ray + 1] = g; rgb[3 * ray + 2] = b;
    }
}

/* ref: volumerendering.cu:87-151 (+ the dL_dws*ws temporary of l.175) */
void orc_composite_train_bw(const float* gO, const float* gD, const float* gRGB, const float* gW, const float* sigmas, const float* rgbs,
                            const float* ws, const float* deltas, const float* ts, const int64_t* rays_a, const float* opacity,
                            const float* depth, const float* rgb, float T_thr, int64_t n_rays, int64_t n_samples, float* dsig, float* drgbs) {
    memset(dsig, 0, sizeof(float) * (size_t)n_samples); memset(drgbs, 0, sizeof(float) * 3 * (size_t)n_samples);
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t n = 0; n < n_rays; ++n) {
        const int64_t ray = rays_a[3 * n], start = rays_a[3 * n + 1]; const int N = (int)rays_a[3 * n + 2];
        if (N == 0) continue;
        float* pre = (float*)malloc(sizeof(float) * (size_t)N);
        float run = 0.f;
        for (int k = 0; k < N; ++k) { run = run + gW[start + k] * ws[start + k]; pre[k] = run; }
        const float tot = pre[N - 1];
        const float R = rgb[3 * ray], G = rgb[3 * ray + 1], B = rgb[3 * ray + 2], O = opacity[ray], D = depth[ray];
        float T = 1.0f, r = 0.f, g = 0.f, b = 0.f, d = 0.f; int samples = 0;
        while (samples < N) {
            const int64_t s = start + samples;
            const float a = 1.0f - expf(-(sigmas[s] * deltas[s]));
            const float w = a * T;
            r = fmaf(w, rgbs[3 * s], r); g = fmaf(w, rgbs[3 * s + 1], g); b = fmaf(w, rgbs[3 * s + 2], b); d = fmaf(w, ts[s], d);
            T = T * (1.0f - a);
            drgbs[3 * s] = gRGB[3 * ray] * w; drgbs[3 * s + 1] = gRGB[3 * ray + 1] * w; drgbs[3 * s + 2] = gRGB[3 * ray + 2] * w;
            dsig[s] = deltas[s] * (gRGB[3 * ray] * (rgbs[3 * s] * T - (R - r)) + gRGB[3 * ray + 1] * (rgbs[3 * s + 1] * T - (G - g)) +
                                   gRGB[3 * ray + 2] * (rgbs[3 * s + 2] * T - (B - b)) + gO[ray] * (1 - O) + gD[ray] * (ts[s] * T - (D - d)) +
                                   T * gW[s] - (tot - pre[samples]));
            if (T <= T_thr) break;
            ++samples;
        }
        free(pre);
    }
}

/* ref: volumerendering.cu:205-249 */
void orc_composite_test_fw(const float* sigmas, const float* rgbs, const float* deltas, const float* ts, int64_t* alive, float T_thr,
                           const int32_t* n_eff, int n_samples, int64_t n_alive, float* opacity, float* depth, float* rgb) {
    for (int64_t n = 0; n < n_alive; ++n) {
        if (n_eff[n] == 0) { alive[n] = -1; continue; }
        const int64_t r = alive[n];
        int s = 0; float T = 1 - opacity[r];
        while (s < n_eff[n]) {
            const int64_t k = n * (int64_t)n_samples + s;
            const float a = 1.0f - expf(-(sigmas[k] * deltas[k]));
            const float w = a * T;
            rgb[3 * r] = fmaf(w, rgbs[3 * k], rgb[3 * r]); rgb[3 * r + 1] = fmaf(w, rgbs[3 * k + 1], rgb[3 * r + 1]);
            rgb[3 * r + 2] = fmaf(w, rgbs[3 * k + 2], rgb[3 * r + 2]);
            depth[r] = fmaf(w, ts[k], depth[r]); opacity[r] = opacity[r] + w;
            T = T * (1.0f - a);
            if (T <= T_thr) { alive[n] = -1; break; }
            ++s;
        }
    }
}

/* ---- distortion loss (ref: losses.cu:9-61, 64-109) ---- */
void orc_distortion_fw(const float* ws, const float* deltas, const float* ts, const int64_t* rays_a, int64_t n_rays, float* loss,
                       float* ws_incl, float* wts_incl) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t n = 0; n < n_rays; ++n) {
        const int64_t ray = rays_a[3 * n], start = rays_a[3 * n + 1]; const int N = (int)rays_a[3 * n + 2];
        float wi = 0.f, wti = 0.f, acc = 0.f;
        for (int k = 0; k < N; ++k) {
            const int64_t s = start + k;
            const float we = wi, wte = wti;                 /* exclusive scans */
            const float wt = ws[s] * ts[s];
            wi = wi + ws[s]; wti = wti + wt;                /* inclusive scans */
            ws_incl[s] = wi; wts_incl[s] = wti;
            acc = acc + (2 * (wti * we - wi * wte) + 0.33333334f * ws[s] * ws[s] * deltas[s]);
        }
        loss[ray] = acc;
    }
}
/* ref: losses.cu:112-142 */
void orc_distortion_bw(const float* gL, const float* ws_incl, const float* wts_incl, const float* ws, const float* deltas, const float* ts,
                       const int64_t* rays_a, int64_t n_rays, float* gW) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t n = 0; n < n_rays; ++n) {
        const int64_t ray = rays_a[3 * n], start = rays_a[3 * n + 1]; const int N = (int)rays_a[3 * n + 2];
        if (N == 0) continue;
        const int64_t end = start + N - 1;
        const float w_sum = ws_incl[end], wt_sum = wts_incl[end];
        for (int64_t s = start; s <= end; ++s) {
            float v = gL[ray] * 2 * ((s == start ? 0.f : (ts[s] * ws_incl[s - 1] - wts_incl[s - 1])) +
                                     (wt_sum - wts_incl[s] - ts[s] * (w_sum - ws_incl[s])));
            v += gL[ray] * 2.0f / 3 * ws[s] * deltas[s];
            gW[s] = v;
        }
    }
}
