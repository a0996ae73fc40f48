// Fused NGP field kernels for sm_100a: hash-grid gather + sigma MLP + SH + rgb MLP in ONE kernel (forward), and the whole MLP
// backward (dgrad chain + weight gradients) in ONE kernel, on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM).
// Replaces, for the standard NGP shape (L*F = 32 encoded features with F = 2, 64-wide sigma net with one hidden layer, 64-wide
// rgb net with 1 or 2 hidden layers), the v1 pipeline of field.cu (encoder kernel -> mma.sync MLP kernels with every
// intermediate in HBM).  Semantics: models/networks.py:96-155 (NGP.density / NGP.forward), TruncExp custom_functions.py:162-173.
//
// Forward, one CTA = 128 samples, 256 threads (two threads per sample row; row r == TMEM lane r), persistent over tiles, 5 CTAs / SM:
//   gather 16 levels x 8 corners (lane pairs, fp32 interpolation) -> X tile [128x32] fp16 in shared memory (canonical un-swizzled
//   core-matrix layout, umma.cuh) -> tcgen05.mma X.W1^T -> TMEM -> ReLU -> H1 tile -> mma H1.W2^T -> h (16) -> sigma = exp(h0);
//   CAT tile = [SH4(dir) | h] -> mma CAT.W3^T -> ReLU -> mma .W4^T -> ReLU -> mma .W5^T -> sigmoid -> rgb.
//   X, H and CAT live one after the other in ONE 16 KiB region and the two accumulators share 64 TMEM columns (36 KiB of shared
//   memory and 48 registers per thread are what lets 5 CTAs share an SM; the gather is bound by resident parallelism).
//   Nothing but xyz/dir in and sigma/rgb out touches HBM at inference.  In training the forward pass saves, per sample, only what
//   the backward pass cannot cheaply recompute: the encoded features X (64 B, the 8 KiB X tile *in the shared-memory layout*, one
//   bulk async store per tile), the view direction (12 B), the fp16 rgb outputs (8 B) and the normalised position for the scatter
//   kernel (16 B) -- 100 B/sample instead of round 1's 540 B/sample of activation blobs.  The hidden activations H1 / CAT / H2 /
//   H3 never leave the SM: the backward kernel re-runs the four forward MMAs on the saved X tile (same instruction shapes, same
//   operands -> bit-identical activations; the tensor pipe was 4-17 % busy) before its dgrad / wgrad chain.
//   (MFN_FIELD_SAVE=full keeps round 1's behaviour -- all five tiles stored, 64 KiB per tile -- for A/B measurements.)
// Backward, one CTA = 128 samples, 256 threads, persistent, 2 CTAs / SM (88 KiB of shared memory, 256 TMEM columns):
//   one 8 KiB bulk async copy (X tile) -> recompute H1, h, CAT = [SH(dir) | h], H2, H3 ->
//   dZ5 = dL/drgb * sigmoid' -> [dgrad mma -> TMEM -> ReLU mask -> dZ tile (in place of the activation it masks)] x 4 -> dX;
//   the weight gradients dW = dZ^T.A are tcgen05 MMAs with M = 64, both operands read MN-major from the very same tiles, and
//   they ACCUMULATE IN TMEM across all tiles of the CTA (fp32); one partial per CTA is written at the end and an extra grid row of
//   the scatter kernel (encoder.cu) sums the partials (deterministic, no atomics).  dX (fp16, loss-scaled, level-major) goes to the
//   hash-grid scatter kernel.
#include "field_internal.h"
#include "grid_common.cuh"
#include "sh4.cuh"
#include "umma.cuh"
#include <stdlib.h>

namespace mfn {
using namespace umma;

constexpr int kFT = 128;                 // samples per tile
// shared-memory byte offsets (both kernels): weights first, all tiles in the canonical row-core layout
constexpr int kW1 = 0;                   // [64 x 32]
constexpr int kW2 = kW1 + 64 * 32 * 2;   // [16 x 64]
constexpr int kW3 = kW2 + 16 * 64 * 2;   // [64 x 32]
constexpr int kW4 = kW3 + 64 * 32 * 2;   // [64 x 64]
constexpr int kW5 = kW4 + 64 * 64 * 2;   // [16 x 64]
constexpr int kWEnd = kW5 + 16 * 64 * 2; // 20480
// saved-activation blob of one tile (== the shared-memory image of the backward kernel's tile area)
constexpr int kBX = 0;                       // X   [128 x 32]
constexpr int kBH1 = kBX + kFT * 32 * 2;     // H1  [128 x 64]
constexpr int kBC = kBH1 + kFT * 64 * 2;     // CAT [128 x 32] = [SH | h]
constexpr int kBH2 = kBC + kFT * 32 * 2;     // H2  [128 x 64]
constexpr int kBH3 = kBH2 + kFT * 64 * 2;    // H3  [128 x 64] (rgb nets with two hidden layers)
constexpr int kBlob = kBH3 + kFT * 64 * 2;   // 65536
static_assert(kBlob == 65536, "blob size");
// forward kernel shared memory: weights | X | H | CAT
constexpr int kFwdX = kWEnd, kFwdH = kFwdX + kFT * 32 * 2, kFwdC = kFwdH + kFT * 64 * 2, kFwdSmem = kFwdC + kFT * 32 * 2;
// backward kernel shared memory: weights | blob image | dZo [128 x 16]
constexpr int kBwdBlob = kWEnd, kBwdDZo = kBwdBlob + kBlob, kBwdSmem = kBwdDZo + kFT * 16 * 2;
// TMEM columns
constexpr int kFwdCols = 128, kAccH = 0, kAccO = 64;
constexpr int kBwdCols = 256, kAccW5 = 64, kAccW4 = 80, kAccW3 = 144, kAccW2 = 176, kAccW1 = 192, kAccO2 = 224;   // kAccO2: h of the recomputed forward
constexpr int kBlobX = kFT * 32 * 2;     // saved per tile when the backward pass recomputes the hidden activations: the X tile only
constexpr int kNumWg = 64 * 32 + 16 * 64 + 64 * 32 + 64 * 64 + 16 * 64;   // 10240 weight-gradient floats per partial

// global row-major [rows][cols] fp16 matrix -> row-core tile in shared memory
__device__ __forceinline__ void stage_weight(unsigned char* dst, const __half* __restrict__ src, int rows, int cols, int tid, int nthreads) {
    const int cpr = cols >> 3;
    for (int q = tid; q < rows * cpr; q += nthreads) {
        const int r = q / cpr, c = (q % cpr) << 3;
        *reinterpret_cast<uint4*>(dst + tile_off(r, c, cols)) = __ldg(reinterpret_cast<const uint4*>(src + (size_t)r * cols + c));
    }
}

template <int NH2>
__device__ __forceinline__ void stage_all_weights(unsigned char* smem, const FusedArgs& a, int tid, int nthreads, bool rgb) {
    stage_weight(smem + kW1, a.w_sigma, 64, 32, tid, nthreads);
    stage_weight(smem + kW2, a.w_sigma + 64 * 32, 16, 64, tid, nthreads);
    if (rgb) {
        stage_weight(smem + kW3, a.w_rgb, 64, 32, tid, nthreads);
        if (NH2 == 2) stage_weight(smem + kW4, a.w_rgb + 64 * 32, 64, 64, tid, nthreads);
        stage_weight(smem + kW5, a.w_rgb + 64 * 32 + (NH2 - 1) * 64 * 64, 16, 64, tid, nthreads);
    }
}

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// one level of the hash grid for one sample, F = 2: 8 gathers of 4 bytes, fp32 trilinear interpolation -> packed half2
__device__ __forceinline__ uint32_t gather_level(const uint32_t* __restrict__ table, const GridMeta& m, int l, float x, float y, float z) {
    const float s = m.scale[l];
    const uint32_t res = m.res[l], off = m.offset[l], size = m.offset[l + 1] - off;
    const float px = fmaf(x, s, 0.5f), py = fmaf(y, s, 0.5f), pz = fmaf(z, s, 0.5f);
    const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
    const float wx = px - fx, wy = py - fy, wz = pz - fz;
    const uint32_t gx = (uint32_t)(int)fx, gy = (uint32_t)(int)fy, gz = (uint32_t)(int)fz;
    const uint32_t* lvl = table + off;
    uint32_t idx[8];
    if ((m.hashed >> l) & 1u) {              // size is a power of two for hashed levels
        const uint32_t mask = size - 1u;
        const uint32_t hy0 = gy * 2654435761u, hy1 = hy0 + 2654435761u, hz0 = gz * 805459861u, hz1 = hz0 + 805459861u;
        const uint32_t h00 = hy0 ^ hz0, h10 = hy1 ^ hz0, h01 = hy0 ^ hz1, h11 = hy1 ^ hz1, gx1 = gx + 1u;
        idx[0] = (gx ^ h00) & mask; idx[1] = (gx1 ^ h00) & mask; idx[2] = (gx ^ h10) & mask; idx[3] = (gx1 ^ h10) & mask;
        idx[4] = (gx ^ h01) & mask; idx[5] = (gx1 ^ h01) & mask; idx[6] = (gx ^ h11) & mask; idx[7] = (gx1 ^ h11) & mask;
    } else {                                  // dense: x + y*res + z*res^2, mod size (only a corner on the x/y/z == res border wraps)
        const uint32_t r2 = res * res;
        const uint32_t b00 = gx + gy * res + gz * r2;
        const bool wrap = b00 + 1u + res + r2 >= size;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            uint32_t i = b00 + (c & 1) + ((c >> 1) & 1) * res + (c >> 2) * r2;
            if (wrap) i %= size;
            idx[c] = i;
        }
    }
    uint32_t v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = __ldg(lvl + idx[c]);
    const float ux = 1.f - wx, uy = 1.f - wy, uz = 1.f - wz;
    float f0 = 0.f, f1 = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {    // same weight expression and accumulation order as encode_level (grid_common.cuh)
        const float w = ((c & 1) ? wx : ux) * (((c >> 1) & 1) ? wy : uy) * ((c >> 2) ? wz : uz);
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&v[c]));
        f0 = fmaf(w, f.x, f0); f1 = fmaf(w, f.y, f1);
    }
    return pack2(f0, f1);
}

// Lane-pair variant: the two lanes of a pair work on the same sample and each fetches the 4 corners of ITS x (gx + xb), so the
// corners (x, x+1) of one (y, z) -- same 32-byte sector 7 times out of 8 -- are requested by the same load instruction: a warp's
// gather touches <= 16 sectors instead of <= 32, and the L1 sector rate is what bounds this phase.  Returns the pair's sum.
template <bool MIXED>
__device__ __forceinline__ uint32_t gather_level_pair(const uint32_t* __restrict__ table, const GridMeta& m, int l, float x, float y, float z, int xb, uint64_t pol) {
    const float s = m.scale[l];
    const uint32_t res = m.res[l], off = m.offset[l], size = m.size[l];
    const float px = fmaf(x, s, 0.5f), py = fmaf(y, s, 0.5f), pz = fmaf(z, s, 0.5f);
    const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
    const float wx = px - fx, wy = py - fy, wz = pz - fz;
    const uint32_t cx = (uint32_t)(int)fx + (uint32_t)xb, gy = (uint32_t)(int)fy, gz = (uint32_t)(int)fz;
    const uint32_t* lvl = table + off;
    uint32_t idx[4];
    if (MIXED) {   // MixedFeature grid: vertices are hashed by their canonical-grid coordinates (grid_common.cuh: canon_vertex); the
                   // lane pair still splits the x corners, it just no longer finds them in one sector
        const float r = m.canon[l];
        const uint32_t mask = size - 1u, ccx = canon_vertex(cx, r);
        const uint32_t hy0 = canon_vertex(gy, r) * 2654435761u, hy1 = canon_vertex(gy + 1u, r) * 2654435761u;
        const uint32_t hz0 = canon_vertex(gz, r) * 805459861u, hz1 = canon_vertex(gz + 1u, r) * 805459861u;
        idx[0] = (ccx ^ hy0 ^ hz0) & mask; idx[1] = (ccx ^ hy1 ^ hz0) & mask; idx[2] = (ccx ^ hy0 ^ hz1) & mask; idx[3] = (ccx ^ hy1 ^ hz1) & mask;
    } else if ((m.hashed >> l) & 1u) {
        const uint32_t mask = size - 1u;
        const uint32_t hy0 = gy * 2654435761u, hy1 = hy0 + 2654435761u, hz0 = gz * 805459861u, hz1 = hz0 + 805459861u;
        idx[0] = (cx ^ hy0 ^ hz0) & mask; idx[1] = (cx ^ hy1 ^ hz0) & mask; idx[2] = (cx ^ hy0 ^ hz1) & mask; idx[3] = (cx ^ hy1 ^ hz1) & mask;
    } else {
        const uint32_t r2 = res * res;
        const uint32_t b = cx + gy * res + gz * r2;
        idx[0] = b; idx[1] = b + res; idx[2] = b + r2; idx[3] = b + res + r2;
        if (idx[3] >= size) {
#pragma unroll
            for (int c = 0; c < 4; ++c) idx[c] %= size;
        }
    }
    uint32_t v[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) v[c] = ldg_nc_hint(lvl + idx[c], pol);   // the table is to stay L2-resident (evict_last)
    const float wxs = xb ? wx : 1.f - wx;
    const float a0 = wxs * (1.f - wy), a1 = wxs * wy;
    const float w[4] = {a0 * (1.f - wz), a1 * (1.f - wz), a0 * wz, a1 * wz};
    float f0 = 0.f, f1 = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&v[c]));
        f0 = fmaf(w[c], f.x, f0); f1 = fmaf(w[c], f.y, f1);
    }
    f0 += __shfl_xor_sync(0xffffffffu, f0, 1);
    f1 += __shfl_xor_sync(0xffffffffu, f1, 1);
    return pack2(f0, f1);
}

// SH (degree 4) of the normalised direction mapped like networks.py:145-146 -> 16 fp16 values = CAT[:, 0:16] of one row
__device__ __forceinline__ void sh_of_dir(float dx, float dy, float dz, uint4& o0, uint4& o1) {
    const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);
    const float ux = (dx / nrm + 1.0f) / 2.0f, uy = (dy / nrm + 1.0f) / 2.0f, uz = (dz / nrm + 1.0f) / 2.0f;
    float s[16];
    sh4_eval(fmaf(ux, 2.f, -1.f), fmaf(uy, 2.f, -1.f), fmaf(uz, 2.f, -1.f), s);
    o0.x = pack2(s[0], s[1]); o0.y = pack2(s[2], s[3]); o0.z = pack2(s[4], s[5]); o0.w = pack2(s[6], s[7]);
    o1.x = pack2(s[8], s[9]); o1.y = pack2(s[10], s[11]); o1.z = pack2(s[12], s[13]); o1.w = pack2(s[14], s[15]);
}

__device__ __forceinline__ float act_out(float x, int act) {
    if (act == MFN_ACT_SIGMOID) return 1.0f / (1.0f + __expf(-x));
    if (act == MFN_ACT_EXP) return __expf(x);
    return x;
}

// TMEM (32 fp32 columns starting at col0) -> ReLU -> fp16 -> columns col0..col0+31 of row `row` of a [128 x 64] row-core tile
// (16 columns per tcgen05.ld: the inference kernel lives on 48 registers to fit 5 CTAs per SM)
__device__ __forceinline__ void relu_epilogue32(uint32_t taddr, unsigned char* tile, int row, int col0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        uint32_t r[16];
        tmem_ld_x16(taddr + col0 + 16 * h, r);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint4 o;
            o.x = pack2(fmaxf(__uint_as_float(r[8 * c + 0]), 0.f), fmaxf(__uint_as_float(r[8 * c + 1]), 0.f));
            o.y = pack2(fmaxf(__uint_as_float(r[8 * c + 2]), 0.f), fmaxf(__uint_as_float(r[8 * c + 3]), 0.f));
            o.z = pack2(fmaxf(__uint_as_float(r[8 * c + 4]), 0.f), fmaxf(__uint_as_float(r[8 * c + 5]), 0.f));
            o.w = pack2(fmaxf(__uint_as_float(r[8 * c + 6]), 0.f), fmaxf(__uint_as_float(r[8 * c + 7]), 0.f));
            *reinterpret_cast<uint4*>(tile + tile_off(row, col0 + 16 * h + c * 8, 64)) = o;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------ forward
// MODE 0: inference (sigma + rgb), 1: training (also saves X tile, direction, fp16 rgb and normalised position), 2: density only (sigma),
// 3: the 16 raw outputs of the grid + sigma MLP in fp16 (tcnn.NetworkWithInputEncoding's forward, networks.py:107),
// 4: training with all five activation tiles stored (round 1's 64 KiB blobs; MFN_FIELD_SAVE=full).
// 256 threads: thread t works on sample row t % 128 (= its TMEM lane); the two threads of a row split the 16 grid levels in the
// gather phase and the 64 accumulator columns in the hidden-layer epilogues.  In training mode every published tile is also sent
// to the blob with one bulk async store (shared -> global) issued by thread 0.
constexpr int kFwdThreads = 256;
constexpr bool kTrainAlias = true;      // training mode too: one tile region, every bulk store waited for before its tile is overwritten
// phase timestamps of CTA 0 (tools/fwd_phases.py): a.dbg != nullptr only in that tool
#define MFN_TS(k) do { if (a.dbg && blockIdx.x == 0 && tid == (k >= 100 ? 255 : 0) && tile_no < 12) a.dbg[tile_no * 16 + (k % 100)] = clock64(); } while (0)

template <int NH2, int MODE, bool MIXED>
__global__ void __launch_bounds__(kFwdThreads, ((MODE == 1 || MODE == 4) && !kTrainAlias) ? 4 : 5)
field_fwd_fused_kernel(const __grid_constant__ FusedArgs a, const __grid_constant__ GridMeta m) {
    constexpr bool TRAIN = (MODE == 1 || MODE == 4), FULL = (MODE == 4), RGB = (MODE < 2 || MODE == 4);
    // Inference / density modes keep no tile alive across stages, so X, H and CAT share ONE 16 KiB region and the 16-column output
    // accumulator shares the hidden accumulator's TMEM columns: 36 KiB + 64 columns per CTA -> 5 CTAs / SM instead of 4 (the gather
    // is bound by resident parallelism).  Training mode streams every tile to the blob and keeps the three regions apart.
    constexpr bool ALIAS = !TRAIN || kTrainAlias;
    constexpr int oX = kFwdX, oH = ALIAS ? kFwdX : kFwdH, oC = ALIAS ? kFwdX : kFwdC;
    constexpr int accO = ALIAS ? kAccH : kAccO;
    constexpr int nCols = ALIAS ? 64 : kFwdCols;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int row = tid & (kFT - 1), hsel = tid >> 7;
    const int64_t n = a.n_dev ? min((int64_t)*a.n_dev, a.n_max) : a.n_max;
    const int64_t n_tiles = (n + kFT - 1) / kFT;
    if (TRAIN && a.n_out && blockIdx.x == 0 && tid == 0) *a.n_out = (int32_t)n;     // the backward pass's own copy of the count
    if ((int64_t)blockIdx.x >= n_tiles) return;      // the grid is sized for n_max; with a device-side count most CTAs may have nothing to do
    stage_all_weights<NH2>(smem, a, tid, kFwdThreads, RGB);
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tmem_base_s, nCols);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_base_s;
    const uint32_t trow = tmem_addr(tbase, (warp & 3) * 32, 0);   // this warp's lane quarter
    const uint32_t sbase = smem_u32(smem);
    uint32_t phase = 0;
    const uint32_t* table = reinterpret_cast<const uint32_t*>(a.table);
    const uint64_t pol_keep = policy_evict_last(), pol_stream = policy_evict_first();

    int tile_no = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_no) {
        const int64_t i = tile * kFT + row;
        const bool valid = i < n;
        unsigned char* blob = TRAIN ? a.blobs + (size_t)tile * (FULL ? kBlob : kBlobX) : nullptr;
        MFN_TS(0);
        // ---- hash-grid gather -> X tile: lane pair (2p, 2p+1) of warp w works on row 16w + p, all 16 levels
        {
            const int grow = tid >> 1, xb = tid & 1;
            const int64_t gi = tile * kFT + grow;
            const bool gvalid = gi < n;
            float x = 0.5f, y = 0.5f, z = 0.5f;
            if (gvalid) {
                x = a.xyzs[3 * gi]; y = a.xyzs[3 * gi + 1]; z = a.xyzs[3 * gi + 2];
                x = __fdiv_rn(__fsub_rn(x, a.mn[0]), __fsub_rn(a.mx[0], a.mn[0]));      // networks.py:105
                y = __fdiv_rn(__fsub_rn(y, a.mn[1]), __fsub_rn(a.mx[1], a.mn[1]));
                z = __fdiv_rn(__fsub_rn(z, a.mn[2]), __fsub_rn(a.mx[2], a.mn[2]));
                if (TRAIN && xb == 1) a.x01[gi] = make_float4(x, y, z, 0.f);
            }
#pragma unroll 2
            for (int l = 0; l < 16; ++l) {
                const uint32_t v = gather_level_pair<MIXED>(table, m, l, x, y, z, xb, pol_keep);    // (invalid rows gather entry 0 harmlessly)
                if (xb == (l & 1)) *reinterpret_cast<uint32_t*>(smem + oX + tile_off(grow, 2 * l, 32)) = gvalid ? v : 0u;
            }
        }
        MFN_TS(1); MFN_TS(109);
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        MFN_TS(2);
        // ---- layer 1: H1 = relu(X . W1^T)
        if (tid == 0) {
            tc_fence_after();
            const uint32_t id = idesc_f16(128, 64, false, false);
#pragma unroll
            for (int k0 = 0; k0 < 32; k0 += 16) mma_f16_ss(tbase + kAccH, desc_kmajor(sbase + oX, 32, k0), desc_kmajor(sbase + kW1, 32, k0), id, k0 > 0);
            if (TRAIN) { bulk_s2g_hint(blob + kBX, smem + oX, kFT * 32 * 2, pol_stream); bulk_commit(); if (ALIAS) bulk_wait_read0(); }   // (shared region: H1 overwrites X)
            mma_commit(&bar);
        }
        mbar_wait(&bar, phase); phase ^= 1u;
        MFN_TS(3);
        tc_fence_after();
        relu_epilogue32(trow + kAccH, smem + oH, row, 32 * hsel);
        MFN_TS(4);
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        MFN_TS(5);
        // ---- layer 2: h = H1 . W2^T (16 outputs, no activation); sigma = exp(h0)  (TruncExp forward)
        if (tid == 0) {
            tc_fence_after();
            const uint32_t id = idesc_f16(128, 16, false, false);
#pragma unroll
            for (int k0 = 0; k0 < 64; k0 += 16) mma_f16_ss(tbase + accO, desc_kmajor(sbase + oH, 64, k0), desc_kmajor(sbase + kW2, 64, k0), id, k0 > 0);
            if (FULL) { bulk_s2g_hint(blob + kBH1, smem + oH, kFT * 64 * 2, pol_stream); bulk_commit(); if (ALIAS) bulk_wait_read0(); }   // (shared region: CAT overwrites H1)
            mma_commit(&bar);
        }
        if (hsel == 0) {
            mbar_wait(&bar, phase);
            tc_fence_after();
            uint32_t r[16];
            tmem_ld_x16(trow + accO, r);
            tmem_ld_wait();
            uint4 o0, o1;
            o0.x = pack2(__uint_as_float(r[0]), __uint_as_float(r[1])); o0.y = pack2(__uint_as_float(r[2]), __uint_as_float(r[3]));
            o0.z = pack2(__uint_as_float(r[4]), __uint_as_float(r[5])); o0.w = pack2(__uint_as_float(r[6]), __uint_as_float(r[7]));
            o1.x = pack2(__uint_as_float(r[8]), __uint_as_float(r[9])); o1.y = pack2(__uint_as_float(r[10]), __uint_as_float(r[11]));
            o1.z = pack2(__uint_as_float(r[12]), __uint_as_float(r[13])); o1.w = pack2(__uint_as_float(r[14]), __uint_as_float(r[15]));
            if (MODE == 3) {
                if (valid) { uint4* ho = reinterpret_cast<uint4*>(a.h_out + 16 * i); ho[0] = o0; ho[1] = o1; }
            } else if (valid) a.sigmas[i] = expf(__low2float(*reinterpret_cast<const __half2*>(&o0.x)));
            if (RGB) {
                *reinterpret_cast<uint4*>(smem + oC + tile_off(row, 16, 32)) = o0;
                *reinterpret_cast<uint4*>(smem + oC + tile_off(row, 24, 32)) = o1;
            }
        } else if (RGB) {
            // the partner thread of the row meanwhile encodes the direction: SH of the normalised direction -> CAT[:, 0:16]
            // (networks.py:145-146); CAT may overlay H1, which the layer-2 MMA has finished reading
            mbar_wait(&bar, phase);
            uint4 o0 = make_uint4(0u, 0u, 0u, 0u), o1 = o0;
            if (valid) {
                const float dx = a.dirs[3 * i], dy = a.dirs[3 * i + 1], dz = a.dirs[3 * i + 2];
                if (MODE == 1) { a.dirs_copy[3 * i] = dx; a.dirs_copy[3 * i + 1] = dy; a.dirs_copy[3 * i + 2] = dz; }   // the backward pass re-encodes it
                sh_of_dir(dx, dy, dz, o0, o1);
            }
            *reinterpret_cast<uint4*>(smem + oC + tile_off(row, 0, 32)) = o0;
            *reinterpret_cast<uint4*>(smem + oC + tile_off(row, 8, 32)) = o1;
        }
        phase ^= 1u;
        MFN_TS(6);
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        MFN_TS(7);
        if (!RGB) continue;   // (uniform) density only / raw outputs
        // ---- rgb layer 1: H2 = relu(CAT . W3^T)
        if (tid == 0) {
            tc_fence_after();
            const uint32_t id = idesc_f16(128, 64, false, false);
#pragma unroll
            for (int k0 = 0; k0 < 32; k0 += 16) mma_f16_ss(tbase + kAccH, desc_kmajor(sbase + oC, 32, k0), desc_kmajor(sbase + kW3, 32, k0), id, k0 > 0);
            if (FULL) { bulk_wait_read0(); bulk_s2g_hint(blob + kBC, smem + oC, kFT * 32 * 2, pol_stream); bulk_commit(); if (ALIAS) bulk_wait_read0(); }
            mma_commit(&bar);
        }
        mbar_wait(&bar, phase); phase ^= 1u;
        tc_fence_after();
        relu_epilogue32(trow + kAccH, smem + oH, row, 32 * hsel);
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (NH2 == 2) {
            // ---- rgb layer 2: H3 = relu(H2 . W4^T)
            if (tid == 0) {
                tc_fence_after();
                const uint32_t id = idesc_f16(128, 64, false, false);
#pragma unroll
                for (int k0 = 0; k0 < 64; k0 += 16) mma_f16_ss(tbase + kAccH, desc_kmajor(sbase + oH, 64, k0), desc_kmajor(sbase + kW4, 64, k0), id, k0 > 0);
                if (FULL) { bulk_s2g_hint(blob + kBH2, smem + oH, kFT * 64 * 2, pol_stream); bulk_commit(); bulk_wait_read0(); }
                mma_commit(&bar);
            }
            mbar_wait(&bar, phase); phase ^= 1u;
            tc_fence_after();
            relu_epilogue32(trow + kAccH, smem + oH, row, 32 * hsel);
            fence_async_smem();
            tc_fence_before();
            __syncthreads();
        }
        // ---- rgb output layer: rgb = act(H . W5^T)[0:3], rounded to fp16 like tcnn's output
        if (tid == 0) {
            tc_fence_after();
            const uint32_t id = idesc_f16(128, 16, false, false);
#pragma unroll
            for (int k0 = 0; k0 < 64; k0 += 16) mma_f16_ss(tbase + accO, desc_kmajor(sbase + oH, 64, k0), desc_kmajor(sbase + kW5, 64, k0), id, k0 > 0);
            if (FULL) { bulk_s2g_hint(blob + (NH2 == 2 ? kBH3 : kBH2), smem + oH, kFT * 64 * 2, pol_stream); bulk_commit(); }
            mma_commit(&bar);
        }
        if (hsel == 0) {
            mbar_wait(&bar, phase);
            tc_fence_after();
            uint32_t r[8];
            tmem_ld_x8(trow + accO, r);
            tmem_ld_wait();
            if (valid) {
                __half hv[4];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    hv[k] = __float2half_rn(act_out(__uint_as_float(r[k]), a.rgb_act));
                    a.rgbs[3 * i + k] = __half2float(hv[k]);
                }
                hv[3] = __float2half_rn(0.f);
                if (TRAIN) a.rgb_h[i] = *reinterpret_cast<const uint2*>(hv);      // the outputs ARE fp16 values: 8 B/sample for the backward pass
            }
        }
        phase ^= 1u;
        if (FULL && tid == 0) bulk_wait_read0();   // the last stores (CAT, last hidden tile) have read shared memory: the next tile may overwrite it
        tc_fence_before();
        __syncthreads();   // TMEM and the tiles are free for the next tile
        MFN_TS(8);
    }
    if (TRAIN && tid == 0) bulk_wait0();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, nCols);
}

// ------------------------------------------------------------------------------------------------------------------ backward
constexpr int kBwdThreads = 256;

// TMEM (fp32 dgrad accumulator) -> mask with relu'(activation tile row) -> fp16 dZ row written IN PLACE of the activation row
__device__ __forceinline__ void mask_epilogue32(uint32_t taddr, unsigned char* tile, int row, int col0) {
    {
        uint32_t r[32];
        tmem_ld_x32(taddr + col0, r);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            uint4* p = reinterpret_cast<uint4*>(tile + tile_off(row, col0 + c * 8, 64));
            const uint4 act = *p;
            const __half2* ah = reinterpret_cast<const __half2*>(&act);
            uint4 o;
            uint32_t* ow = &o.x;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 av = __half22float2(ah[k]);
                ow[k] = pack2(av.x > 0.f ? __uint_as_float(r[8 * c + 2 * k]) : 0.f, av.y > 0.f ? __uint_as_float(r[8 * c + 2 * k + 1]) : 0.f);
            }
            *p = o;
        }
    }
}

// RC = true: only the X tile was saved; H1, h, CAT, H2, H3 are recomputed here (four more MMAs + epilogues per tile, nothing read
// from HBM but 64 + 12 + 8 B/sample).  RC = false: round 1's path, the five tiles come back as one 64 KiB bulk copy.
// phase timestamps of CTA 0 / thread 0 of the backward kernel (tools/fwd_phases.py): after every MMA-completion wait and every barrier
#define MFN_BTS() do { if (a.dbg && blockIdx.x == 0 && tid == 0 && tile_no < 6 && ts_k < 40) a.dbg[256 + tile_no * 40 + (ts_k++)] = clock64(); } while (0)
template <int NH2, bool RC>
__global__ void __launch_bounds__(kBwdThreads, 2)
field_bwd_fused_kernel(const __grid_constant__ FusedArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar_mma, bar_load;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = tid & (kFT - 1), hsel = tid >> 7;   // two threads per sample row: they split the accumulator columns
    const int64_t n = a.n_dev ? min((int64_t)*a.n_dev, a.n_max) : a.n_max;
    const int64_t n_tiles = (n + kFT - 1) / kFT;
    stage_all_weights<NH2>(smem, a, tid, kBwdThreads, true);
    if (tid == 0) { mbar_init(&bar_mma, 1); mbar_init(&bar_load, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tmem_base_s, kBwdCols);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_base_s;
    const uint32_t trow = tmem_addr(tbase, (warp & 3) * 32, 0);
    const uint32_t sbase = smem_u32(smem);
    unsigned char* sBlob = smem + kBwdBlob;
    const uint32_t sX = sbase + kBwdBlob + kBX, sH1 = sbase + kBwdBlob + kBH1, sC = sbase + kBwdBlob + kBC, sH2 = sbase + kBwdBlob + kBH2,
                   sH3 = sbase + kBwdBlob + kBH3, sDZo = sbase + kBwdDZo;
    const uint32_t sHL = (NH2 == 2) ? sH3 : sH2;                  // last hidden activation of the rgb net
    unsigned char* pHL = sBlob + ((NH2 == 2) ? kBH3 : kBH2);
    uint32_t ph_mma = 0, ph_load = 0;
    uint32_t acc = 0;                                              // 0 on the CTA's first tile: weight-gradient MMAs overwrite
    bool bad = false;

    int tile_no = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_no) {
        int ts_k = 0;
        MFN_BTS();
        const int64_t i = tile * kFT + row;
        const bool valid = i < n;
        if (tid == 0) {   // RC: the saved X tile; else the whole activation blob of this tile -- one bulk async copy either way
            constexpr uint32_t bytes = RC ? kBlobX : (NH2 == 2 ? kBlob : kBH3);
            mbar_arrive_expect_tx(&bar_load, bytes);
            bulk_g2s_hint(sBlob, a.blobs + (size_t)tile * (RC ? kBlobX : kBlob), bytes, &bar_load, policy_evict_first());
        }
        // ---- dZ5 = loss_scale * dL/drgb * act'(rgb)   (16 columns, 3 live)
        if (hsel == 0) {
            float g[3] = {0.f, 0.f, 0.f};
            if (valid) {
                const uint2 yh = a.rgb_h[i];
                const float2 y01 = __half22float2(*reinterpret_cast<const __half2*>(&yh.x));
                const float y3[3] = {y01.x, y01.y, __low2float(*reinterpret_cast<const __half2*>(&yh.y))};
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float yv = y3[k];
                    float d = a.dL_drgbs[3 * i + k] * a.loss_scale;
                    if (a.rgb_act == MFN_ACT_SIGMOID) d *= yv * (1.f - yv);
                    else if (a.rgb_act == MFN_ACT_EXP) d *= yv;
                    g[k] = d;
                }
            }
            uint4 o0 = make_uint4(pack2(g[0], g[1]), pack2(g[2], 0.f), 0u, 0u);
            const __half2* hh = reinterpret_cast<const __half2*>(&o0);
            bad |= !isfinite(__low2float(hh[0])) || !isfinite(__high2float(hh[0])) || !isfinite(__low2float(hh[1]));
            *reinterpret_cast<uint4*>(smem + kBwdDZo + tile_off(row, 0, 16)) = o0;
            *reinterpret_cast<uint4*>(smem + kBwdDZo + tile_off(row, 8, 16)) = make_uint4(0u, 0u, 0u, 0u);
        } else if (RC) {   // the partner thread re-encodes the view direction: CAT[:, 0:16] (same code, same bits as the forward pass)
            uint4 o0 = make_uint4(0u, 0u, 0u, 0u), o1 = o0;
            if (valid) sh_of_dir(a.dirs_copy[3 * i], a.dirs_copy[3 * i + 1], a.dirs_copy[3 * i + 2], o0, o1);
            *reinterpret_cast<uint4*>(sBlob + kBC + tile_off(row, 0, 32)) = o0;
            *reinterpret_cast<uint4*>(sBlob + kBC + tile_off(row, 8, 32)) = o1;
        }
        mbar_wait(&bar_load, ph_load); ph_load ^= 1u;
        fence_async_smem();
        tc_fence_before();
        __syncthreads(); MFN_BTS();
        if (RC) {
            // ---- recomputed forward: H1 = relu(X.W1^T) ; h = H1.W2^T -> CAT[:, 16:32] ; H2 = relu(CAT.W3^T) ; H3 = relu(H2.W4^T)
            if (tid == 0) {
                tc_fence_after();
                const uint32_t id = idesc_f16(128, 64, false, false);
#pragma unroll
                for (int k0 = 0; k0 < 32; k0 += 16) mma_f16_ss(tbase + kAccH, desc_kmajor(sX, 32, k0), desc_kmajor(sbase + kW1, 32, k0), id, k0 > 0);
                mma_commit(&bar_mma);
            }
            mbar_wait(&bar_mma, ph_mma); ph_mma ^= 1u; MFN_BTS();
            tc_fence_after();
            relu_epilogue32(trow + kAccH, sBlob + kBH1, row, 32 * hsel);
            fence_async_smem();
            tc_fence_before();
            __syncthreads(); MFN_BTS();
            if (tid == 0) {
                tc_fence_after();
                const uint32_t id = idesc_f16(128, 16, false, false);
#pragma unroll
                for (int k0 = 0; k0 < 64; k0 += 16) mma_f16_ss(tbase + kAccO2, desc_kmajor(sH1, 64, k0), desc_kmajor(sbase + kW2, 64, k0), id, k0 > 0);
                mma_commit(&bar_mma);
            }
            mbar_wait(&bar_mma, ph_mma); ph_mma ^= 1u; MFN_BTS();
            tc_fence_after();
            if (hsel == 0) {
                uint32_t r[16];
                tmem_ld_x16(trow + kAccO2, r);
                tmem_ld_wait();
                uint4 o0, o1;
                o0.x = pack2(__uint_as_float(r[0]), __uint_as_float(r[1])); o0.y = pack2(__uint_as_float(r[2]), __uint_as_float(r[3]));
                o0.z = pack2(__uint_as_float(r[4]), __uint_as_float(r[5])); o0.w = pack2(__uint_as_float(r[6]), __uint_as_float(r[7]));
                o1.x = pack2(__uint_as_float(r[8]), __uint_as_float(r[9])); o1.y = pack2(__uint_as_float(r[10]), __uint_as_float(r[11]));
                o1.z = pack2(__uint_as_float(r[12]), __uint_as_float(r[13])); o1.w = pack2(__uint_as_float(r[14]), __uint_as_float(r[15]));
                *reinterpret_cast<uint4*>(sBlob + kBC + tile_off(row, 16, 32)) = o0;
                *reinterpret_cast<uint4*>(sBlob + kBC + tile_off(row, 24, 32)) = o1;
            }
            fence_async_smem();
            tc_fence_before();
            __syncthreads(); MFN_BTS();
            if (tid == 0) {
                tc_fence_after();
                const uint32_t id = idesc_f16(128, 64, false, false);
#pragma unroll
                for (int k0 = 0; k0 < 32; k0 += 16) mma_f16_ss(tbase + kAccH, desc_kmajor(sC, 32, k0), desc_kmajor(sbase + kW3, 32, k0), id, k0 > 0);
                mma_commit(&bar_mma);
            }
            mbar_wait(&bar_mma, ph_mma); ph_mma ^= 1u; MFN_BTS();
            tc_fence_after();
            relu_epilogue32(trow + kAccH, sBlob + kBH2, row, 32 * hsel);
            fence_async_smem();
            tc_fence_before();
            __syncthreads(); MFN_BTS();
            if (NH2 == 2) {
                if (tid == 0) {
                    tc_fence_after();
                    const uint32_t id = idesc_f16(128, 64, false, false);
#pragma unroll
                    for (int k0 = 0; k0 < 64; k0 += 16) mma_f16_ss(tbase + kAccH, desc_kmajor(sH2, 64, k0), desc_kmajor(sbase + kW4, 64, k0), id, k0 > 0);
                    mma_commit(&bar_mma);
                }
                mbar_wait(&bar_mma, ph_mma); ph_mma ^= 1u; MFN_BTS();
                tc_fence_after();
                relu_epilogue32(trow + kAccH, sBlob + kBH3, row, 32 * hsel);
                fence_async_smem();
                tc_fence_before();
                __syncthreads(); MFN_BTS();
            }
        }
        // ---- stage A: dH_last = dZ5 . W5 ;  dW5^T += H_last^T . dZ5
        if (tid == 0) {
            tc_fence_after();
            mma_f16_ss(tbase + kAccH, desc_kmajor(sDZo, 16, 0), desc_mnmajor(sbase + kW5, 64, 0), idesc_f16(128, 64, false, true), 0u);
            const uint32_t idw = idesc_f16(64, 16, true, true);
#pragma unroll
            for (int k0 = 0; k0 < kFT; k0 += 16) mma_f16_ss(tbase + kAccW5, desc_mnmajor(sHL, 64, k0), desc_mnmajor(sDZo, 16, k0), idw, acc | (k0 > 0));
            mma_commit(&bar_mma);
        }
        mbar_wait(&bar_mma, ph_mma); ph_mma ^= 1u; MFN_BTS();
        tc_fence_after();
        mask_epilogue32(trow + kAccH, pHL, row, 32 * hsel);                 // dZ of the last hidden layer, in place
        fence_async_smem();
        tc_fence_before();
        __syncthreads(); MFN_BTS();
        if (NH2 == 2) {
            // ---- stage B: dH2 = dZ4 . W4 ;  dW4^T += H2^T . dZ4
            if (tid == 0) {
                tc_fence_after();
                const uint32_t id = idesc_f16(128, 64, false, true);
#pragma unroll
                for (int k0 = 0; k0 < 64; k0 += 16) mma_f16_ss(tbase + kAccH, desc_kmajor(sH3, 64, k0), desc_mnmajor(sbase + kW4, 64, k0), id, k0 > 0);
                const uint32_t idw = idesc_f16(64, 64, true, true);
#pragma unroll
                for (int k0 = 0; k0 < kFT; k0 += 16) mma_f16_ss(tbase + kAccW4, desc_mnmajor(sH2, 64, k0), desc_mnmajor(sH3, 64, k0), idw, acc | (k0 > 0));
                mma_commit(&bar_mma);
            }
            mbar_wait(&bar_mma, ph_mma); ph_mma ^= 1u; MFN_BTS();
            tc_fence_after();
            mask_epilogue32(trow + kAccH, sBlob + kBH2, row, 32 * hsel);    // dZ3 in place of H2
            fence_async_smem();
            tc_fence_before();
            __syncthreads(); MFN_BTS();
        }
        // ---- stage C: dCAT = dZ3 . W3 (32 columns) ;  dW3 += dZ3^T . CAT
        if (tid == 0) {
            tc_fence_after();
            const uint32_t id = idesc_f16(128, 32, false, true);
#pragma unroll
            for (int k0 = 0; k0 < 64; k0 += 16) mma_f16_ss(tbase + kAccH, desc_kmajor(sH2, 64, k0), desc_mnmajor(sbase + kW3, 32, k0), id, k0 > 0);
            const uint32_t idw = idesc_f16(64, 32, true, true);
#pragma unroll
            for (int k0 = 0; k0 < kFT; k0 += 16) mma_f16_ss(tbase + kAccW3, desc_mnmajor(sH2, 64, k0), desc_mnmajor(sC, 32, k0), idw, acc | (k0 > 0));
            mma_commit(&bar_mma);
        }
        mbar_wait(&bar_mma, ph_mma); ph_mma ^= 1u; MFN_BTS();
        tc_fence_after();
        if (hsel == 0) {   // dh = dCAT[:, 16:32] ; dh[0] += loss_scale * dL/dsigma * exp(clamp(h0, -15, 15))   (TruncExp backward)
            uint32_t r[16];
            tmem_ld_x16(trow + kAccH + 16, r);
            tmem_ld_wait();
            float d0 = __uint_as_float(r[0]);
            if (valid) {
                const float h0 = __half2float(*reinterpret_cast<const __half*>(sBlob + kBC + tile_off(row, 16, 32)));
                d0 += a.dL_dsigmas[i] * expf(fminf(fmaxf(h0, -15.f), 15.f)) * a.loss_scale;
            }
            uint4 o0, o1;
            o0.x = pack2(d0, __uint_as_float(r[1])); o0.y = pack2(__uint_as_float(r[2]), __uint_as_float(r[3]));
            o0.z = pack2(__uint_as_float(r[4]), __uint_as_float(r[5])); o0.w = pack2(__uint_as_float(r[6]), __uint_as_float(r[7]));
            o1.x = pack2(__uint_as_float(r[8]), __uint_as_float(r[9])); o1.y = pack2(__uint_as_float(r[10]), __uint_as_float(r[11]));
            o1.z = pack2(__uint_as_float(r[12]), __uint_as_float(r[13])); o1.w = pack2(__uint_as_float(r[14]), __uint_as_float(r[15]));
            bad |= !isfinite(__low2float(*reinterpret_cast<const __half2*>(&o0.x)));
            *reinterpret_cast<uint4*>(smem + kBwdDZo + tile_off(row, 0, 16)) = o0;
            *reinterpret_cast<uint4*>(smem + kBwdDZo + tile_off(row, 8, 16)) = o1;
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads(); MFN_BTS();
        // ---- stage D: dH1 = dZ2 . W2 ;  dW2^T += H1^T . dZ2
        if (tid == 0) {
            tc_fence_after();
            mma_f16_ss(tbase + kAccH, desc_kmajor(sDZo, 16, 0), desc_mnmajor(sbase + kW2, 64, 0), idesc_f16(128, 64, false, true), 0u);
            const uint32_t idw = idesc_f16(64, 16, true, true);
#pragma unroll
            for (int k0 = 0; k0 < kFT; k0 += 16) mma_f16_ss(tbase + kAccW2, desc_mnmajor(sH1, 64, k0), desc_mnmajor(sDZo, 16, k0), idw, acc | (k0 > 0));
            mma_commit(&bar_mma);
        }
        mbar_wait(&bar_mma, ph_mma); ph_mma ^= 1u; MFN_BTS();
        tc_fence_after();
        mask_epilogue32(trow + kAccH, sBlob + kBH1, row, 32 * hsel);        // dZ1 in place of H1
        fence_async_smem();
        tc_fence_before();
        __syncthreads(); MFN_BTS();
        // ---- stage E: dX = dZ1 . W1 (32 columns) ;  dW1 += dZ1^T . X
        if (tid == 0) {
            tc_fence_after();
            const uint32_t id = idesc_f16(128, 32, false, true);
#pragma unroll
            for (int k0 = 0; k0 < 64; k0 += 16) mma_f16_ss(tbase + kAccH, desc_kmajor(sH1, 64, k0), desc_mnmajor(sbase + kW1, 32, k0), id, k0 > 0);
            const uint32_t idw = idesc_f16(64, 32, true, true);
#pragma unroll
            for (int k0 = 0; k0 < kFT; k0 += 16) mma_f16_ss(tbase + kAccW1, desc_mnmajor(sH1, 64, k0), desc_mnmajor(sX, 32, k0), idw, acc | (k0 > 0));
            mma_commit(&bar_mma);
        }
        mbar_wait(&bar_mma, ph_mma); ph_mma ^= 1u; MFN_BTS();
        tc_fence_after();
        {   // dX row -> dfeats, level-major [16][stride] half2: coalesced here and in the scatter kernel; each thread of a row does 8 levels
            uint32_t r[16];
            tmem_ld_x16(trow + kAccH + 16 * hsel, r);
            tmem_ld_wait();
            if (valid) {
                uint32_t* dst = reinterpret_cast<uint32_t*>(a.dfeats) + i + (size_t)(8 * hsel) * a.dfeats_stride;
                bool ovf = false;
#pragma unroll
                for (int l = 0; l < 8; ++l) {
                    const uint32_t h2 = pack2(__uint_as_float(r[2 * l]), __uint_as_float(r[2 * l + 1]));
                    ovf |= ((h2 & 0x7c00u) == 0x7c00u) || ((h2 & 0x7c000000u) == 0x7c000000u);   // inf / nan in either half
                    dst[(size_t)l * a.dfeats_stride] = h2;
                }
                bad |= ovf;
            }
        }
        acc = 1u;
        fence_async_smem();   // generic reads/writes of the tile area are ordered before the next bulk copy into it
        tc_fence_before();
        __syncthreads(); MFN_BTS();
    }
    if (bad && a.overflow) *a.overflow = 1;
    // ---- flush this CTA's weight-gradient accumulators (M = 64 accumulators: row m lives in TMEM lane 32*(m/16) + m%16)
    float* part = a.partials + (size_t)blockIdx.x * kNumWg;
    if (acc == 0u) {
        for (int q = tid; q < kNumWg; q += kBwdThreads) part[q] = 0.f;
    } else if (warp < 4) {
        tc_fence_after();
        const int mrow = warp * 16 + (lane & 15);
        const bool own = lane < 16;
        constexpr int oW1 = 0, oW2 = 64 * 32, oW3 = oW2 + 16 * 64, oW4 = oW3 + 64 * 32, oW5 = oW4 + (NH2 - 1) * 64 * 64;
        {
            uint32_t r[32];
            tmem_ld_x32(trow + kAccW1, r); tmem_ld_wait();          // dW1[out = m][in = j]
            if (own) for (int j = 0; j < 32; ++j) part[oW1 + mrow * 32 + j] = __uint_as_float(r[j]);
            tmem_ld_x32(trow + kAccW3, r); tmem_ld_wait();          // dW3[out = m][in = j]
            if (own) for (int j = 0; j < 32; ++j) part[oW3 + mrow * 32 + j] = __uint_as_float(r[j]);
            if (NH2 == 2) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {                          // dW4^T[in = m][out = j]
                    tmem_ld_x32(trow + kAccW4 + 32 * h, r); tmem_ld_wait();
                    if (own) for (int j = 0; j < 32; ++j) part[oW4 + (32 * h + j) * 64 + mrow] = __uint_as_float(r[j]);
                }
            }
        }
        {
            uint32_t r[16];
            tmem_ld_x16(trow + kAccW2, r); tmem_ld_wait();          // dW2^T[in = m][out = j]
            if (own) for (int j = 0; j < 16; ++j) part[oW2 + j * 64 + mrow] = __uint_as_float(r[j]);
            tmem_ld_x16(trow + kAccW5, r); tmem_ld_wait();          // dW5^T[in = m][out = j]
            if (own) for (int j = 0; j < 16; ++j) part[oW5 + j * 64 + mrow] = __uint_as_float(r[j]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, kBwdCols);
}

// ------------------------------------------------------------------------------------------------------------------ host side
static int g_num_sms = 0;
static int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0, v = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        g_num_sms = v > 0 ? v : kNumSMs;
    }
    return g_num_sms;
}

bool fused_field_supported(const mfn_field_cfg* c) {
    return (c->grid.grid_type == MFN_GRID_HASH || c->grid.grid_type == MFN_GRID_MIXED) && c->grid.n_levels == 16 && c->grid.n_features == 2 && c->sigma_width == 64 && c->sigma_hidden == 1 && c->rgb_width == 64 &&
           (c->rgb_hidden == 1 || c->rgb_hidden == 2);
}
int fused_bwd_max_ctas() { return 2 * num_sms(); }
// MFN_FIELD_SAVE=full: round 1's behaviour (all five activation tiles stored, 64 KiB per tile) for A/B measurements
bool fused_save_full() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("MFN_FIELD_SAVE"); v = (e && e[0] == 'f') ? 1 : 0; }
    return v == 1;
}
size_t fused_blob_bytes(int64_t n_max) { return (size_t)ceil_div(n_max, kFT) * (fused_save_full() ? kBlob : kBlobX); }
size_t fused_partial_bytes() { return (size_t)fused_bwd_max_ctas() * kNumWg * sizeof(float); }

template <int NH2, int MODE, bool MIXED>
static void launch_fwd_m(const FusedArgs& a, const GridMeta& m, cudaStream_t st) {
    constexpr int smem_bytes = ((MODE == 1 || MODE == 4) && !kTrainAlias) ? kFwdSmem : kFwdX + kFT * 64 * 2;      // one shared tile region
    constexpr int max_ctas = ((MODE == 1 || MODE == 4) && !kTrainAlias) ? 4 : 5;
    static bool once = (cudaFuncSetAttribute(field_fwd_fused_kernel<NH2, MODE, MIXED>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes), true);
    (void)once;
    const int64_t tiles = ceil_div(a.n_max, kFT);
    static int ctas_per_sm = 0;
    if (ctas_per_sm == 0) { const char* e = getenv("MFN_FWD_CTAS"); ctas_per_sm = e ? atoi(e) : max_ctas; if (ctas_per_sm < 1 || ctas_per_sm > max_ctas) ctas_per_sm = max_ctas; }
    const int64_t cap = ctas_per_sm * (int64_t)num_sms();
    field_fwd_fused_kernel<NH2, MODE, MIXED><<<(unsigned)(tiles < cap ? tiles : cap), kFwdThreads, smem_bytes, st>>>(a, m);
}
template <int NH2, int MODE>
static void launch_fwd(const FusedArgs& a, const GridMeta& m, cudaStream_t st) {
    if (m.mixed) launch_fwd_m<NH2, MODE, true>(a, m, st); else launch_fwd_m<NH2, MODE, false>(a, m, st);
}

// mode: 0 inference, 1 training, 2 density only, 3 raw 16 outputs of the sigma network
int fused_field_forward(const FusedArgs& a, const GridMeta& m, int rgb_hidden, int mode, cudaStream_t st) {
    ProfScope ps(mode >= 2 ? "density_fwd" : "field_fwd", st);
    if (mode == 3) { launch_fwd<1, 3>(a, m, st); return check_launch("mfn_geo_fwd(fused)", st); }
    if (mode == 1 && fused_save_full()) mode = 4;
    if (rgb_hidden == 2) {
        if (mode == 0) launch_fwd<2, 0>(a, m, st); else if (mode == 1) launch_fwd<2, 1>(a, m, st); else if (mode == 4) launch_fwd<2, 4>(a, m, st); else launch_fwd<2, 2>(a, m, st);
    } else {
        if (mode == 0) launch_fwd<1, 0>(a, m, st); else if (mode == 1) launch_fwd<1, 1>(a, m, st); else if (mode == 4) launch_fwd<1, 4>(a, m, st); else launch_fwd<1, 2>(a, m, st);
    }
    return check_launch("mfn_field_fwd(fused)", st);
}

template <int NH2, bool RC>
static int launch_bwd(const FusedArgs& a, cudaStream_t st) {
    static bool once = (cudaFuncSetAttribute(field_bwd_fused_kernel<NH2, RC>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem), true);
    (void)once;
    const int64_t tiles = ceil_div(a.n_max, kFT);
    const int64_t cap = fused_bwd_max_ctas();
    const unsigned grid = (unsigned)(tiles < cap ? tiles : cap);
    field_bwd_fused_kernel<NH2, RC><<<grid, kBwdThreads, kBwdSmem, st>>>(a);
    return (int)grid;
}

int fused_field_backward(const FusedArgs& a, int rgb_hidden, float* d_sigma_params, float* d_rgb_params, WgradReduce* wr, cudaStream_t st) {
    int grid;
    {
        ProfScope ps("field_bwd", st);
        if (fused_save_full()) grid = rgb_hidden == 2 ? launch_bwd<2, false>(a, st) : launch_bwd<1, false>(a, st);
        else grid = rgb_hidden == 2 ? launch_bwd<2, true>(a, st) : launch_bwd<1, true>(a, st);
    }
    wr->partials = a.partials; wr->n_parts = grid; wr->stride = kNumWg;
    wr->n_sigma = 3072; wr->n_rgb = 64 * 32 + (rgb_hidden - 1) * 64 * 64 + 16 * 64;
    wr->d_sigma = d_sigma_params; wr->d_rgb = d_rgb_params;
    return check_launch("mfn_field_bwd(fused)", st);
}

}  // namespace mfn
