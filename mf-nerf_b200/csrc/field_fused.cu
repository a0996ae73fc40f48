// Fused NGP field kernels for sm_100a: hash-grid gather + sigma MLP + SH + rgb MLP in ONE kernel (forward), and the whole MLP
// backward (recomputed activations + dgrad chain + weight gradients) in ONE kernel, on the 5th-generation tensor cores
// (tcgen05.mma, accumulators in TMEM).  Shapes: L*F = 32 encoded features with F = 2 (plain hash grid or the fork's MixedFeature
// grid), 64-wide sigma net with one hidden layer, rgb net 64 OR 128 wide (RW) with 1 or 2 hidden layers (NH2) -- i.e. the
// reference's defaults (opt.py) and MF-NeRF's own scripts (--grid MixedFeature --N_tables 8 --rgb_channels 128 --rgb_layers 2,
// benchmarking/benchmark_*_mf.sh).  Semantics: models/networks.py:96-155 (NGP.density / NGP.forward), TruncExp
// custom_functions.py:162-173.
//
// Forward, one CTA = 128 samples, 256 threads (two threads per sample row; row r == TMEM lane r), persistent over tiles,
// 5 CTAs / SM (RW = 64: 36 KiB of shared memory, 64 TMEM columns, 48 registers) or 2 CTAs / SM (RW = 128: 82 KiB, 128 columns):
//   gather 16 levels x 8 corners (lane pairs, fp32 interpolation) -> X tile [128x32] fp16 in shared memory (canonical un-swizzled
//   core-matrix layout, umma.cuh) -> tcgen05.mma X.W1^T -> TMEM -> ReLU -> H1 tile -> mma H1.W2^T -> h (16) -> sigma = exp(h0);
//   CAT tile = [SH4(dir) | h] -> mma CAT.W3^T -> ReLU -> mma .W4^T -> ReLU -> mma .W5^T -> sigmoid -> rgb.
//   X, H and CAT live one after the other in ONE tile region and the accumulators share the TMEM columns.
//   Nothing but xyz/dir in and sigma/rgb out touches HBM at inference.  In training the forward pass saves, per sample, only what
//   the backward pass cannot cheaply recompute: the encoded features X (64 B, the 8 KiB X tile *in the shared-memory layout*, one
//   bulk async store per tile), the view direction (12 B), the fp16 rgb outputs (8 B) and the normalised position for the scatter
//   kernel (16 B) -- 100 B/sample instead of round 1's 540 B/sample of activation blobs.  The hidden activations H1 / CAT / H2 /
//   H3 never leave the SM: the backward kernel re-runs the four forward MMAs on the saved X tile (same instruction shapes, same
//   operands -> bit-identical activations) before its dgrad / wgrad chain.  (A/B and traffic numbers: DESIGN.md section 4.)
// Backward, one CTA = 128 samples, 256 threads, persistent, 2 CTAs / SM (RW = 64: 96 KiB, 256 TMEM columns) or 1 CTA / SM (RW = 128):
//   X tile of the NEXT tile prefetched with one 8 KiB bulk async copy while this one is processed (double buffer) ->
//   recompute H1, h, CAT = [SH(dir) | h], H2, H3 ->
//   dZ5 = dL/drgb * sigmoid' -> [dgrad mma -> TMEM -> ReLU mask -> dZ tile (in place of the activation it masks)] x 4 -> dX;
//   the weight gradients dW = dZ^T.A are tcgen05 MMAs with M = 64 / 128, both operands read MN-major from the very same tiles, and
//   they ACCUMULATE IN TMEM across all tiles of the CTA (fp32); one partial per CTA is written at the end and an extra grid row of
//   the scatter kernel (encoder.cu) sums the partials (deterministic, no atomics).  dX (fp16, loss-scaled, level-major) goes to the
//   hash-grid scatter kernel.  The MMA-issuing thread advances its shared-memory descriptors by constant increments (one uniform
//   add per operand per MMA instead of rebuilding the 64-bit descriptor: the issue path is what a stage waits on).
#include "field_internal.h"
#include "grid_common.cuh"
#include "sh4.cuh"
#include "umma.cuh"
#include <stdlib.h>
#include <atomic>

namespace mfn {
using namespace umma;

constexpr int kFT = 128;                 // samples per tile
constexpr int kBlobX = kFT * 32 * 2;     // saved per tile for the backward pass: the X tile

// shared-memory byte offsets, TMEM columns and partial sizes as a function of the rgb net's width
template <int RW>
struct Lay {
    static constexpr int HW = RW > 64 ? RW : 64;                 // widest hidden tile (columns)
    // weights first, all tiles in the canonical row-core layout
    static constexpr int W1 = 0;                                 // [64 x 32]
    static constexpr int W2 = W1 + 64 * 32 * 2;                  // [16 x 64]
    static constexpr int W3 = W2 + 16 * 64 * 2;                  // [RW x 32]
    static constexpr int W4 = W3 + RW * 32 * 2;                  // [RW x RW]
    static constexpr int W5 = W4 + RW * RW * 2;                  // [16 x RW]
    static constexpr int WEnd = W5 + 16 * RW * 2;                // 20480 (RW = 64) / 51200 (RW = 128)
    // forward kernel: weights | ONE tile region shared by X, H1, CAT, H2, H3
    static constexpr int FwdT = WEnd, FwdSmem = FwdT + kFT * HW * 2;
    static constexpr int FwdCols = HW;                           // TMEM columns (64 / 128)
    static constexpr int FwdCtas = RW == 64 ? 5 : 2;
    // backward kernel: weights | X (two buffers) | H1 | CAT | H2 | H3 | dZo [128 x 16]
    static constexpr int BX = WEnd, BH1 = BX + 2 * kBlobX, BC = BH1 + kFT * 64 * 2, BH2 = BC + kFT * 32 * 2, BH3 = BH2 + kFT * RW * 2,
                         BDZo = BH3 + kFT * RW * 2, BwdSmem = BDZo + kFT * 16 * 2;
    static constexpr int AccH = 0, AccW5 = HW, AccW4 = AccW5 + 16, AccW3 = AccW4 + RW, AccW2 = AccW3 + 32, AccW1 = AccW2 + 16, AccO2 = AccW1 + 32,
                         Used = AccO2 + 16;
    static constexpr int BwdCols = Used <= 256 ? 256 : 512;
    static constexpr int BwdCtas = RW == 64 ? 2 : 1;
    static constexpr int NumWg = 64 * 32 + 16 * 64 + RW * 32 + RW * RW + 16 * RW;   // weight-gradient floats per partial (10240 / 25600)
};
static_assert(Lay<64>::WEnd == 20480 && Lay<64>::Used == 240 && Lay<64>::BwdSmem == 98304, "layout (rgb width 64)");
static_assert(Lay<128>::WEnd == 51200 && Lay<128>::Used == 368 && Lay<128>::BwdSmem <= 227 * 1024, "layout (rgb width 128)");

// global row-major [rows][cols] fp16 matrix -> row-core tile in shared memory
__device__ __forceinline__ void stage_weight(unsigned char* dst, const __half* __restrict__ src, int rows, int cols, int tid, int nthreads) {
    const int cpr = cols >> 3;
    for (int q = tid; q < rows * cpr; q += nthreads) {
        const int r = q / cpr, c = (q % cpr) << 3;
        *reinterpret_cast<uint4*>(dst + tile_off(r, c, cols)) = __ldg(reinterpret_cast<const uint4*>(src + (size_t)r * cols + c));
    }
}

template <int NH2, int RW>
__device__ __forceinline__ void stage_all_weights(unsigned char* smem, const FusedArgs& a, int tid, int nthreads, bool rgb) {
    using L = Lay<RW>;
    stage_weight(smem + L::W1, a.w_sigma, 64, 32, tid, nthreads);
    stage_weight(smem + L::W2, a.w_sigma + 64 * 32, 16, 64, tid, nthreads);
    if (rgb) {
        stage_weight(smem + L::W3, a.w_rgb, RW, 32, tid, nthreads);
        if (NH2 == 2) stage_weight(smem + L::W4, a.w_rgb + RW * 32, RW, RW, tid, nthreads);
        stage_weight(smem + L::W5, a.w_rgb + RW * 32 + (NH2 - 1) * RW * RW, 16, RW, tid, nthreads);
    }
}

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// Lane-pair gather of one level, F = 2: the two lanes of a pair work on the same sample and each fetches the 4 corners of ITS x
// (gx + xb), so the corners (x, x+1) of one (y, z) -- same 32-byte sector 7 times out of 8 -- are requested by the same load
// instruction: a warp's gather touches <= 16 sectors instead of <= 32, and the L1 sector (tag) rate is what bounds this phase
// (profiles/l2_peaks_r02.json: 1 divergent sector per SM per clock).  fp32 trilinear interpolation; returns the pair's sum.
// Instruction diet (the phase is issue-bound once 5 CTAs per SM hide the latency): floor() without the conversion pipe -- for
// |p| < 2^22, t = RZ(p + 1.5 * 2^23) is an integer-valued float in [2^23, 2^24) whose low mantissa bits ARE floor(p), so
// floor(p) = bits(t) - 0x4B400000 (one integer add) and (float)floor(p) = t - 1.5 * 2^23 (exact): two fp32-pipe operations
// instead of FRND + F2I on the quarter-rate pipe, same values bit for bit; 32-bit entry indices (offset + index < 2^32) so that
// an address is one IMAD.WIDE.
__device__ __forceinline__ float floor_split(float p, int& i) {
    const float t = __fadd_rz(p, 12582912.f);
    i = __float_as_int(t) - 0x4B400000;
    return __fsub_rn(t, 12582912.f);
}
// one level in two halves so that the caller can put the loads of SEVERAL levels in flight before consuming any of them (the
// phase is a chain of L2 round trips per thread: its length is what a tile's gather costs)
struct LevelW { float wx, wy, wz; };
template <bool MIXED>
__device__ __forceinline__ LevelW level_indices(const GridMeta& m, int l, float x, float y, float z, int xb, uint32_t (&idx)[4]) {
    const float s = m.scale[l];
    const uint32_t off = m.offset[l], size = m.size[l];
    const float px = fmaf(x, s, 0.5f), py = fmaf(y, s, 0.5f), pz = fmaf(z, s, 0.5f);
    int ix, iy, iz;
    const float fx = floor_split(px, ix), fy = floor_split(py, iy), fz = floor_split(pz, iz);
    LevelW w;
    w.wx = px - fx; w.wy = py - fy; w.wz = pz - fz;
    const uint32_t cx = (uint32_t)ix + (uint32_t)xb, gy = (uint32_t)iy, gz = (uint32_t)iz;
    if (MIXED) {   // MixedFeature grid: vertices are hashed by their canonical-grid coordinates (grid_common.cuh: canon_vertex); the
                   // lane pair still splits the x corners, it just no longer finds them in one sector
        const float r = m.canon[l];
        const uint32_t mask = size - 1u, ccx = canon_vertex(cx, r);
        const uint32_t hy0 = canon_vertex(gy, r) * 2654435761u, hy1 = canon_vertex(gy + 1u, r) * 2654435761u;
        const uint32_t hz0 = canon_vertex(gz, r) * 805459861u, hz1 = canon_vertex(gz + 1u, r) * 805459861u;
        idx[0] = (ccx ^ hy0 ^ hz0) & mask; idx[1] = (ccx ^ hy1 ^ hz0) & mask; idx[2] = (ccx ^ hy0 ^ hz1) & mask; idx[3] = (ccx ^ hy1 ^ hz1) & mask;
    } else if ((m.hashed >> l) & 1u) {       // size is a power of two for hashed levels
        const uint32_t mask = size - 1u;
        const uint32_t hy0 = gy * 2654435761u, hy1 = hy0 + 2654435761u, hz0 = gz * 805459861u, hz1 = hz0 + 805459861u;
        idx[0] = (cx ^ hy0 ^ hz0) & mask; idx[1] = (cx ^ hy1 ^ hz0) & mask; idx[2] = (cx ^ hy0 ^ hz1) & mask; idx[3] = (cx ^ hy1 ^ hz1) & mask;
    } else {                                  // dense: x + y*res + z*res^2, mod size (only a corner on the x/y/z == res border wraps)
        const uint32_t res = m.res[l], r2 = res * res;
        const uint32_t b = cx + gy * res + gz * r2;
        idx[0] = b; idx[1] = b + res; idx[2] = b + r2; idx[3] = b + res + r2;
        if (idx[3] >= size) {
#pragma unroll
            for (int c = 0; c < 4; ++c) idx[c] %= size;
        }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) idx[c] += off;
    return w;
}
// fp32 trilinear interpolation of this lane's 4 corners, summed over the lane pair
__device__ __forceinline__ uint32_t level_interp(const uint32_t (&v)[4], const LevelW& lw, int xb) {
    const float wxs = xb ? lw.wx : 1.f - lw.wx;
    const float a0 = wxs * (1.f - lw.wy), a1 = wxs * lw.wy;
    const float w[4] = {a0 * (1.f - lw.wz), a1 * (1.f - lw.wz), a0 * lw.wz, a1 * lw.wz};
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
template <bool MIXED>
__device__ __forceinline__ uint32_t gather_level_pair(const uint32_t* __restrict__ table, const GridMeta& m, int l, float x, float y, float z, int xb, uint64_t pol) {
    uint32_t idx[4], v[4];
    const LevelW lw = level_indices<MIXED>(m, l, x, y, z, xb, idx);
#pragma unroll
    for (int c = 0; c < 4; ++c) v[c] = ldg_nc_hint(table + idx[c], pol);   // the table is to stay L2-resident (evict_last)
    return level_interp(v, lw, xb);
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

// 16 fp32 TMEM values -> 16 fp16 (two uint4)
__device__ __forceinline__ void pack16(const uint32_t (&r)[16], uint4& o0, uint4& o1) {
    o0.x = pack2(__uint_as_float(r[0]), __uint_as_float(r[1])); o0.y = pack2(__uint_as_float(r[2]), __uint_as_float(r[3]));
    o0.z = pack2(__uint_as_float(r[4]), __uint_as_float(r[5])); o0.w = pack2(__uint_as_float(r[6]), __uint_as_float(r[7]));
    o1.x = pack2(__uint_as_float(r[8]), __uint_as_float(r[9])); o1.y = pack2(__uint_as_float(r[10]), __uint_as_float(r[11]));
    o1.z = pack2(__uint_as_float(r[12]), __uint_as_float(r[13])); o1.w = pack2(__uint_as_float(r[14]), __uint_as_float(r[15]));
}

__device__ __forceinline__ uint32_t relu2(uint32_t packed) {
    const __half2 z = __float2half2_rn(0.f);
    const __half2 v = __hmax2(*reinterpret_cast<const __half2*>(&packed), z);
    return *reinterpret_cast<const uint32_t*>(&v);
}
// TMEM (32 fp32 columns starting at col0) -> ReLU -> fp16 -> columns col0..col0+31 of row `row` of a [128 x C] row-core tile
// (16 columns per tcgen05.ld: the 64-wide forward kernel lives on 48 registers to fit 5 CTAs per SM)
template <int C>
__device__ __forceinline__ void relu_epilogue32(uint32_t taddr, unsigned char* tile, int row, int col0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        uint32_t r[16];
        tmem_ld_x16(taddr + col0 + 16 * h, r);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 2; ++c) {      // round first, clamp the PAIR afterwards: fp16(max(x, 0)) == max(fp16(x), 0) (rounding is monotone, max(-0, +0) = +0)
            uint4 o;
            o.x = relu2(pack2(__uint_as_float(r[8 * c + 0]), __uint_as_float(r[8 * c + 1])));
            o.y = relu2(pack2(__uint_as_float(r[8 * c + 2]), __uint_as_float(r[8 * c + 3])));
            o.z = relu2(pack2(__uint_as_float(r[8 * c + 4]), __uint_as_float(r[8 * c + 5])));
            o.w = relu2(pack2(__uint_as_float(r[8 * c + 6]), __uint_as_float(r[8 * c + 7])));
            *reinterpret_cast<uint4*>(tile + tile_off(row, col0 + 16 * h + c * 8, C)) = o;
        }
    }
}
// a hidden layer of width C: thread `hsel` of a row does columns [hsel * C/2, (hsel + 1) * C/2)
template <int C>
__device__ __forceinline__ void relu_epilogue(uint32_t taddr, unsigned char* tile, int row, int hsel) {
#pragma unroll
    for (int q = 0; q < C / 64; ++q) relu_epilogue32<C>(taddr, tile, row, hsel * (C / 2) + 32 * q);
}

// MMA issue with incrementally advanced descriptors: the start-address field (bits 0-13, in 16-byte units) never carries into the
// next field for shared-memory addresses below 256 KiB, so stepping along K is one 32-bit add on the low word.
struct DescIter {
    uint32_t lo, hi;
    __device__ __forceinline__ explicit DescIter(uint64_t d) : lo((uint32_t)d), hi((uint32_t)(d >> 32)) {}
    __device__ __forceinline__ uint64_t get() const { return ((uint64_t)hi << 32) | lo; }
    __device__ __forceinline__ void step(uint32_t bytes) { lo += bytes >> 4; }
};
// D[tmem] (+)= A . B over `n` K-steps of 16; a_step / b_step = byte distance of consecutive K-steps inside the operand tiles
__device__ __forceinline__ void mma_chain(uint32_t d_tmem, uint64_t a_desc, uint32_t a_step, uint64_t b_desc, uint32_t b_step, uint32_t idesc, int n,
                                          uint32_t acc_first) {
    DescIter a(a_desc), b(b_desc);
#pragma unroll
    for (int k = 0; k < n; ++k) {
        mma_f16_ss(d_tmem, a.get(), b.get(), idesc, k > 0 ? 1u : acc_first);
        a.step(a_step); b.step(b_step);
    }
}
// K-major operand: 16 K-elements = 2 core matrices of 128 bytes; MN-major operand with C columns: 16 K-rows = 2 row groups of C/8 core matrices
__host__ __device__ constexpr uint32_t kstep_kmajor() { return 256u; }
__host__ __device__ constexpr uint32_t kstep_mnmajor(int C) { return 2u * (uint32_t)(C >> 3) * 128u; }

// ------------------------------------------------------------------------------------------------------------------ forward
// MODE 0: inference (sigma + rgb), 1: training (also saves X tile, direction, fp16 rgb and normalised position), 2: density only (sigma),
// 3: the 16 raw outputs of the grid + sigma MLP in fp16 (tcnn.NetworkWithInputEncoding's forward, networks.py:107),
// 4: inference inside the render wavefront (render.cu): row i = sample i % N_samples of alive slot i / N_samples; x = o + t d and the
//    direction come from the ray (the marcher stores t and dt only: 8 B instead of 32 B per row), rows past N_eff are skipped.
// 256 threads: thread t works on sample row t % 128 (= its TMEM lane); the two threads of a row split the 16 grid levels in the
// gather phase and the accumulator columns in the hidden-layer epilogues.
// scheduler slots: {next tile, CTAs that have left}; zero at module load and zero again after every launch (the last leaver resets)
constexpr int kSchedSlots = 64;
__device__ unsigned int g_tile_sched[kSchedSlots][2];
__device__ __forceinline__ void sched_leave(unsigned int* sched) {
    if (!sched) return;
    __threadfence();
    if (atomicInc(sched + 1, gridDim.x - 1u) == gridDim.x - 1u) { sched[0] = 0u; __threadfence(); }   // atomicInc wraps the leaver count to 0 itself
}
constexpr int kFwdThreads = 256;
#ifndef MFN_GATHER_TWO
#define MFN_GATHER_TWO 0
#endif
constexpr bool kGatherTwoLevels = MFN_GATHER_TWO != 0;
// phase timestamps of CTA 0 (tools/fwd_phases.py): a.dbg != nullptr only in that tool
#define MFN_TS(k) do { if (a.dbg && blockIdx.x == 0 && tid == (k >= 100 ? 255 : 0) && tile_no < 12) a.dbg[tile_no * 16 + (k % 100)] = clock64(); } while (0)

// kernel entry / exit of every CTA (tools/fwd_phases.py): dbg[1024 + 4 * cta + {0: entry ns, 1: exit ns, 2: SM id, 3: tiles done}]
__device__ __forceinline__ long long global_ns() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ uint32_t sm_id() { uint32_t v; asm volatile("mov.u32 %0, %%smid;" : "=r"(v)); return v; }
#define MFN_EDGE_IN() do { if (a.dbg && tid == 0) { long long* e_ = a.dbg + 1024 + 4 * blockIdx.x; e_[0] = global_ns(); e_[2] = sm_id(); } } while (0)
#define MFN_EDGE_OUT() do { if (a.dbg && tid == 0) { long long* e_ = a.dbg + 1024 + 4 * blockIdx.x; e_[1] = global_ns(); e_[3] = tile_no; } } while (0)

template <int NH2, int MODE, bool MIXED, int RW>
__global__ void __launch_bounds__(kFwdThreads, Lay<RW>::FwdCtas)
field_fwd_fused_kernel(const __grid_constant__ FusedArgs a, const __grid_constant__ GridMeta m) {
    using L = Lay<RW>;
    constexpr bool TRAIN = (MODE == 1), RAYS = (MODE == 4), RGB = (MODE < 2) || RAYS;
    // No tile is alive across stages, so X, H1, CAT, H2 and H3 share ONE tile region and the 16-column output accumulator shares
    // the hidden accumulator's TMEM columns (the gather is bound by resident parallelism: every KiB counts).  In training mode the
    // X tile leaves with a bulk async store that is waited for (its shared-memory read) before H1 overwrites the region.
    constexpr int oT = L::FwdT;
    constexpr int nCols = L::FwdCols;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int row = tid & (kFT - 1), hsel = tid >> 7;
    const int64_t n = a.n_dev ? min((int64_t)*a.n_dev, a.n_max) : a.n_max;
    const int64_t n_tiles = (n + kFT - 1) / kFT;
    if (TRAIN && a.n_out && blockIdx.x == 0 && tid == 0) *a.n_out = (int32_t)n;     // the backward pass's own copy of the count
    // Dynamic tile scheduler: tiles are handed out by a global counter (a.sched[0]) instead of the static stride blockIdx.x + k * gridDim.x.
    // With 5 CTAs per SM and 6.5 tiles per CTA the static split left an SM with 4, 3, ... resident CTAs for the last tile time
    // (CTAs with 6 tiles done, CTAs with 7 not: measured 115 us of SM time for 97 us of work); the next tile is requested one tile
    // ahead, so the atomic's latency is never waited for.  The last CTA to leave resets the counter (a.sched[1] counts leavers).
    __shared__ int next_tile_s;
    if (tid == 0) next_tile_s = a.sched ? (int)atomicAdd(a.sched, 1u) : (int)blockIdx.x;
    __syncthreads();
    if (next_tile_s >= n_tiles) {      // the grid is sized for n_max; with a device-side count most CTAs may have nothing to do
        if (tid == 0) sched_leave(a.sched);
        return;
    }
    MFN_EDGE_IN();      // the grid is sized for n_max; with a device-side count most CTAs may have nothing to do
    stage_all_weights<NH2, RW>(smem, a, tid, kFwdThreads, RGB);
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

    // render mode: N_samples of this iteration (uniform) and the magic multiplier of the division row -> alive slot
    // (exact for rows < 2^26 and N_samples <= 64: the error of umulhi(i, ceil(2^32 / ns)) stays below 1 / 64)
    uint32_t r_ns = 1u, r_magic = 0u;
    const int32_t* r_alive = nullptr;
    if (RAYS) {
        r_ns = (uint32_t)max(a.ray_plan[3], 1);
        r_magic = (uint32_t)((0x100000000ull + r_ns - 1u) / r_ns);
        r_alive = a.ray_alive_lists + (size_t)a.ray_plan[2] * a.ray_list_stride;
    }
    int tile_no = 0;
    for (int64_t tile = next_tile_s; tile < n_tiles; tile = next_tile_s, ++tile_no) {
        const int64_t i = tile * kFT + row;
        bool valid = i < n;
        int ray_of_row = 0;
        if (RAYS && valid) {
            const uint32_t slot = r_ns == 1u ? (uint32_t)i : __umulhi((uint32_t)i, r_magic);
            valid = (int)((uint32_t)i - slot * r_ns) < a.ray_n_eff[slot];
            ray_of_row = r_alive[slot];
        }
        int tile_after = (int)tile + (int)gridDim.x;
        if (a.sched && tid == 0) tile_after = (int)atomicAdd(a.sched, 1u);    // consumed at the end of this tile
        MFN_TS(0);
        // ---- hash-grid gather -> X tile: lane pair (2p, 2p+1) of warp w works on row 16w + p, all 16 levels
        {
            const int grow = tid >> 1, xb = tid & 1;
            const int64_t gi = tile * kFT + grow;
            bool gvalid = gi < n;
            float x = 0.5f, y = 0.5f, z = 0.5f;
            if (RAYS && gvalid) {
                const uint32_t slot = r_ns == 1u ? (uint32_t)gi : __umulhi((uint32_t)gi, r_magic);
                gvalid = (int)((uint32_t)gi - slot * r_ns) < a.ray_n_eff[slot];
                if (gvalid) {      // the marcher's own expression (march.cuh: probe_cell), bit for bit
                    const int r = r_alive[slot];
                    const float t = a.ray_ts[gi];
                    x = __fmaf_rn(a.rays_d[3 * r], t, a.rays_o[3 * r]); y = __fmaf_rn(a.rays_d[3 * r + 1], t, a.rays_o[3 * r + 1]);
                    z = __fmaf_rn(a.rays_d[3 * r + 2], t, a.rays_o[3 * r + 2]);
                }
            }
            if (gvalid) {
                if (!RAYS) { x = a.xyzs[3 * gi]; y = a.xyzs[3 * gi + 1]; z = a.xyzs[3 * gi + 2]; }
                x = __fdiv_rn(__fsub_rn(x, a.mn[0]), __fsub_rn(a.mx[0], a.mn[0]));      // networks.py:105
                y = __fdiv_rn(__fsub_rn(y, a.mn[1]), __fsub_rn(a.mx[1], a.mn[1]));
                z = __fdiv_rn(__fsub_rn(z, a.mn[2]), __fsub_rn(a.mx[2], a.mn[2]));
                if (TRAIN && xb == 1) a.x01[gi] = make_float4(x, y, z, 0.f);
            }
            // feature pair of level l = columns 2l, 2l+1 of row `grow`: core matrix l / 4, byte 4 * (l % 4) of the row's 16-byte chunk;
            // lane xb of the pair stores the levels of its parity
            unsigned char* xrow = smem + oT + tile_off(grow, 0, 32) + 4 * xb;
            if (RAYS && !__any_sync(0xffffffffu, gvalid)) {      // 16 padding rows: nothing to gather
#pragma unroll
                for (int l = 0; l < 16; l += 2) *reinterpret_cast<uint32_t*>(xrow + (l >> 2) * 128 + (l & 2) * 4) = 0u;
            } else
#pragma unroll 2
            for (int l = 0; l < 16; l += 2) {
                uint32_t v0, v1;
                if (kGatherTwoLevels) {      // 8 loads in flight per thread (invalid rows gather entry 0 harmlessly)
                    uint32_t i0[4], i1[4];
                    const LevelW w0 = level_indices<MIXED>(m, l, x, y, z, xb, i0);
                    const LevelW w1 = level_indices<MIXED>(m, l + 1, x, y, z, xb, i1);
#pragma unroll
                    for (int c = 0; c < 4; ++c) i0[c] = ldg_nc_hint(table + i0[c], pol_keep);
#pragma unroll
                    for (int c = 0; c < 4; ++c) i1[c] = ldg_nc_hint(table + i1[c], pol_keep);
                    v0 = level_interp(i0, w0, xb);
                    v1 = level_interp(i1, w1, xb);
                } else {
                    v0 = gather_level_pair<MIXED>(table, m, l, x, y, z, xb, pol_keep);
                    v1 = gather_level_pair<MIXED>(table, m, l + 1, x, y, z, xb, pol_keep);
                }
                *reinterpret_cast<uint32_t*>(xrow + (l >> 2) * 128 + (l & 2) * 4) = gvalid ? (xb ? v1 : v0) : 0u;
            }
        }
        MFN_TS(1); MFN_TS(109);
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        MFN_TS(2);
        // the view direction of this row (partner thread), requested now so that the layer-1 MMA wait hides its latency: a load issued
        // just before a stage boundary is waited for by the boundary's proxy fence
        float dirx = 0.f, diry = 0.f, dirz = 1.f;
        if (RGB && hsel == 1 && valid) {
            const float* dp = RAYS ? a.rays_d + 3 * (size_t)ray_of_row : a.dirs + 3 * i;
            dirx = dp[0]; diry = dp[1]; dirz = dp[2];
        }
        // ---- layer 1: H1 = relu(X . W1^T)
        if (tid == 0) {
            tc_fence_after();
            mma_chain(tbase, desc_kmajor(sbase + oT, 32, 0), kstep_kmajor(), desc_kmajor(sbase + L::W1, 32, 0), kstep_kmajor(), idesc_f16(128, 64, false, false), 2, 0u);
            if (TRAIN) { bulk_s2g_hint(a.blobs + (size_t)tile * kBlobX, smem + oT, kBlobX, pol_stream); bulk_commit(); bulk_wait_read0(); }   // (H1 overwrites X)
            mma_commit(&bar);
        }
        mbar_wait(&bar, phase); phase ^= 1u;
        MFN_TS(3);
        tc_fence_after();
        relu_epilogue<64>(trow, smem + oT, row, hsel);
        MFN_TS(4);
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        MFN_TS(5);
        // ---- layer 2: h = H1 . W2^T (16 outputs, no activation); sigma = exp(h0)  (TruncExp forward)
        if (tid == 0) {
            tc_fence_after();
            mma_chain(tbase, desc_kmajor(sbase + oT, 64, 0), kstep_kmajor(), desc_kmajor(sbase + L::W2, 64, 0), kstep_kmajor(), idesc_f16(128, 16, false, false), 4, 0u);
            mma_commit(&bar);
        }
        if (hsel == 0) {
            mbar_wait(&bar, phase);
            tc_fence_after();
            uint32_t r[16];
            tmem_ld_x16(trow, r);
            tmem_ld_wait();
            uint4 o0, o1;
            pack16(r, o0, o1);
            if (MODE == 3) {
                if (valid) { uint4* ho = reinterpret_cast<uint4*>(a.h_out + 16 * i); ho[0] = o0; ho[1] = o1; }
            } else if (valid) a.sigmas[i] = expf(__low2float(*reinterpret_cast<const __half2*>(&o0.x)));
            if (RGB) {   // CAT may overlay H1, which the layer-2 MMA has finished reading
                *reinterpret_cast<uint4*>(smem + oT + tile_off(row, 16, 32)) = o0;
                *reinterpret_cast<uint4*>(smem + oT + tile_off(row, 24, 32)) = o1;
            }
        } else if (RGB) {
            // the partner thread of the row meanwhile encodes the direction: SH of the normalised direction -> CAT[:, 0:16]
            // (networks.py:145-146)
            mbar_wait(&bar, phase);
            uint4 o0 = make_uint4(0u, 0u, 0u, 0u), o1 = o0;
            if (valid) {
                if (TRAIN) { a.dirs_copy[3 * i] = dirx; a.dirs_copy[3 * i + 1] = diry; a.dirs_copy[3 * i + 2] = dirz; }   // the backward pass re-encodes it
                sh_of_dir(dirx, diry, dirz, o0, o1);
            }
            *reinterpret_cast<uint4*>(smem + oT + tile_off(row, 0, 32)) = o0;
            *reinterpret_cast<uint4*>(smem + oT + tile_off(row, 8, 32)) = o1;
        }
        phase ^= 1u;
        MFN_TS(6);
        if (!RGB && tid == 0) next_tile_s = tile_after;
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        MFN_TS(7);
        if (!RGB) continue;   // (uniform) density only / raw outputs
        // ---- rgb layer 1: H2 = relu(CAT . W3^T)
        if (tid == 0) {
            tc_fence_after();
            mma_chain(tbase, desc_kmajor(sbase + oT, 32, 0), kstep_kmajor(), desc_kmajor(sbase + L::W3, 32, 0), kstep_kmajor(), idesc_f16(128, RW, false, false), 2, 0u);
            mma_commit(&bar);
        }
        mbar_wait(&bar, phase); phase ^= 1u;
        tc_fence_after();
        relu_epilogue<RW>(trow, smem + oT, row, hsel);
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (NH2 == 2) {
            // ---- rgb layer 2: H3 = relu(H2 . W4^T)
            if (tid == 0) {
                tc_fence_after();
                mma_chain(tbase, desc_kmajor(sbase + oT, RW, 0), kstep_kmajor(), desc_kmajor(sbase + L::W4, RW, 0), kstep_kmajor(), idesc_f16(128, RW, false, false), RW / 16, 0u);
                mma_commit(&bar);
            }
            mbar_wait(&bar, phase); phase ^= 1u;
            tc_fence_after();
            relu_epilogue<RW>(trow, smem + oT, row, hsel);
            fence_async_smem();
            tc_fence_before();
            __syncthreads();
        }
        // ---- rgb output layer: rgb = act(H . W5^T)[0:3], rounded to fp16 like tcnn's output
        if (tid == 0) {
            tc_fence_after();
            mma_chain(tbase, desc_kmajor(sbase + oT, RW, 0), kstep_kmajor(), desc_kmajor(sbase + L::W5, RW, 0), kstep_kmajor(), idesc_f16(128, 16, false, false), RW / 16, 0u);
            mma_commit(&bar);
        }
        if (hsel == 0) {
            mbar_wait(&bar, phase);
            tc_fence_after();
            uint32_t r[8];
            tmem_ld_x8(trow, r);
            tmem_ld_wait();
            if (valid) {
                __align__(8) __half hv[4];
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
        if (tid == 0) next_tile_s = tile_after;
        tc_fence_before();
        __syncthreads();   // TMEM and the tile region are free for the next tile
        MFN_TS(8);
    }
    MFN_EDGE_OUT();
    if (tid == 0) sched_leave(a.sched);
    if (TRAIN && tid == 0) bulk_wait0();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, nCols);
}

// ------------------------------------------------------------------------------------------------------------------ backward
constexpr int kBwdThreads = 288;        // 8 epilogue warps (two threads per sample row) + 1 warp whose lane 0 issues every MMA and bulk copy:
                                        // its issue path never has a global load of its own outstanding (measured: the stage that followed the
                                        // per-row prefetch waited ~1.5 k cycles when the issuing thread was also a row thread)
constexpr int kBwdIssuer = 256;

// TMEM (fp32 dgrad accumulator) -> mask with relu'(activation tile row) -> fp16 dZ row written IN PLACE of the activation row.
// The activation tile is also an operand of the weight-gradient MMAs issued right behind the dgrad MMA (dW = A^T . dZ_prev), so the
// epilogue runs in two halves: TMEM load, mask and packing into registers as soon as the DGRAD accumulator is complete (bar_mma), the
// stores only once the weight-gradient MMAs have finished reading the tile (bar_dw) -- their issue and execution (8 MMAs) overlap the
// first half instead of sitting on the stage's critical path.
template <int C>
__device__ __forceinline__ void mask_epilogue(uint32_t taddr, unsigned char* tile, int row, int hsel, uint64_t* bar_dw, uint32_t& ph_dw) {
    constexpr int Q = C / 64;
    uint4 o[Q][4];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int col0 = hsel * (C / 2) + 32 * q;
        uint32_t r[32];
        tmem_ld_x32(taddr + col0, r);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint4 act = *reinterpret_cast<const uint4*>(tile + tile_off(row, col0 + c * 8, C));
            const uint32_t* aw = &act.x;
            uint32_t* ow = &o[q][c].x;
#pragma unroll
            for (int k = 0; k < 4; ++k) {      // the activations are ReLU outputs (>= +0): "> 0" is "bits != 0", no conversion needed
                ow[k] = pack2((aw[k] & 0xffffu) ? __uint_as_float(r[8 * c + 2 * k]) : 0.f, (aw[k] >> 16) ? __uint_as_float(r[8 * c + 2 * k + 1]) : 0.f);
            }
        }
    }
    mbar_wait(bar_dw, ph_dw); ph_dw ^= 1u;
#pragma unroll
    for (int q = 0; q < Q; ++q)
#pragma unroll
        for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(tile + tile_off(row, hsel * (C / 2) + 32 * q + c * 8, C)) = o[q][c];
}

// phase timestamps of CTA 0 / thread 0 of the backward kernel (tools/fwd_phases.py): after every MMA-completion wait and every barrier
#define MFN_BTS() do { if (a.dbg && blockIdx.x == 0 && tid == kBwdIssuer && tile_no < 6 && ts_k < 40) a.dbg[256 + tile_no * 40 + (ts_k++)] = clock64(); } while (0)
// one stage boundary: generic-proxy writes of the tiles -> async proxy (MMA), TMEM reads done, CTA barrier
#define MFN_STAGE_SYNC() do { fence_async_smem(); tc_fence_before(); __syncthreads(); MFN_BTS(); } while (0)
#define MFN_MMA_WAIT() do { mbar_wait(&bar_mma, ph_mma); ph_mma ^= 1u; MFN_BTS(); tc_fence_after(); } while (0)

template <int NH2, int RW>
__global__ void __launch_bounds__(kBwdThreads, Lay<RW>::BwdCtas)
field_bwd_fused_kernel(const __grid_constant__ FusedArgs a) {
    using L = Lay<RW>;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar_mma, bar_dw, bar_tail, bar_load[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = tid & (kFT - 1), hsel = tid >> 7;   // two threads per sample row: they split the accumulator columns (hsel == 2: issuer warp)
    const bool worker = tid < 256, issuer = tid == kBwdIssuer;
    const int64_t n = a.n_dev ? min((int64_t)*a.n_dev, a.n_max) : a.n_max;
    const int64_t n_tiles = (n + kFT - 1) / kFT;
    const uint64_t pol_stream = policy_evict_first();
    const float loss_scale = a.loss_scale_dev ? *a.loss_scale_dev : a.loss_scale;
    if (issuer) {
        mbar_init(&bar_mma, 1); mbar_init(&bar_dw, 1); mbar_init(&bar_tail, 1); mbar_init(&bar_load[0], 1); mbar_init(&bar_load[1], 1); mbar_fence_init();
        if ((int64_t)blockIdx.x < n_tiles) {      // first X tile: in flight while the weights are staged
            mbar_arrive_expect_tx(&bar_load[0], kBlobX);
            bulk_g2s_hint(smem + L::BX, a.blobs + (size_t)blockIdx.x * kBlobX, kBlobX, &bar_load[0], pol_stream);
        }
    }
    stage_all_weights<NH2, RW>(smem, a, tid, kBwdThreads, true);
    if (warp == 0) tmem_alloc(&tmem_base_s, L::BwdCols);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_base_s;
    const uint32_t trow = tmem_addr(tbase, (warp & 3) * 32, 0);
    const uint32_t sbase = smem_u32(smem);
    const uint32_t sH1 = sbase + L::BH1, sC = sbase + L::BC, sH2 = sbase + L::BH2, sH3 = sbase + L::BH3, sDZo = sbase + L::BDZo;
    const uint32_t sHL = (NH2 == 2) ? sH3 : sH2;                  // last hidden activation of the rgb net
    unsigned char* pHL = smem + ((NH2 == 2) ? L::BH3 : L::BH2);
    uint32_t ph_mma = 0, ph_dw = 0, ph_tail = 0, ph_load0 = 0, ph_load1 = 0;
    uint32_t acc = 0;                                              // 0 on the CTA's first tile: weight-gradient MMAs overwrite
    bool bad = false;
    // per-row inputs of a tile, requested one tile ahead (their latency used to open every tile: ~1.5 k of its ~12 k cycles):
    // thread 0 of a row: fp16 rgb outputs + dL/drgb; its partner: the view direction
    float pf[3] = {0.f, 0.f, 0.f};
    uint2 pfy = make_uint2(0u, 0u);
    auto prefetch_row = [&](int64_t t) {
        const int64_t ip = t * kFT + row;
        if (ip < n) {
            if (hsel == 0) { pfy = a.rgb_h[ip]; pf[0] = a.dL_drgbs[3 * ip]; pf[1] = a.dL_drgbs[3 * ip + 1]; pf[2] = a.dL_drgbs[3 * ip + 2]; }
            else { pf[0] = a.dirs_copy[3 * ip]; pf[1] = a.dirs_copy[3 * ip + 1]; pf[2] = a.dirs_copy[3 * ip + 2]; }
        }
    };
    if (worker && (int64_t)blockIdx.x < n_tiles) prefetch_row(blockIdx.x);

    int tile_no = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_no) {
        int ts_k = 0;
        MFN_BTS();
        const int64_t i = tile * kFT + row;
        const bool valid = worker && i < n;
        const int xbuf = tile_no & 1;
        const uint32_t sX = sbase + L::BX + xbuf * kBlobX;
        if (issuer && tile_no > 0) { mbar_wait(&bar_tail, ph_tail); ph_tail ^= 1u; }   // the previous tile's last MMAs (dW1) have read its X buffer
        if (issuer && tile + gridDim.x < n_tiles) {   // prefetch the NEXT tile's X into the other buffer = the previous tile's
            mbar_arrive_expect_tx(&bar_load[xbuf ^ 1], kBlobX);
            bulk_g2s_hint(smem + L::BX + (xbuf ^ 1) * kBlobX, a.blobs + (size_t)(tile + gridDim.x) * kBlobX, kBlobX, &bar_load[xbuf ^ 1], pol_stream);
        }
        // ---- dZ5 = loss_scale * dL/drgb * act'(rgb)   (16 columns, 3 live)
        if (hsel == 0) {
            float g[3] = {0.f, 0.f, 0.f};
            if (valid) {
                const uint2 yh = pfy;
                const float2 y01 = __half22float2(*reinterpret_cast<const __half2*>(&yh.x));
                const float y3[3] = {y01.x, y01.y, __low2float(*reinterpret_cast<const __half2*>(&yh.y))};
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float yv = y3[k];
                    float d = pf[k] * loss_scale;
                    if (a.rgb_act == MFN_ACT_SIGMOID) d *= yv * (1.f - yv);
                    else if (a.rgb_act == MFN_ACT_EXP) d *= yv;
                    g[k] = d;
                }
            }
            uint4 o0 = make_uint4(pack2(g[0], g[1]), pack2(g[2], 0.f), 0u, 0u);
            const __half2* hh = reinterpret_cast<const __half2*>(&o0);
            bad |= !isfinite(__low2float(hh[0])) || !isfinite(__high2float(hh[0])) || !isfinite(__low2float(hh[1]));
            *reinterpret_cast<uint4*>(smem + L::BDZo + tile_off(row, 0, 16)) = o0;
            *reinterpret_cast<uint4*>(smem + L::BDZo + tile_off(row, 8, 16)) = make_uint4(0u, 0u, 0u, 0u);
        } else if (hsel == 1) {   // the partner thread re-encodes the view direction: CAT[:, 0:16] (same code, same bits as the forward pass)
            uint4 o0 = make_uint4(0u, 0u, 0u, 0u), o1 = o0;
            if (valid) sh_of_dir(pf[0], pf[1], pf[2], o0, o1);
            *reinterpret_cast<uint4*>(smem + L::BC + tile_off(row, 0, 32)) = o0;
            *reinterpret_cast<uint4*>(smem + L::BC + tile_off(row, 8, 32)) = o1;
        }
        if (xbuf == 0) { mbar_wait(&bar_load[0], ph_load0); ph_load0 ^= 1u; } else { mbar_wait(&bar_load[1], ph_load1); ph_load1 ^= 1u; }
        MFN_STAGE_SYNC();
        const float dsig = (valid && hsel == 0) ? a.dL_dsigmas[i] : 0.f;        // used in stage C; requested after the boundary (see below)
        // ---- recomputed forward: H1 = relu(X.W1^T) ; h = H1.W2^T -> CAT[:, 16:32] ; H2 = relu(CAT.W3^T) ; H3 = relu(H2.W4^T)
        if (issuer) {
            tc_fence_after();
            mma_chain(tbase + L::AccH, desc_kmajor(sX, 32, 0), kstep_kmajor(), desc_kmajor(sbase + L::W1, 32, 0), kstep_kmajor(), idesc_f16(128, 64, false, false), 2, 0u);
            mma_commit(&bar_mma);
        }
        MFN_MMA_WAIT();
        if (worker) relu_epilogue<64>(trow + L::AccH, smem + L::BH1, row, hsel);
        MFN_STAGE_SYNC();
        if (issuer) {
            tc_fence_after();
            mma_chain(tbase + L::AccO2, desc_kmajor(sH1, 64, 0), kstep_kmajor(), desc_kmajor(sbase + L::W2, 64, 0), kstep_kmajor(), idesc_f16(128, 16, false, false), 4, 0u);
            mma_commit(&bar_mma);
        }
        MFN_MMA_WAIT();
        if (hsel == 0) {
            uint32_t r[16];
            tmem_ld_x16(trow + L::AccO2, r);
            tmem_ld_wait();
            uint4 o0, o1;
            pack16(r, o0, o1);
            *reinterpret_cast<uint4*>(smem + L::BC + tile_off(row, 16, 32)) = o0;
            *reinterpret_cast<uint4*>(smem + L::BC + tile_off(row, 24, 32)) = o1;
        }
        MFN_STAGE_SYNC();
        if (issuer) {
            tc_fence_after();
            mma_chain(tbase + L::AccH, desc_kmajor(sC, 32, 0), kstep_kmajor(), desc_kmajor(sbase + L::W3, 32, 0), kstep_kmajor(), idesc_f16(128, RW, false, false), 2, 0u);
            mma_commit(&bar_mma);
        }
        MFN_MMA_WAIT();
        if (worker) relu_epilogue<RW>(trow + L::AccH, smem + L::BH2, row, hsel);
        MFN_STAGE_SYNC();
        if (NH2 == 2) {
            if (issuer) {
                tc_fence_after();
                mma_chain(tbase + L::AccH, desc_kmajor(sH2, RW, 0), kstep_kmajor(), desc_kmajor(sbase + L::W4, RW, 0), kstep_kmajor(), idesc_f16(128, RW, false, false), RW / 16, 0u);
                mma_commit(&bar_mma);
            }
            MFN_MMA_WAIT();
            if (worker) relu_epilogue<RW>(trow + L::AccH, smem + L::BH3, row, hsel);
            MFN_STAGE_SYNC();
        }
        // ---- stage A: dH_last = dZ5 . W5 ;  dW5^T += H_last^T . dZ5
        if (issuer) {
            tc_fence_after();
            mma_f16_ss(tbase + L::AccH, desc_kmajor(sDZo, 16, 0), desc_mnmajor(sbase + L::W5, RW, 0), idesc_f16(128, RW, false, true), 0u);
            mma_commit(&bar_mma);
            mma_chain(tbase + L::AccW5, desc_mnmajor(sHL, RW, 0), kstep_mnmajor(RW), desc_mnmajor(sDZo, 16, 0), kstep_mnmajor(16), idesc_f16(RW, 16, true, true), kFT / 16, acc);
            mma_commit(&bar_dw);
        }
        MFN_MMA_WAIT();
        if (worker) mask_epilogue<RW>(trow + L::AccH, pHL, row, hsel, &bar_dw, ph_dw);   // dZ of the last hidden layer, in place
        MFN_STAGE_SYNC();
        // next tile's per-row inputs: requested right AFTER a stage boundary -- the proxy fence of a boundary waits for the thread's
        // outstanding global loads (measured: +1.5 k cycles when they were issued just before one), the MMA wait that follows hides them
        if (worker && tile + gridDim.x < n_tiles) prefetch_row(tile + gridDim.x);
        if (NH2 == 2) {
            // ---- stage B: dH2 = dZ4 . W4 ;  dW4^T += H2^T . dZ4
            if (issuer) {
                tc_fence_after();
                mma_chain(tbase + L::AccH, desc_kmajor(sH3, RW, 0), kstep_kmajor(), desc_mnmajor(sbase + L::W4, RW, 0), kstep_mnmajor(RW), idesc_f16(128, RW, false, true), RW / 16, 0u);
                mma_commit(&bar_mma);
                mma_chain(tbase + L::AccW4, desc_mnmajor(sH2, RW, 0), kstep_mnmajor(RW), desc_mnmajor(sH3, RW, 0), kstep_mnmajor(RW), idesc_f16(RW, RW, true, true), kFT / 16, acc);
                mma_commit(&bar_dw);
            }
            MFN_MMA_WAIT();
            if (worker) mask_epilogue<RW>(trow + L::AccH, smem + L::BH2, row, hsel, &bar_dw, ph_dw);    // dZ3 in place of H2
            MFN_STAGE_SYNC();
        }
        // ---- stage C: dCAT = dZ3 . W3 (32 columns) ;  dW3 += dZ3^T . CAT
        if (issuer) {
            tc_fence_after();
            mma_chain(tbase + L::AccH, desc_kmajor(sH2, RW, 0), kstep_kmajor(), desc_mnmajor(sbase + L::W3, 32, 0), kstep_mnmajor(32), idesc_f16(128, 32, false, true), RW / 16, 0u);
            mma_commit(&bar_mma);
            // (dW3 reads dZ3 and CAT, which nothing writes before the next tile: no barrier of its own -- every later commit covers it)
            mma_chain(tbase + L::AccW3, desc_mnmajor(sH2, RW, 0), kstep_mnmajor(RW), desc_mnmajor(sC, 32, 0), kstep_mnmajor(32), idesc_f16(RW, 32, true, true), kFT / 16, acc);
        }
        MFN_MMA_WAIT();
        if (hsel == 0) {   // dh = dCAT[:, 16:32] ; dh[0] += loss_scale * dL/dsigma * exp(clamp(h0, -15, 15))   (TruncExp backward)
            uint32_t r[16];
            tmem_ld_x16(trow + L::AccH + 16, r);
            tmem_ld_wait();
            float d0 = __uint_as_float(r[0]);
            if (valid) {
                const float h0 = __half2float(*reinterpret_cast<const __half*>(smem + L::BC + tile_off(row, 16, 32)));
                d0 += dsig * expf(fminf(fmaxf(h0, -15.f), 15.f)) * loss_scale;
            }
            r[0] = __float_as_uint(d0);
            uint4 o0, o1;
            pack16(r, o0, o1);
            bad |= !isfinite(__low2float(*reinterpret_cast<const __half2*>(&o0.x)));
            *reinterpret_cast<uint4*>(smem + L::BDZo + tile_off(row, 0, 16)) = o0;
            *reinterpret_cast<uint4*>(smem + L::BDZo + tile_off(row, 8, 16)) = o1;
        }
        MFN_STAGE_SYNC();
        // ---- stage D: dH1 = dZ2 . W2 ;  dW2^T += H1^T . dZ2
        if (issuer) {
            tc_fence_after();
            mma_f16_ss(tbase + L::AccH, desc_kmajor(sDZo, 16, 0), desc_mnmajor(sbase + L::W2, 64, 0), idesc_f16(128, 64, false, true), 0u);
            mma_commit(&bar_mma);
            mma_chain(tbase + L::AccW2, desc_mnmajor(sH1, 64, 0), kstep_mnmajor(64), desc_mnmajor(sDZo, 16, 0), kstep_mnmajor(16), idesc_f16(64, 16, true, true), kFT / 16, acc);
            mma_commit(&bar_dw);
        }
        MFN_MMA_WAIT();
        if (worker) mask_epilogue<64>(trow + L::AccH, smem + L::BH1, row, hsel, &bar_dw, ph_dw);        // dZ1 in place of H1
        MFN_STAGE_SYNC();
        // ---- stage E: dX = dZ1 . W1 (32 columns) ;  dW1 += dZ1^T . X
        if (issuer) {
            tc_fence_after();
            mma_chain(tbase + L::AccH, desc_kmajor(sH1, 64, 0), kstep_kmajor(), desc_mnmajor(sbase + L::W1, 32, 0), kstep_mnmajor(32), idesc_f16(128, 32, false, true), 4, 0u);
            mma_commit(&bar_mma);
            // dW1 reads dZ1 (H1 buffer) and this tile's X buffer.  H1 is written again by the next tile's first epilogue, behind a commit
            // that covers dW1; the X buffer by the bulk copy the issuer starts at the top of the next tile -- after waiting on bar_tail
            mma_chain(tbase + L::AccW1, desc_mnmajor(sH1, 64, 0), kstep_mnmajor(64), desc_mnmajor(sX, 32, 0), kstep_mnmajor(32), idesc_f16(64, 32, true, true), kFT / 16, acc);
            mma_commit(&bar_tail);
        }
        MFN_MMA_WAIT();
        if (worker) {   // dX row -> dfeats, level-major [16][stride] half2: coalesced here and in the scatter kernel; each thread of a row does 8 levels
            uint32_t r[16];
            tmem_ld_x16(trow + L::AccH + 16 * hsel, r);
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
        MFN_STAGE_SYNC();   // generic reads/writes of the tile area are ordered before the next tile's writes and bulk copy
    }
    if (bad && a.overflow) *a.overflow = 1;
    if (issuer && tile_no > 0) mbar_wait(&bar_tail, ph_tail);      // the last tile's trailing weight-gradient MMAs have landed in TMEM
    tc_fence_before();
    __syncthreads();
    // ---- flush this CTA's weight-gradient accumulators.  M = 64 accumulators: row m lives in TMEM lane 32*(m/16) + m%16;
    //      M = 128 accumulators: row m lives in lane m.
    float* part = a.partials + (size_t)blockIdx.x * L::NumWg;
    if (acc == 0u) {
        for (int q = tid; q < L::NumWg; q += kBwdThreads) part[q] = 0.f;
    } else if (warp < 4) {
        tc_fence_after();
        const int m64 = warp * 16 + (lane & 15);
        const bool own64 = lane < 16;
        const int mrw = (RW == 64) ? m64 : warp * 32 + lane;
        const bool ownrw = (RW == 64) ? own64 : true;
        constexpr int oW1 = 0, oW2 = 64 * 32, oW3 = oW2 + 16 * 64, oW4 = oW3 + RW * 32, oW5 = oW4 + (NH2 - 1) * RW * RW;
        {
            uint32_t r[32];
            tmem_ld_x32(trow + L::AccW1, r); tmem_ld_wait();          // dW1[out = m][in = j]
            if (own64) for (int j = 0; j < 32; ++j) part[oW1 + m64 * 32 + j] = __uint_as_float(r[j]);
            tmem_ld_x32(trow + L::AccW3, r); tmem_ld_wait();          // dW3[out = m][in = j]
            if (ownrw) for (int j = 0; j < 32; ++j) part[oW3 + mrw * 32 + j] = __uint_as_float(r[j]);
            if (NH2 == 2) {
#pragma unroll
                for (int h = 0; h < RW / 32; ++h) {                    // dW4^T[in = m][out = j]
                    tmem_ld_x32(trow + L::AccW4 + 32 * h, r); tmem_ld_wait();
                    if (ownrw) for (int j = 0; j < 32; ++j) part[oW4 + (32 * h + j) * RW + mrw] = __uint_as_float(r[j]);
                }
            }
        }
        {
            uint32_t r[16];
            tmem_ld_x16(trow + L::AccW2, r); tmem_ld_wait();          // dW2^T[in = m][out = j]
            if (own64) for (int j = 0; j < 16; ++j) part[oW2 + j * 64 + m64] = __uint_as_float(r[j]);
            tmem_ld_x16(trow + L::AccW5, r); tmem_ld_wait();          // dW5^T[in = m][out = j]
            if (ownrw) for (int j = 0; j < 16; ++j) part[oW5 + j * RW + mrw] = __uint_as_float(r[j]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, L::BwdCols);
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
    return (c->grid.grid_type == MFN_GRID_HASH || c->grid.grid_type == MFN_GRID_MIXED) && c->grid.n_levels == 16 && c->grid.n_features == 2 &&
           c->sigma_width == 64 && c->sigma_hidden == 1 && (c->rgb_width == 64 || c->rgb_width == 128) && (c->rgb_hidden == 1 || c->rgb_hidden == 2);
}
static int bwd_max_ctas(int rgb_width) { return (rgb_width == 64 ? Lay<64>::BwdCtas : Lay<128>::BwdCtas) * num_sms(); }
size_t fused_blob_bytes(int64_t n_max) { return (size_t)ceil_div(n_max, kFT) * kBlobX; }
size_t fused_partial_bytes(int rgb_width) {
    return (size_t)bwd_max_ctas(rgb_width) * (rgb_width == 64 ? Lay<64>::NumWg : Lay<128>::NumWg) * sizeof(float);
}

template <int NH2, int MODE, bool MIXED, int RW>
static void launch_fwd_m(const FusedArgs& a, const GridMeta& m, cudaStream_t st) {
    constexpr int smem_bytes = Lay<RW>::FwdSmem;
    constexpr int max_ctas = Lay<RW>::FwdCtas;
    static bool once = (cudaFuncSetAttribute(field_fwd_fused_kernel<NH2, MODE, MIXED, RW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes), true);
    (void)once;
    const int64_t tiles = ceil_div(a.n_max, kFT);
    static int ctas_per_sm = 0;
    if (ctas_per_sm == 0) { const char* e = getenv("MFN_FWD_CTAS"); ctas_per_sm = e ? atoi(e) : max_ctas; if (ctas_per_sm < 1 || ctas_per_sm > max_ctas) ctas_per_sm = max_ctas; }
    const int64_t cap = ctas_per_sm * (int64_t)num_sms();
    // one scheduler slot per launch, 16 per mode in rotation: launches of one mode that are in flight at the same time (different
    // streams) never share a slot unless 16 of them overlap; a captured launch keeps its slot, its replays are stream-ordered
    static unsigned int* sched_base = nullptr;
    static std::atomic<unsigned> seq{0};
    static const bool dynamic = !(getenv("MFN_FWD_STATIC") && atoi(getenv("MFN_FWD_STATIC")));
    if (!sched_base) { void* p = nullptr; if (cudaGetSymbolAddress(&p, g_tile_sched) == cudaSuccess) sched_base = (unsigned int*)p; }
    FusedArgs b = a;
    b.sched = (dynamic && sched_base) ? sched_base + 2 * ((MODE & 3) * 16 + (seq.fetch_add(1) & 15u)) : nullptr;
    field_fwd_fused_kernel<NH2, MODE, MIXED, RW><<<(unsigned)(tiles < cap ? tiles : cap), kFwdThreads, smem_bytes, st>>>(b, m);
}
template <int NH2, int MODE, int RW>
static void launch_fwd(const FusedArgs& a, const GridMeta& m, cudaStream_t st) {
    if (m.mixed) launch_fwd_m<NH2, MODE, true, RW>(a, m, st); else launch_fwd_m<NH2, MODE, false, RW>(a, m, st);
}
template <int NH2, int RW>
static void launch_fwd_mode(const FusedArgs& a, const GridMeta& m, int mode, cudaStream_t st) {
    if (mode == 0) launch_fwd<NH2, 0, RW>(a, m, st); else if (mode == 4) launch_fwd<NH2, 4, RW>(a, m, st); else launch_fwd<NH2, 1, RW>(a, m, st);
}

// mode: 0 inference, 1 training, 2 density only, 3 raw 16 outputs of the sigma network, 4 inference inside the render wavefront
int fused_field_forward(const FusedArgs& a, const GridMeta& m, int rgb_width, int rgb_hidden, int mode, cudaStream_t st) {
    ProfScope ps(mode >= 2 ? "density_fwd" : "field_fwd", st);
    if (mode == 3) { launch_fwd<1, 3, 64>(a, m, st); return check_launch("mfn_geo_fwd(fused)", st); }      // (the sigma network only: the rgb width does not matter)
    if (mode == 2) { launch_fwd<1, 2, 64>(a, m, st); return check_launch("mfn_density_fwd(fused)", st); }
    if (rgb_width == 128) { if (rgb_hidden == 2) launch_fwd_mode<2, 128>(a, m, mode, st); else launch_fwd_mode<1, 128>(a, m, mode, st); }
    else { if (rgb_hidden == 2) launch_fwd_mode<2, 64>(a, m, mode, st); else launch_fwd_mode<1, 64>(a, m, mode, st); }
    return check_launch("mfn_field_fwd(fused)", st);
}

template <int NH2, int RW>
static int launch_bwd(const FusedArgs& a, cudaStream_t st) {
    static bool once = (cudaFuncSetAttribute(field_bwd_fused_kernel<NH2, RW>, cudaFuncAttributeMaxDynamicSharedMemorySize, Lay<RW>::BwdSmem), true);
    (void)once;
    const int64_t tiles = ceil_div(a.n_max, kFT);
    int64_t cap = bwd_max_ctas(RW);
    { static const char* e = getenv("MFN_BWD_CTAS"); if (e && atoi(e) > 0 && atoi(e) < cap) cap = atoi(e); }      // (experiments: chains per SM)
    const unsigned grid = (unsigned)(tiles < cap ? tiles : cap);
    field_bwd_fused_kernel<NH2, RW><<<grid, kBwdThreads, Lay<RW>::BwdSmem, st>>>(a);
    return (int)grid;
}

int fused_field_backward(const FusedArgs& a, int rgb_width, int rgb_hidden, float* d_sigma_params, float* d_rgb_params, WgradReduce* wr, cudaStream_t st) {
    int grid;
    {
        ProfScope ps("field_bwd", st);
        if (rgb_width == 128) grid = rgb_hidden == 2 ? launch_bwd<2, 128>(a, st) : launch_bwd<1, 128>(a, st);
        else grid = rgb_hidden == 2 ? launch_bwd<2, 64>(a, st) : launch_bwd<1, 64>(a, st);
    }
    wr->partials = a.partials; wr->n_parts = grid; wr->stride = rgb_width == 128 ? Lay<128>::NumWg : Lay<64>::NumWg;
    wr->n_sigma = 3072; wr->n_rgb = rgb_width * 32 + (rgb_hidden - 1) * rgb_width * rgb_width + 16 * rgb_width;
    wr->d_sigma = d_sigma_params; wr->d_rgb = d_rgb_params;
    return check_launch("mfn_field_bwd(fused)", st);
}

}  // namespace mfn
