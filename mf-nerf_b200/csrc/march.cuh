// Occupancy-bitfield ray marching, shared device code  (ref: models/csrc/raymarching.cu:7-32, 166-279, 335-404)
//
// Bit-exactness contract.  N_samples, ts, deltas and xyzs must equal the reference's to the last bit, so
// every fp32 operation below is spelled with a round-to-nearest intrinsic that nvcc can neither contract
// nor re-associate, in exactly the shape the reference's own SASS has when built with its stock flags
// (-O2, -fmad=true, -prec-div=true; checked with cuobjdump on oracle/_ref/vren_ref*.so):
//     x        = fma(d, t, o)                               (raymarching.cu:205)
//     dt       = max(dt_min, min(t*esf, dt_max))            (helper_math clamp; raymarching.cu:11-13)
//     cell     = trunc(max(0, min(fma(x, 1/bound, 1)*0.5*G, G-1)))            (l.215-217)
//     t_exit_x = fma(bound, fma((n+0.5 + 0.5*sign)*(1/G), 2, -1), -x) * (1/d) (l.225)
//     t_next   = dt + t                                     (l.222, l.231)
// Key structural fact used by the warp marcher: both the "occupied" step and the "skip" loop advance t
// with the same map t -> t + dt(t), so the sequence of candidate positions along a ray (the "lattice")
// does not depend on occupancy.  32 lanes evaluate 32 consecutive lattice points at once and the
// sequential visit order of the reference is then replayed with ballots.
#pragma once
#include "common.cuh"

namespace mfn {

struct MarchConst {
    float dt_min, dt_max, esf;
    float gs, gs_inv, gsm1;   // (float)G, 1/G, G-1
    float scale;              // mip_bound clamp
    float bound0, bound0_inv; // mip 0: min(2^-1, scale) and its reciprocal
    int cascades, grid_size;
    uint32_t g3;
    const uint32_t* lut;      // optional shared-memory table: lut[v] = v's bits spread to every third position (morton_lut_fill)
};

// 10-bit coordinate -> its bits at positions 0, 3, 6, ...: three shared-memory loads replace ~24 integer instructions per probe
__device__ __forceinline__ void morton_lut_fill(uint32_t* lut, int grid_size, int tid, int nthreads) {
    for (int v = tid; v < grid_size; v += nthreads) lut[v] = spread3_mul((uint32_t)v);
}

// `dt_scale` is `scale` for the train marcher and (float)cascades for the test marcher -- the reference
// passes `cascades` to calc_dt in raymarching_test_kernel (raymarching.cu:370,399), a quirk we keep.
__device__ __forceinline__ MarchConst make_march_const(int cascades, int grid_size, float scale, float esf, int max_samples, float dt_scale) {
    MarchConst c;
    c.esf = esf;
    c.dt_min = __fdiv_rn(1.73205080757f, (float)max_samples);
    c.dt_max = __fdiv_rn(__fmul_rn(dt_scale, 3.4641015529632568359f), (float)grid_size);
    c.gs = (float)grid_size;
    c.gs_inv = __fdiv_rn(1.0f, c.gs);
    c.gsm1 = __fadd_rn(c.gs, -1.0f);
    c.scale = scale;
    c.bound0 = fminf(0.5f, scale);
    c.bound0_inv = __fdiv_rn(1.0f, c.bound0);
    c.cascades = cascades;
    c.grid_size = grid_size;
    c.g3 = (uint32_t)grid_size * grid_size * grid_size;
    c.lut = nullptr;
    return c;
}

__device__ __forceinline__ float march_dt(float t, const MarchConst& c) {
    return fmaxf(fminf(__fmul_rn(t, c.esf), c.dt_max), c.dt_min);
}
__device__ __forceinline__ float march_next(float t, const MarchConst& c) { return __fadd_rn(march_dt(t, c), t); }

struct RayConst {
    float ox, oy, oz, dx, dy, dz, ix, iy, iz, sx, sy, sz;
};
__device__ __forceinline__ RayConst make_ray(const float* __restrict__ o, const float* __restrict__ d, int64_t r) {
    RayConst q;
    q.ox = o[3 * r]; q.oy = o[3 * r + 1]; q.oz = o[3 * r + 2];
    q.dx = d[3 * r]; q.dy = d[3 * r + 1]; q.dz = d[3 * r + 2];
    q.ix = __fdiv_rn(1.0f, q.dx); q.iy = __fdiv_rn(1.0f, q.dy); q.iz = __fdiv_rn(1.0f, q.dz);
    q.sx = copysignf(1.0f, q.dx); q.sy = copysignf(1.0f, q.dy); q.sz = copysignf(1.0f, q.dz);
    return q;
}

struct Probe {
    float x, y, z, dt, t_target;
    bool occ;
};

// Evaluate one lattice point: position, step, cascade, cell, occupancy bit and (if empty) the t at which
// the ray leaves the cell.  (ref: raymarching.cu:205-228)
// ONE_CASCADE: cascades == 1 makes mip = clamp(.., 0, 0) = 0 whatever the position, so the two frexpf and the division are
// hoisted into MarchConst (bound0 = min(2^-1, scale), bound0_inv = 1/bound0 -- the same IEEE operations, evaluated once).
template <bool ONE_CASCADE = false, bool LUT = false>
__device__ __forceinline__ Probe probe_cell(float t, const RayConst& q, const MarchConst& c, const uint8_t* __restrict__ bitfield) {
    Probe p;
    p.x = __fmaf_rn(q.dx, t, q.ox); p.y = __fmaf_rn(q.dy, t, q.oy); p.z = __fmaf_rn(q.dz, t, q.oz);
    p.dt = march_dt(t, c);
    int mip = 0;
    float bound = c.bound0, bound_inv = c.bound0_inv;
    if (!ONE_CASCADE) {
        int e_pos, e_dt;
        (void)frexpf(fmaxf(fabsf(p.x), fmaxf(fabsf(p.y), fabsf(p.z))), &e_pos);
        (void)frexpf(__fmul_rn(p.dt, c.gs), &e_dt);
        const int mip_pos = min(c.cascades - 1, max(0, e_pos + 1));
        const int mip_dt = min(c.cascades - 1, max(0, e_dt));
        mip = max(mip_pos, mip_dt);
        bound = fminf(scalbnf(1.0f, mip - 1), c.scale);
        bound_inv = __fdiv_rn(1.0f, bound);
    }
    const float fx = __fmul_rn(__fmul_rn(__fmaf_rn(p.x, bound_inv, 1.0f), 0.5f), c.gs);
    const float fy = __fmul_rn(__fmul_rn(__fmaf_rn(p.y, bound_inv, 1.0f), 0.5f), c.gs);
    const float fz = __fmul_rn(__fmul_rn(__fmaf_rn(p.z, bound_inv, 1.0f), 0.5f), c.gs);
    const int nx = (int)fmaxf(0.0f, fminf(fx, c.gsm1));
    const int ny = (int)fmaxf(0.0f, fminf(fy, c.gsm1));
    const int nz = (int)fmaxf(0.0f, fminf(fz, c.gsm1));
    const uint32_t code = LUT ? (c.lut[nx] | (c.lut[ny] << 1) | (c.lut[nz] << 2)) : morton_encode((uint32_t)nx, (uint32_t)ny, (uint32_t)nz);
    const uint32_t idx = (uint32_t)mip * c.g3 + code;
    p.occ = (__ldg(bitfield + (idx >> 3)) >> (idx & 7u)) & 1u;
    const float ax = __fmaf_rn(q.sx, 0.5f, __fadd_rn((float)nx, 0.5f));
    const float ay = __fmaf_rn(q.sy, 0.5f, __fadd_rn((float)ny, 0.5f));
    const float az = __fmaf_rn(q.sz, 0.5f, __fadd_rn((float)nz, 0.5f));
    const float tx = __fmul_rn(__fmaf_rn(bound, __fmaf_rn(__fmul_rn(ax, c.gs_inv), 2.0f, -1.0f), -p.x), q.ix);
    const float ty = __fmul_rn(__fmaf_rn(bound, __fmaf_rn(__fmul_rn(ay, c.gs_inv), 2.0f, -1.0f), -p.y), q.iy);
    const float tz = __fmul_rn(__fmaf_rn(bound, __fmaf_rn(__fmul_rn(az, c.gs_inv), 2.0f, -1.0f), -p.z), q.iz);
    p.t_target = __fadd_rn(t, fmaxf(0.0f, fminf(tx, fminf(ty, tz))));
    return p;
}

// ---------------------------------------------------------------------------------------------------
// Warp marcher.  One warp walks one ray, 32 lattice points per batch; `emit(rank, t, dt)` is called by the lane
// that owns an accepted sample, rank = its index along the ray.  Stops after `max_emit` samples or when t leaves
// [0, t2).  Returns (warp-uniform) the number of samples emitted.
//
// Instruction diet (the kernel is issue-bound: every lane-redundant instruction is paid 32 times):
//   * lattice in O(1) when dt is constant (CONST_DT: exp_step_factor == 0).  Inside one binade the recurrence
//     t' = fl(t + dt) advances by a constant q = fl(t + dt) - t, so lane j proposes c_j = fma(j, q, t_base) (exact: a
//     multiple of the binade's ulp) and the warp VERIFIES the recurrence, fl(c_j + dt) == c_{j+1} for all j -- by
//     induction from c_0 = t_base this proves c_j is the sequential value, bit for bit.  A failed check (binade
//     crossing, a tie) falls back to the 31-step chain for that batch.
//   * the landing point of an empty-space skip, "first lattice point with t >= t_target", from the same closed form
//     (estimate, then an exact +-1 correction against c_j); binary search by shuffles in the fallback.
//   * the reference's sequential visit order is replayed with warp-uniform bit operations on the ballots; only a skip
//     needs a shuffle (to fetch the landing index of the lane it starts from).
// ---------------------------------------------------------------------------------------------------
template <bool ONE_CASCADE, bool CONST_DT, typename Emit>
__device__ __forceinline__ int march_ray_warp(float t_start, float t2, int max_emit, const RayConst& q, const MarchConst& c,
                                              const uint8_t* __restrict__ bitfield, int lane, Emit emit) {
    constexpr uint32_t FULL = 0xffffffffu;
    int n = 0;
    float t_base = t_start;
    float pending = -INFINITY;  // skip target carried over from the previous batch
    // (ref loop condition: 0<=t && t<t2 && N_samples<max_samples, raymarching.cu:204)
    while (t_base >= 0.0f && t_base < t2 && n < max_emit) {
        float my_t, t_next_base, qstep = 0.f;
        bool closed = false;
        if (CONST_DT) {
            qstep = __fsub_rn(__fadd_rn(c.dt_min, t_base), t_base);
            my_t = __fmaf_rn((float)lane, qstep, t_base);
            const float succ = __fadd_rn(c.dt_min, my_t);                       // what the recurrence gives after my_t
            const float nxt_c = __shfl_down_sync(FULL, my_t, 1);
            closed = __all_sync(FULL, lane == 31 || succ == nxt_c);
            t_next_base = __shfl_sync(FULL, succ, 31);
        }
        if (!closed) {   // sequential chain (always for exp_step_factor > 0)
            float tt = t_base; my_t = t_base;
#pragma unroll 4
            for (int j = 0; j < 32; ++j) {
                if (j == lane) my_t = tt;
                tt = march_next(tt, c);
            }
            t_next_base = tt;
        }
        const bool active = my_t < t2;
        Probe p;
        p.occ = false; p.t_target = 0.f; p.dt = 0.f;
        if (active) p = probe_cell<ONE_CASCADE>(my_t, q, c, bitfield);
        const uint32_t act_mask = __ballot_sync(FULL, active);
        const uint32_t occ_mask = __ballot_sync(FULL, active && p.occ);
        // landing index of a skip that starts from this lane: smallest j > lane with t_j >= t_target, 32 = beyond this batch
        int land = 32;
        if (closed) {
            if (active && !p.occ) {
                int j = (int)(__fsub_rn(p.t_target, t_base) / qstep);           // estimate; the exact test follows
                j = max(lane + 1, min(j, 33));
                // move down while the previous point already satisfies t >= target, then up while this one does not
                while (j - 1 > lane && __fmaf_rn((float)(j - 1), qstep, t_base) >= p.t_target) --j;
                while (j < 32 && __fmaf_rn((float)j, qstep, t_base) < p.t_target) ++j;
                land = min(j, 32);
            }
        } else {
            int lo = lane + 1, hi = 32;
#pragma unroll
            for (int it = 0; it < 5; ++it) {
                const int mid = (lo + hi) >> 1;
                const float tm = __shfl_sync(FULL, my_t, mid < 32 ? mid : 31);
                if (lo < hi) { if (tm >= p.t_target) hi = mid; else lo = mid + 1; }
            }
            land = lo;
        }
        // first lattice point of this batch that the sequential marcher visits
        int cur = 0;
        if (pending > -INFINITY) {
            const uint32_t m = __ballot_sync(FULL, my_t >= pending);
            cur = m ? (__ffs(m) - 1) : 32;
            if (cur < 32) pending = -INFINITY;
        }
        uint32_t take = 0;
        bool finished = false;
        while (cur < 32) {
            if (!((act_mask >> cur) & 1u)) { finished = true; break; }  // t >= t2: the ray left the box
            if ((occ_mask >> cur) & 1u) {                                 // run of consecutive occupied points: all taken
                const uint32_t rest = ~(occ_mask >> cur);
                const int run = rest ? (__ffs(rest) - 1) : 32;
                take |= (run >= 32 ? FULL : ((1u << run) - 1u)) << cur;
                cur += run;
            } else {                                                      // empty cell: skip to its exit
                const int nx = __shfl_sync(FULL, land, cur);
                if (nx >= 32) pending = __shfl_sync(FULL, p.t_target, cur);
                cur = nx;
            }
        }
        int cnt = __popc(take);
        if (n + cnt >= max_emit) {  // sample budget reached inside this batch: keep the first ones only
            const int keep = max_emit - n;
            if (cnt > keep) {
                uint32_t m = take;  // drop the lowest keep-1 set bits; what is left starts at the keep-th one
                for (int i = 1; i < keep; ++i) m &= m - 1;
                const int last = __ffs(m) - 1;
                take &= (last >= 31) ? FULL : ((1u << (last + 1)) - 1u);
                cnt = keep;
            }
            finished = true;
        }
        if ((take >> lane) & 1u) emit(n + __popc(take & ((1u << lane) - 1u)), my_t, p.dt);
        n += cnt;
        if (finished) break;
        t_base = t_next_base;
    }
    return n;
}

// Fast path of the warp marcher: same lattice, same probes, but NO replay of the visit order.  In exact arithmetic every lattice
// point of an occupied cell is visited (an empty cell's skip lands on the first point behind its exit); in fp32 the one thing
// that can go wrong is a skip target that overshoots the first point of the next occupied run.  So the warp only PROVES that this
// did not happen: for every occupied point j whose predecessor is not occupied (a run start), every empty point i since the last
// visited occupied point (or the ray start) has t_target(i) <= t_j.  Then the chain, which starts at a visited point (the ray's
// first point, or the successor of a visited occupied point), can only land at or before j, lands on empty points until it
// reaches j, and advances by at least one point per hop: j is visited, and by induction every occupied point is taken.
// `carry` is the running maximum of t_target over the empty points of the current gap (as ordered bits: t_target >= t >= 0).
// Returns the number of samples, or -1 when one run start could not be proven: the caller then re-marches the ray with
// march_ray_warp (the exact replay).  Expected rate: an exit within an ulp or two of a lattice point, a few rays in 10^3.
template <bool ONE_CASCADE, bool CONST_DT, bool LUT, typename Emit>
__device__ __forceinline__ int march_ray_warp_fast(float t_start, float t2, int max_emit, const RayConst& q, const MarchConst& c,
                                                   const uint8_t* __restrict__ bitfield, int lane, Emit emit) {
    constexpr uint32_t FULL = 0xffffffffu;
    int n = 0;
    float t_base = t_start;
    uint32_t carry = 0u;
    while (t_base >= 0.0f && t_base < t2 && n < max_emit) {
        float my_t, t_next_base;
        bool closed = false;
        if (CONST_DT) {
            const float qstep = __fsub_rn(__fadd_rn(c.dt_min, t_base), t_base);
            my_t = __fmaf_rn((float)lane, qstep, t_base);
            const float succ = __fadd_rn(c.dt_min, my_t);
            const float nxt_c = __shfl_down_sync(FULL, my_t, 1);
            closed = __all_sync(FULL, lane == 31 || succ == nxt_c);
            t_next_base = __shfl_sync(FULL, succ, 31);
        }
        if (!closed) {
            float tt = t_base; my_t = t_base;
#pragma unroll 4
            for (int j = 0; j < 32; ++j) {
                if (j == lane) my_t = tt;
                tt = march_next(tt, c);
            }
            t_next_base = tt;
        }
        const bool active = my_t < t2;
        Probe p;
        p.occ = false; p.t_target = 0.f; p.dt = 0.f;
        if (active) p = probe_cell<ONE_CASCADE, LUT>(my_t, q, c, bitfield);
        const uint32_t act_mask = __ballot_sync(FULL, active);
        const uint32_t occ_mask = __ballot_sync(FULL, active && p.occ);
        const uint32_t e_bits = (active && !p.occ) ? __float_as_uint(p.t_target) : 0u;
        uint32_t starts = occ_mask & ~(occ_mask << 1);
        while (starts) {                                      // (warp-uniform) usually 0 or 1 run start per batch
            const int j = __ffs(starts) - 1;
            starts &= starts - 1u;
            const uint32_t below = occ_mask & ((1u << j) - 1u);
            const int prev = below ? 31 - __clz(below) : -1;  // last occupied lane in front of j
            const bool in_gap = lane > prev && lane < j;
            uint32_t m = __reduce_max_sync(FULL, in_gap ? e_bits : 0u);
            if (prev < 0) m = max(m, carry);
            if (m > __shfl_sync(FULL, __float_as_uint(my_t), j)) return -1;
        }
        const int last_occ = occ_mask ? 31 - __clz(occ_mask) : -1;
        const uint32_t trail = __reduce_max_sync(FULL, lane > last_occ ? e_bits : 0u);
        carry = last_occ < 0 ? max(carry, trail) : trail;
        uint32_t take = occ_mask;
        bool finished = act_mask != FULL;                    // a point with t >= t2 exists: the chain ends at or before the next batch
        int cnt = __popc(take);
        if (n + cnt >= max_emit) {
            const int keep = max_emit - n;
            if (cnt > keep) {
                uint32_t m = take;
                for (int i = 1; i < keep; ++i) m &= m - 1;
                const int last = __ffs(m) - 1;
                take &= (last >= 31) ? FULL : ((1u << (last + 1)) - 1u);
                cnt = keep;
            }
            finished = true;
        }
        if ((take >> lane) & 1u) emit(n + __popc(take & ((1u << lane) - 1u)), my_t, p.dt);
        n += cnt;
        if (finished) break;
        t_base = t_next_base;
    }
    return n;
}

// Sequential marcher (one thread per ray) -- the reference's own control flow; used by the test-time
// marcher where each call only asks for a handful of samples per ray.
template <bool ONE_CASCADE = false, bool LUT = false, typename Emit>
__device__ __forceinline__ int march_ray_thread(float t, float t2, int max_emit, const RayConst& q, const MarchConst& c,
                                                const uint8_t* __restrict__ bitfield, Emit emit, float* t_after) {
    int s = 0;
    float t_last = t;
    while (t < t2 && s < max_emit) {
        const Probe p = probe_cell<ONE_CASCADE, LUT>(t, q, c, bitfield);
        if (p.occ) {
            emit(s, t, p.dt, p.x, p.y, p.z);
            t = __fadd_rn(p.dt, t);
            t_last = t;
            ++s;
        } else {
            do { t = march_next(t, c); } while (t < p.t_target);
        }
    }
    *t_after = t_last;
    return s;
}

}  // namespace mfn
