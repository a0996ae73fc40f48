// Data-parallel gradient exchange + optimiser as ONE kernel over NVLink / NVSwitch peer memory (the reference trains
// with DDP, train.py:284-285: bucketed NCCL all-reduce of the flat fp32 gradients, then a replicated optimiser step on every rank).
//
// Here every rank owns the shard [shard_begin, shard_begin + n) of the flat parameter vector (ZeRO-1 style) and ONE kernel per rank does,
// element by element of its shard:
//     g      = sum over ranks of grads_r[j]            N peer loads summed in rank order (the default: measured faster at 2, 4 and 8 B200s,
//                                                      DESIGN.md section 6) -- or one multimem.ld_reduce when the caller passes the multicast
//                                                      addresses (the NVSwitch adds the N copies in the fabric)
//     Adam on the fp32 master copy (local HBM: p, m, v of the shard only), unscale and skip-on-overflow like optim.cu
//     shadow_r[j] <- fp16(p) on every rank             N peer stores (or one multimem.st): the fp16 parameters every field kernel gathers from
// i.e. reduce-scatter + optimiser + all-gather in one pass, 4 + 2 bytes per parameter over the links instead of NCCL's three collectives
// with the optimiser in between (every rank clears its own gradient buffer locally after the closing barrier: clearing all copies from
// the owner -- tried first -- doubles the store traffic on the links).  The overflow flags of all ranks are OR-ed by every CTA (N loads) so that
// every replica takes the same skip-or-step decision (GradScaler semantics under DDP).
//
// Synchronisation is the caller's: every rank's gradient buffer must be complete before the kernel starts anywhere, and every rank's kernel
// must have finished before any rank reads its shadow or accumulates new gradients -- the engine brackets the launch with two device-side
// barriers over the same symmetric-memory allocation (torch.distributed._symmetric_memory, engine.py:_dp_setup).
#include "common.cuh"
#include <cuda_fp16.h>
#include <stdlib.h>
#include "../../include/mfnerf_b200.h"

namespace mfn {

constexpr int kMaxRanks = 16;
struct PeerPtrs { unsigned long long p[kMaxRanks]; };

__device__ __forceinline__ float4 mc_ld_reduce_add(const float* mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st_v4(void* mc, float a, float b, float c, float d) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ void adam4(float4& p, float4& m, float4& v, const float4& g, float gs, float lr, float b1, float b2, float eps, float bc1, float bc2) {
    float* P = &p.x; float* M = &m.x; float* V = &v.x; const float* G = &g.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {     // the arithmetic of optim.cu's adam_kernel, expression for expression
        const float gr = G[k] * gs;
        M[k] = b1 * M[k] + (1.f - b1) * gr;
        V[k] = b2 * V[k] + (1.f - b2) * gr * gr;
        P[k] -= lr * (M[k] / bc1) / (sqrtf(V[k] / bc2) + eps);
    }
}

// One "chunk" = 8 parameters: two 16-byte gradient reductions, one 16-byte shadow store.  A thread handles U chunks per iteration and
// issues ALL their loads (local p, m, v and the W peers' / the multicast gradients: ~16 sixteen-byte requests) before consuming the first,
// so that a SMALL grid -- one CTA per SM by default, like a collective library's channels -- keeps the links busy and leaves the SMs to the
// marching front that overlaps the exchange (a full-size grid made both slower).  W = number of ranks when it is 2, 4 or 8, 0 = any other
// count (runtime loop, U = 1).  MC: multimem path.
template <bool MC, int W, int U>
__global__ void __launch_bounds__(256)
dp_exchange_adam_kernel(const PeerPtrs grads, const PeerPtrs shadow, const PeerPtrs flags, float* grads_mc, __half* shadow_mc, int world,
                        float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, int64_t shard_begin, int64_t n, const float* __restrict__ lr_dev,
                        float beta1, float beta2, float eps, const float* __restrict__ amp, int32_t* __restrict__ skip_out) {
    // one thread per CTA reads the ranks' overflow flags (every thread doing so put 300 k uncached loads of ONE peer address on the links:
    // that, not the data, was most of the first version's 194 us)
    // (lane r of warp 0 reads rank r's flag: the N peer round trips overlap instead of queueing behind each other)
    __shared__ int s_bad;
    if (threadIdx.x < 32) {
        int bad = 0;
        for (int r = (int)threadIdx.x; r < world; r += 32) bad |= *reinterpret_cast<const volatile int32_t*>(flags.p[r]);
        bad = __any_sync(0xffffffffu, bad != 0);
        if (threadIdx.x == 0) s_bad = bad;
    }
    __syncthreads();
    const bool skip = s_bad != 0;
    if (blockIdx.x == 0 && threadIdx.x == 0 && skip_out) *skip_out = skip ? 1 : 0;
    if (skip) return;           // nothing moves anywhere; the caller clears its own gradient buffer after the closing barrier either way
    const float lr = lr_dev[0], bc1 = amp[4], bc2 = amp[5], gs = (1.f / (float)world) / amp[0];
    const int64_t n8 = n / 8;
    constexpr int WW = W > 0 ? W : 1;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n8; i0 += stride * U) {
        float4 pp[U][2], mm[U][2], vv[U][2], g[U][2];
        float4 a[U][WW][2];
#pragma unroll
        for (int u = 0; u < U; ++u) {      // chunk u of this thread: i0 + u * stride (coalesced across the warp for every u)
            const int64_t i = i0 + u * stride;
            if (i < n8) {
                const int64_t j = shard_begin + 8 * i;
                pp[u][0] = reinterpret_cast<const float4*>(p)[2 * i]; pp[u][1] = reinterpret_cast<const float4*>(p)[2 * i + 1];
                mm[u][0] = reinterpret_cast<const float4*>(m)[2 * i]; mm[u][1] = reinterpret_cast<const float4*>(m)[2 * i + 1];
                vv[u][0] = reinterpret_cast<const float4*>(v)[2 * i]; vv[u][1] = reinterpret_cast<const float4*>(v)[2 * i + 1];
                if (MC) { g[u][0] = mc_ld_reduce_add(grads_mc + j); g[u][1] = mc_ld_reduce_add(grads_mc + j + 4); }
                else if (W > 0) {
#pragma unroll
                    for (int r = 0; r < WW; ++r) { const float4* gp = reinterpret_cast<const float4*>(grads.p[r]) + (j >> 2); a[u][r][0] = __ldcs(gp); a[u][r][1] = __ldcs(gp + 1); }
                } else {
                    g[u][0] = make_float4(0.f, 0.f, 0.f, 0.f); g[u][1] = g[u][0];
                    for (int r = 0; r < world; ++r) {
                        const float4* gp = reinterpret_cast<const float4*>(grads.p[r]) + (j >> 2);
                        const float4 x = __ldcs(gp), y = __ldcs(gp + 1);
                        g[u][0].x += x.x; g[u][0].y += x.y; g[u][0].z += x.z; g[u][0].w += x.w; g[u][1].x += y.x; g[u][1].y += y.y; g[u][1].z += y.z; g[u][1].w += y.w;
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + u * stride;
            if (i >= n8) break;
            const int64_t j = shard_begin + 8 * i;
            if (!MC && W > 0) {
                g[u][0] = a[u][0][0]; g[u][1] = a[u][0][1];
#pragma unroll
                for (int r = 1; r < WW; ++r) {      // rank order: the same sum on every run
                    g[u][0].x += a[u][r][0].x; g[u][0].y += a[u][r][0].y; g[u][0].z += a[u][r][0].z; g[u][0].w += a[u][r][0].w;
                    g[u][1].x += a[u][r][1].x; g[u][1].y += a[u][r][1].y; g[u][1].z += a[u][r][1].z; g[u][1].w += a[u][r][1].w;
                }
            }
            adam4(pp[u][0], mm[u][0], vv[u][0], g[u][0], gs, lr, beta1, beta2, eps, bc1, bc2);
            adam4(pp[u][1], mm[u][1], vv[u][1], g[u][1], gs, lr, beta1, beta2, eps, bc1, bc2);
            reinterpret_cast<float4*>(p)[2 * i] = pp[u][0]; reinterpret_cast<float4*>(p)[2 * i + 1] = pp[u][1];
            reinterpret_cast<float4*>(m)[2 * i] = mm[u][0]; reinterpret_cast<float4*>(m)[2 * i + 1] = mm[u][1];
            reinterpret_cast<float4*>(v)[2 * i] = vv[u][0]; reinterpret_cast<float4*>(v)[2 * i + 1] = vv[u][1];
            const __half2 h0 = __floats2half2_rn(pp[u][0].x, pp[u][0].y), h1 = __floats2half2_rn(pp[u][0].z, pp[u][0].w),
                          h2 = __floats2half2_rn(pp[u][1].x, pp[u][1].y), h3 = __floats2half2_rn(pp[u][1].z, pp[u][1].w);
            const float f0 = __uint_as_float(*reinterpret_cast<const uint32_t*>(&h0)), f1 = __uint_as_float(*reinterpret_cast<const uint32_t*>(&h1)),
                        f2 = __uint_as_float(*reinterpret_cast<const uint32_t*>(&h2)), f3 = __uint_as_float(*reinterpret_cast<const uint32_t*>(&h3));
            if (MC) mc_st_v4(shadow_mc + j, f0, f1, f2, f3);      // (a store moves bits: 8 halves travel as 4 "floats")
            else {
#pragma unroll
                for (int r = 0; r < (W > 0 ? W : kMaxRanks); ++r)
                    if (W > 0 || r < world) *reinterpret_cast<float4*>(reinterpret_cast<__half*>(shadow.p[r]) + j) = make_float4(f0, f1, f2, f3);
            }
        }
    }
    // no fence here: the stores are performed before the grid completes, and the barrier kernel that follows on the same stream publishes
    // them to the peers (a system-scope fence per thread cost more than the rest of the kernel)
}

}  // namespace mfn

using namespace mfn;

extern "C" int mfn_dp_exchange_adam(int world, const uint64_t* grads_ptrs_host, const uint64_t* shadow_ptrs_host, const uint64_t* flag_ptrs_host,
                                    uint64_t grads_mc, uint64_t shadow_mc, float* params_shard, float* exp_avg_shard, float* exp_avg_sq_shard,
                                    int64_t shard_begin, int64_t n, const float* lr_dev, float beta1, float beta2, float eps, const float* amp_state,
                                    int32_t* skip_out, void* stream) {
    if (world < 1 || world > kMaxRanks || n < 0 || shard_begin < 0 || (n % 8) || (shard_begin % 8)) {
        set_error("mfn_dp_exchange_adam: bad argument (1 <= world <= 16, shard begin and size multiples of 8)"); return MFN_ERR_ARG;
    }
    if (n == 0) return MFN_OK;
    if (!grads_ptrs_host || !shadow_ptrs_host || !flag_ptrs_host || !params_shard || !exp_avg_shard || !exp_avg_sq_shard || !lr_dev || !amp_state) {
        set_error("mfn_dp_exchange_adam: null pointer"); return MFN_ERR_ARG;
    }
    if ((grads_mc != 0) != (shadow_mc != 0)) { set_error("mfn_dp_exchange_adam: give both multicast addresses or neither"); return MFN_ERR_ARG; }
    PeerPtrs g{}, s{}, f{};
    for (int r = 0; r < world; ++r) {
        g.p[r] = grads_ptrs_host[r]; s.p[r] = shadow_ptrs_host[r]; f.p[r] = flag_ptrs_host[r];
        if (!g.p[r] || !s.p[r] || !f.p[r] || (g.p[r] & 15) || (s.p[r] & 15)) { set_error("mfn_dp_exchange_adam: peer pointers must be non-null and 16-byte aligned"); return MFN_ERR_ARG; }
    }
    if ((grads_mc & 15) || (shadow_mc & 15) || ((uintptr_t)params_shard & 15) || ((uintptr_t)exp_avg_shard & 15) || ((uintptr_t)exp_avg_sq_shard & 15)) {
        set_error("mfn_dp_exchange_adam: buffers must be 16-byte aligned"); return MFN_ERR_ARG;
    }
    static int ctas = 0;
    if (ctas == 0) { const char* e = getenv("MFN_DPX_CTAS"); ctas = e ? atoi(e) : kNumSMs; if (ctas < 1) ctas = kNumSMs; }
    int64_t blocks = ceil_div(n / 8, 256);
    if (blocks > ctas) blocks = ctas;
    if (blocks < 1) blocks = 1;
    ProfScope ps("adam", (cudaStream_t)stream);
#define MFN_DPX(MC_, W_, U_) dp_exchange_adam_kernel<MC_, W_, U_><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(g, s, f, (float*)grads_mc, (__half*)shadow_mc, \
        world, params_shard, exp_avg_shard, exp_avg_sq_shard, shard_begin, n, lr_dev, beta1, beta2, eps, amp_state, skip_out)
    if (grads_mc) MFN_DPX(true, 0, 4);
    else if (world == 2) MFN_DPX(false, 2, 4);
    else if (world == 4) MFN_DPX(false, 4, 2);
    else if (world == 8) MFN_DPX(false, 8, 1);
    else MFN_DPX(false, 0, 1);
#undef MFN_DPX
    return check_launch("mfn_dp_exchange_adam", (cudaStream_t)stream);
}
