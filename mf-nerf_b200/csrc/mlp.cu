// Fully fused small MLPs (the reference's tcnn "FullyFusedMLP": no biases, fp16 weights stored first->last as
// row-major (out x in) matrices, ReLU hidden activations, output padded to 16; networks.py:48-57,69-79).
//
// v1 tensor-core path: one CTA = 128 samples, weights resident in shared memory for the whole (persistent) kernel,
// activations chained through mma.sync fragments in registers (forward) / shared memory (backward), fp32
// accumulation (tcnn accumulates in fp16 -- ours is the more precise of the two; tolerance stated in tests/).
// Backward = dgrad chain + wgrad (dW = dZ^T A) with the weight-gradient tiles accumulated in registers across all
// tiles of the CTA and flushed once with fp32 atomics.
//   FLOPs/sample: 2*(in*W + (h-1)*W*W + W*16) forward, 3x that forward+backward.
#include "mma_utils.cuh"
#include "field_internal.h"
#include "../../include/mfnerf_b200.h"

namespace mfn {

constexpr int kTile = 128;
constexpr int kMlpThreads = 128;
constexpr int kOutPad = 16;

template <int WIDTH, int K_IN>
struct MlpLayout {
    static constexpr int ldIn = K_IN + 8, ldH = WIDTH + 8, ldO = kOutPad + 8;
    static constexpr int w0 = 0;
    static constexpr int wh = WIDTH * ldIn;  // start of the hidden->hidden matrices
    __host__ __device__ static int wl(int n_hidden) { return wh + (n_hidden - 1) * WIDTH * ldH; }
    __host__ __device__ static int wend(int n_hidden) { return wl(n_hidden) + kOutPad * ldH; }
    static size_t fwd_bytes(int n_hidden) { return 2 * (size_t)(wend(n_hidden) + kTile * ldIn); }
    static size_t bwd_bytes(int n_hidden) { return 2 * (size_t)(wend(n_hidden) + kTile * ldIn + 2 * n_hidden * kTile * ldH + kTile * ldO); }
};

// copy a dense row-major [rows][cols] fp16 matrix from global into padded smem (cols % 8 == 0)
__device__ __forceinline__ void stage_matrix(__half* dst, int ld, const __half* __restrict__ src, int rows, int cols, int tid, int nthreads) {
    const int cpr = cols / 8;
    for (int i = tid; i < rows * cpr; i += nthreads) {
        const int r = i / cpr, c = (i % cpr) * 8;
        *reinterpret_cast<int4*>(dst + r * ld + c) = __ldg(reinterpret_cast<const int4*>(src + (size_t)r * cols + c));
    }
}
// same for a tile of rows [row0, row0+128) of an [n][cols] activation matrix; rows >= n are zero filled
__device__ __forceinline__ void stage_rows(__half* dst, int ld, const __half* __restrict__ src, int64_t row0, int64_t n, int cols, int tid, int src_stride = 0) {
    const int cpr = cols / 8;
    if (src_stride == 0) src_stride = cols;
    for (int i = tid; i < kTile * cpr; i += kMlpThreads) {
        const int r = i / cpr, c = (i % cpr) * 8;
        int4 v = make_int4(0, 0, 0, 0);
        if (row0 + r < n) v = __ldg(reinterpret_cast<const int4*>(src + (size_t)(row0 + r) * src_stride + c));
        *reinterpret_cast<int4*>(dst + r * ld + c) = v;
    }
}

template <int WIDTH, int K_IN>
__device__ __forceinline__ void stage_weights(__half* sW, const __half* __restrict__ W, int n_hidden, int tid) {
    using L = MlpLayout<WIDTH, K_IN>;
    stage_matrix(sW + L::w0, L::ldIn, W, WIDTH, K_IN, tid, kMlpThreads);
    const __half* g = W + WIDTH * K_IN;
    for (int l = 1; l < n_hidden; ++l, g += WIDTH * WIDTH) stage_matrix(sW + L::wh + (l - 1) * WIDTH * L::ldH, L::ldH, g, WIDTH, WIDTH, tid, kMlpThreads);
    stage_matrix(sW + L::wl(n_hidden), L::ldH, g, kOutPad, WIDTH, tid, kMlpThreads);
}

// c[NT][4] += A * B^T with B stored [n][k] (the forward weight layout)
template <int NT, int KS>
__device__ __forceinline__ void gemm_nk(float (&c)[NT][4], const uint32_t (&a)[KS][4], const __half* sB, int ldB, int lane) {
#pragma unroll
    for (int p = 0; p < NT / 2; ++p) {
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            uint32_t b[4];
            load_b(b, sB, ldB, p * 16, ks * 16, lane);
            mma_16816(c[2 * p], a[ks], b[0], b[1]);
            mma_16816(c[2 * p + 1], a[ks], b[2], b[3]);
        }
    }
}
// c[NT][4] += A * B with B stored [k][n] (dgrad: the same weight matrix read transposed)
template <int NT, int KS>
__device__ __forceinline__ void gemm_kn(float (&c)[NT][4], const uint32_t (&a)[KS][4], const __half* sB, int ldB, int lane) {
#pragma unroll
    for (int p = 0; p < NT / 2; ++p) {
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            uint32_t b[4];
            load_b_t(b, sB, ldB, p * 16, ks * 16, lane);
            mma_16816(c[2 * p], a[ks], b[0], b[1]);
            mma_16816(c[2 * p + 1], a[ks], b[2], b[3]);
        }
    }
}
template <int NT>
__device__ __forceinline__ void zero_acc(float (&c)[NT][4]) {
#pragma unroll
    for (int i = 0; i < NT; ++i) { c[i][0] = 0.f; c[i][1] = 0.f; c[i][2] = 0.f; c[i][3] = 0.f; }
}
// C fragments (m16 x NT*8, fp32) -> ReLU -> fp16 A fragments for the next layer (k = NT*8)
template <int NT>
__device__ __forceinline__ void relu_pack(const float (&c)[NT][4], uint32_t (&a)[NT / 2][4]) {
#pragma unroll
    for (int j = 0; j < NT / 2; ++j) {
        a[j][0] = pack_h2(fmaxf(c[2 * j][0], 0.f), fmaxf(c[2 * j][1], 0.f));
        a[j][1] = pack_h2(fmaxf(c[2 * j][2], 0.f), fmaxf(c[2 * j][3], 0.f));
        a[j][2] = pack_h2(fmaxf(c[2 * j + 1][0], 0.f), fmaxf(c[2 * j + 1][1], 0.f));
        a[j][3] = pack_h2(fmaxf(c[2 * j + 1][2], 0.f), fmaxf(c[2 * j + 1][3], 0.f));
    }
}
// write A fragments (rows row0+g / row0+g+8) to a row-major [n][cols] fp16 matrix in global memory
template <int KS>
__device__ __forceinline__ void store_frag_rows(__half* __restrict__ dst, int cols, int64_t row_g, int64_t n, const uint32_t (&a)[KS][4], int lane) {
    const int t2 = (lane & 3) * 2;
#pragma unroll
    for (int j = 0; j < KS; ++j) {
        if (row_g < n) {
            *reinterpret_cast<uint32_t*>(dst + (size_t)row_g * cols + 16 * j + t2) = a[j][0];
            *reinterpret_cast<uint32_t*>(dst + (size_t)row_g * cols + 16 * j + 8 + t2) = a[j][2];
        }
        if (row_g + 8 < n) {
            *reinterpret_cast<uint32_t*>(dst + (size_t)(row_g + 8) * cols + 16 * j + t2) = a[j][1];
            *reinterpret_cast<uint32_t*>(dst + (size_t)(row_g + 8) * cols + 16 * j + 8 + t2) = a[j][3];
        }
    }
}

__device__ __forceinline__ float out_activation(float x, int act) {
    if (act == MFN_ACT_SIGMOID) return 1.0f / (1.0f + __expf(-x));
    if (act == MFN_ACT_EXP) return __expf(x);
    return x;
}

template <int WIDTH, int K_IN>
__global__ void __launch_bounds__(kMlpThreads)
mlp_fwd_kernel(const __half* __restrict__ in, int in_stride, const __half* __restrict__ W, int n_hidden, int out_act, int64_t n_max,
               const int32_t* __restrict__ n_dev, __half* __restrict__ out, int out_stride, float* __restrict__ out_rgb32,
               __half* __restrict__ acts) {
    const int64_t n = n_dev ? min((int64_t)*n_dev, n_max) : n_max;
    using L = MlpLayout<WIDTH, K_IN>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half* sW = reinterpret_cast<__half*>(smem_raw);
    __half* sIn = sW + L::wend(n_hidden);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    stage_weights<WIDTH, K_IN>(sW, W, n_hidden, tid);
    const int64_t n_tiles = (n + kTile - 1) / kTile;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * kTile;
        __syncthreads();  // previous tile fully consumed (and weights staged)
        stage_rows(sIn, L::ldIn, in, row0, n, K_IN, tid, in_stride);
        __syncthreads();
#pragma unroll 1
        for (int mt = 0; mt < 2; ++mt) {
            const int m0 = warp * 32 + mt * 16;
            const int64_t row_g = row0 + m0 + (lane >> 2);
            uint32_t aIn[K_IN / 16][4];
#pragma unroll
            for (int ks = 0; ks < K_IN / 16; ++ks) load_a(aIn[ks], sIn, L::ldIn, m0, ks * 16, lane);
            float c[WIDTH / 8][4];
            zero_acc(c);
            gemm_nk<WIDTH / 8, K_IN / 16>(c, aIn, sW + L::w0, L::ldIn, lane);
            uint32_t aH[WIDTH / 16][4];
            relu_pack(c, aH);
            if (acts) store_frag_rows<WIDTH / 16>(acts, WIDTH, row_g, n, aH, lane);
            for (int l = 1; l < n_hidden; ++l) {
                zero_acc(c);
                gemm_nk<WIDTH / 8, WIDTH / 16>(c, aH, sW + L::wh + (l - 1) * WIDTH * L::ldH, L::ldH, lane);
                relu_pack(c, aH);
                if (acts) store_frag_rows<WIDTH / 16>(acts + (size_t)l * n_max * WIDTH, WIDTH, row_g, n, aH, lane);
            }
            float co[2][4];
            zero_acc(co);
            gemm_nk<2, WIDTH / 16>(co, aH, sW + L::wl(n_hidden), L::ldH, lane);
            const int t2 = (lane & 3) * 2;
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                const uint32_t lo = pack_h2(out_activation(co[nt][0], out_act), out_activation(co[nt][1], out_act));
                const uint32_t hi = pack_h2(out_activation(co[nt][2], out_act), out_activation(co[nt][3], out_act));
                if (row_g < n) *reinterpret_cast<uint32_t*>(out + (size_t)row_g * out_stride + nt * 8 + t2) = lo;
                if (row_g + 8 < n) *reinterpret_cast<uint32_t*>(out + (size_t)(row_g + 8) * out_stride + nt * 8 + t2) = hi;
                if (out_rgb32 && nt == 0 && t2 < 4) {  // fp32 copy of outputs 0..2 (the values as rounded to fp16)
                    const float2 flo = __half22float2(*reinterpret_cast<const __half2*>(&lo)), fhi = __half22float2(*reinterpret_cast<const __half2*>(&hi));
                    if (row_g < n) { out_rgb32[3 * row_g + t2] = flo.x; if (t2 == 0) out_rgb32[3 * row_g + 1] = flo.y; }
                    if (row_g + 8 < n) { out_rgb32[3 * (row_g + 8) + t2] = fhi.x; if (t2 == 0) out_rgb32[3 * (row_g + 8) + 1] = fhi.y; }
                }
            }
        }
    }
}

// masks a C tile with relu'(act) read from smem, rounds to fp16, writes dZ to smem and packs it as A fragments
template <int NT>
__device__ __forceinline__ void relu_bwd_pack(const float (&c)[NT][4], const __half* sAct, __half* sDZ, int ld, int m0, int lane, uint32_t (&a)[NT / 2][4]) {
    const int g = lane >> 2, t2 = (lane & 3) * 2;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const int col = nt * 8 + t2;
        const __half2 a_lo = *reinterpret_cast<const __half2*>(sAct + (m0 + g) * ld + col);
        const __half2 a_hi = *reinterpret_cast<const __half2*>(sAct + (m0 + g + 8) * ld + col);
        const uint32_t lo = pack_h2(__low2float(a_lo) > 0.f ? c[nt][0] : 0.f, __high2float(a_lo) > 0.f ? c[nt][1] : 0.f);
        const uint32_t hi = pack_h2(__low2float(a_hi) > 0.f ? c[nt][2] : 0.f, __high2float(a_hi) > 0.f ? c[nt][3] : 0.f);
        *reinterpret_cast<uint32_t*>(sDZ + (m0 + g) * ld + col) = lo;
        *reinterpret_cast<uint32_t*>(sDZ + (m0 + g + 8) * ld + col) = hi;
        a[nt / 2][(nt & 1) * 2 + 0] = lo;
        a[nt / 2][(nt & 1) * 2 + 1] = hi;
    }
}

// dW tile accumulation: acc[MT][NT][4] += dZ^T (m = out feature, k = sample) * A_prev (k = sample, n = in feature)
template <int MT, int NT>
__device__ __forceinline__ void wgrad_accum(float (&acc)[MT][NT][4], const __half* sDZ, int ldz, int m_base, const __half* sA, int lda, int n_base, int lane) {
#pragma unroll 2
    for (int ks = 0; ks < kTile / 16; ++ks) {
        uint32_t a[MT][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) load_a_t(a[mt], sDZ, ldz, m_base + mt * 16, ks * 16, lane);
#pragma unroll
        for (int p = 0; p < NT / 2; ++p) {
            uint32_t b[4];
            load_b_t(b, sA, lda, n_base + p * 16, ks * 16, lane);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                mma_16816(acc[mt][2 * p], a[mt], b[0], b[1]);
                mma_16816(acc[mt][2 * p + 1], a[mt], b[2], b[3]);
            }
        }
    }
}
template <int MT, int NT>
__device__ __forceinline__ void wgrad_flush(float (&acc)[MT][NT][4], float* __restrict__ dW, int cols, int m_base, int n_base, int lane, bool reset) {
    const int g = lane >> 2, t2 = (lane & 3) * 2;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            float* p = dW + (size_t)(m_base + mt * 16 + g) * cols + n_base + nt * 8 + t2;
            atomicAdd(p, acc[mt][nt][0]); atomicAdd(p + 1, acc[mt][nt][1]);
            atomicAdd(p + 8 * cols, acc[mt][nt][2]); atomicAdd(p + 8 * cols + 1, acc[mt][nt][3]);
            if (reset) { acc[mt][nt][0] = 0.f; acc[mt][nt][1] = 0.f; acc[mt][nt][2] = 0.f; acc[mt][nt][3] = 0.f; }
        }
}

template <int WIDTH, int K_IN, bool FLUSH>
__global__ void __launch_bounds__(kMlpThreads)
mlp_bwd_kernel(const __half* __restrict__ dOut, const __half* __restrict__ in, int in_stride, const __half* __restrict__ acts,
               const __half* __restrict__ outv, int out_stride, const __half* __restrict__ W, int n_hidden, int out_act, int64_t n_max,
               const int32_t* __restrict__ n_dev, __half* __restrict__ dIn, int din_stride, float* __restrict__ dW) {
    const int64_t n = n_dev ? min((int64_t)*n_dev, n_max) : n_max;
    using L = MlpLayout<WIDTH, K_IN>;
    constexpr int MT = WIDTH / 64;  // m16 tiles of dW rows per warp (4 warps split WIDTH rows)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half* sW = reinterpret_cast<__half*>(smem_raw);
    __half* sIn = sW + L::wend(n_hidden);
    __half* sAct = sIn + kTile * L::ldIn;                 // [n_hidden][128][ldH]
    __half* sDZ = sAct + n_hidden * kTile * L::ldH;       // [n_hidden][128][ldH]
    __half* sDZo = sDZ + n_hidden * kTile * L::ldH;       // [128][ldO]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    stage_weights<WIDTH, K_IN>(sW, W, n_hidden, tid);
    float accF[MT][K_IN / 8][4];      // dW of the first matrix   [WIDTH][K_IN]
    float accH[MT][WIDTH / 8][4];     // dW of the hidden matrix  [WIDTH][WIDTH] (n_hidden == 2)
    float accL[1][WIDTH / 32][4];     // dW of the last matrix    [16][WIDTH], n split over the 4 warps
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) { zero_acc(accF[mt]); zero_acc(accH[mt]); }
    zero_acc(accL[0]);
    const int64_t n_tiles = (n + kTile - 1) / kTile;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * kTile;
        __syncthreads();
        stage_rows(sIn, L::ldIn, in, row0, n, K_IN, tid, in_stride);
        for (int l = 0; l < n_hidden; ++l) stage_rows(sAct + l * kTile * L::ldH, L::ldH, acts + (size_t)l * n_max * WIDTH, row0, n, WIDTH, tid);
        // dL/dz of the output layer: dOut (* sigmoid' when the output activation is a sigmoid)
        for (int i = tid; i < kTile * kOutPad / 2; i += kMlpThreads) {
            const int r = i / (kOutPad / 2), c = (i % (kOutPad / 2)) * 2;
            __half2 v = __floats2half2_rn(0.f, 0.f);
            if (row0 + r < n) {
                v = *reinterpret_cast<const __half2*>(dOut + (size_t)(row0 + r) * kOutPad + c);
                if (out_act == MFN_ACT_SIGMOID) {
                    const float2 y = __half22float2(*reinterpret_cast<const __half2*>(outv + (size_t)(row0 + r) * out_stride + c));
                    const float2 g = __half22float2(v);
                    v = __floats2half2_rn(g.x * y.x * (1.f - y.x), g.y * y.y * (1.f - y.y));
                } else if (out_act == MFN_ACT_EXP) {
                    const float2 y = __half22float2(*reinterpret_cast<const __half2*>(outv + (size_t)(row0 + r) * out_stride + c));
                    const float2 g = __half22float2(v);
                    v = __floats2half2_rn(g.x * y.x, g.y * y.y);
                }
            }
            *reinterpret_cast<__half2*>(sDZo + r * L::ldO + c) = v;
        }
        __syncthreads();
        // ---- dgrad chain, one m16 tile at a time ----
#pragma unroll 1
        for (int mt = 0; mt < 2; ++mt) {
            const int m0 = warp * 32 + mt * 16;
            uint32_t aO[1][4];
            load_a(aO[0], sDZo, L::ldO, m0, 0, lane);
            float c[WIDTH / 8][4];
            zero_acc(c);
            gemm_kn<WIDTH / 8, 1>(c, aO, sW + L::wl(n_hidden), L::ldH, lane);
            uint32_t aZ[WIDTH / 16][4];
            relu_bwd_pack<WIDTH / 8>(c, sAct + (n_hidden - 1) * kTile * L::ldH, sDZ + (n_hidden - 1) * kTile * L::ldH, L::ldH, m0, lane, aZ);
            for (int l = n_hidden - 1; l >= 1; --l) {
                zero_acc(c);
                gemm_kn<WIDTH / 8, WIDTH / 16>(c, aZ, sW + L::wh + (l - 1) * WIDTH * L::ldH, L::ldH, lane);
                relu_bwd_pack<WIDTH / 8>(c, sAct + (l - 1) * kTile * L::ldH, sDZ + (l - 1) * kTile * L::ldH, L::ldH, m0, lane, aZ);
            }
            if (dIn) {
                float ci[K_IN / 8][4];
                zero_acc(ci);
                gemm_kn<K_IN / 8, WIDTH / 16>(ci, aZ, sW + L::w0, L::ldIn, lane);
                const int64_t row_g = row0 + m0 + (lane >> 2);
                const int t2 = (lane & 3) * 2;
#pragma unroll
                for (int nt = 0; nt < K_IN / 8; ++nt) {
                    if (row_g < n) *reinterpret_cast<uint32_t*>(dIn + (size_t)row_g * din_stride + nt * 8 + t2) = pack_h2(ci[nt][0], ci[nt][1]);
                    if (row_g + 8 < n) *reinterpret_cast<uint32_t*>(dIn + (size_t)(row_g + 8) * din_stride + nt * 8 + t2) = pack_h2(ci[nt][2], ci[nt][3]);
                }
            }
        }
        __syncthreads();
        // ---- wgrad: every warp owns WIDTH/4 output rows of the big matrices and WIDTH/4 input columns of the last one ----
        wgrad_accum<MT, K_IN / 8>(accF, sDZ, L::ldH, warp * (WIDTH / 4), sIn, L::ldIn, 0, lane);
        if (n_hidden == 2) wgrad_accum<MT, WIDTH / 8>(accH, sDZ + kTile * L::ldH, L::ldH, warp * (WIDTH / 4), sAct, L::ldH, 0, lane);
        wgrad_accum<1, WIDTH / 32>(accL, sDZo, L::ldO, 0, sAct + (n_hidden - 1) * kTile * L::ldH, L::ldH, warp * (WIDTH / 4), lane);
        if (FLUSH) {
            wgrad_flush<MT, K_IN / 8>(accF, dW, K_IN, warp * (WIDTH / 4), 0, lane, true);
            if (n_hidden == 2) wgrad_flush<MT, WIDTH / 8>(accH, dW + WIDTH * K_IN, WIDTH, warp * (WIDTH / 4), 0, lane, true);
            wgrad_flush<1, WIDTH / 32>(accL, dW + WIDTH * K_IN + (n_hidden - 1) * WIDTH * WIDTH, WIDTH, 0, warp * (WIDTH / 4), lane, true);
        }
    }
    if (!FLUSH) {
        wgrad_flush<MT, K_IN / 8>(accF, dW, K_IN, warp * (WIDTH / 4), 0, lane, false);
        if (n_hidden == 2) wgrad_flush<MT, WIDTH / 8>(accH, dW + WIDTH * K_IN, WIDTH, warp * (WIDTH / 4), 0, lane, false);
        wgrad_flush<1, WIDTH / 32>(accL, dW + WIDTH * K_IN + (n_hidden - 1) * WIDTH * WIDTH, WIDTH, 0, warp * (WIDTH / 4), lane, false);
    }
}

template <typename K>
static int max_ctas_per_sm(K kernel, size_t smem) {
    int nb = 0;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, kMlpThreads, smem);
    return nb < 1 ? 1 : nb;
}

template <int WIDTH, int K_IN>
static int launch_fwd(const MlpFwdArgs& a, cudaStream_t st) {
    const size_t smem = MlpLayout<WIDTH, K_IN>::fwd_bytes(a.n_hidden);
    static int per_sm = max_ctas_per_sm(mlp_fwd_kernel<WIDTH, K_IN>, MlpLayout<WIDTH, K_IN>::fwd_bytes(2));
    const int64_t tiles = ceil_div(a.n_max, kTile);
    const int grid = (int)(tiles < (int64_t)per_sm * kNumSMs ? tiles : (int64_t)per_sm * kNumSMs);
    ProfScope ps(a.tag ? a.tag : "mlp_fwd", st);
    mlp_fwd_kernel<WIDTH, K_IN><<<grid, kMlpThreads, smem, st>>>(a.in, a.in_stride, a.W, a.n_hidden, a.out_act, a.n_max, a.n_dev, a.out,
                                                              a.out_stride ? a.out_stride : kOutPad, a.out_rgb32, a.acts);
    return check_launch("mfn_mlp_fwd", st);
}
template <int WIDTH, int K_IN, bool FLUSH>
static int launch_bwd(const MlpBwdArgs& a, cudaStream_t st) {
    const size_t smem = MlpLayout<WIDTH, K_IN>::bwd_bytes(a.n_hidden);
    static int per_sm = max_ctas_per_sm(mlp_bwd_kernel<WIDTH, K_IN, FLUSH>, MlpLayout<WIDTH, K_IN>::bwd_bytes(2));
    const int64_t tiles = ceil_div(a.n_max, kTile);
    const int grid = (int)(tiles < (int64_t)per_sm * kNumSMs ? tiles : (int64_t)per_sm * kNumSMs);
    ProfScope ps(a.tag ? a.tag : "mlp_bwd", st);
    mlp_bwd_kernel<WIDTH, K_IN, FLUSH><<<grid, kMlpThreads, smem, st>>>(a.dOut, a.in, a.in_stride, a.acts, a.outv, a.out_stride ? a.out_stride : kOutPad, a.W,
                                                                     a.n_hidden, a.out_act, a.n_max, a.n_dev, a.dIn, a.din_stride ? a.din_stride : K_IN, a.dW);
    return check_launch("mfn_mlp_bwd", st);
}

static bool mlp_cfg_ok(int in_dim, int width, int n_hidden, const char* name) {
    if ((in_dim != 16 && in_dim != 32 && in_dim != 64) || (width != 64 && width != 128) || n_hidden < 1 || n_hidden > 2) {
        set_error("%s: unsupported MLP shape in=%d width=%d hidden_layers=%d (in: 16/32/64, width: 64/128, hidden layers: 1-2)", name, in_dim, width, n_hidden);
        return false;
    }
    return true;
}

}  // namespace mfn

using namespace mfn;

extern "C" int64_t mfn_mlp_param_count(int in_dim, int width, int n_hidden) {
    return (int64_t)width * in_dim + (int64_t)(n_hidden - 1) * width * width + (int64_t)kOutPad * width;
}

#define MFN_MLP_DISPATCH(CALL64, CALL128)                                  \
    if (width == 64) {                                                     \
        if (in_dim == 16) { CALL64(16) } else if (in_dim == 32) { CALL64(32) } else { CALL64(64) } \
    } else {                                                               \
        if (in_dim == 16) { CALL128(16) } else if (in_dim == 32) { CALL128(32) } else { CALL128(64) } \
    }

namespace mfn {
int mlp_forward(const MlpFwdArgs& a, int in_dim, int width, cudaStream_t st) {
    if (!mlp_cfg_ok(in_dim, width, a.n_hidden, "mfn_mlp_fwd")) return MFN_ERR_ARG;
    if (a.n_max <= 0) return MFN_OK;
#define F64(K) return launch_fwd<64, K>(a, st);
#define F128(K) return launch_fwd<128, K>(a, st);
    MFN_MLP_DISPATCH(F64, F128)
#undef F64
#undef F128
}
int mlp_backward(const MlpBwdArgs& a, int in_dim, int width, cudaStream_t st) {
    if (!mlp_cfg_ok(in_dim, width, a.n_hidden, "mfn_mlp_bwd")) return MFN_ERR_ARG;
    if (a.n_max <= 0) return MFN_OK;
#define B64(K) return launch_bwd<64, K, false>(a, st);
#define B128(K) return launch_bwd<128, K, true>(a, st);
    MFN_MLP_DISPATCH(B64, B128)
#undef B64
#undef B128
}
}  // namespace mfn

extern "C" int mfn_mlp_fwd(const void* in, const void* weights, int in_dim, int width, int n_hidden, int out_act, int64_t n, void* out,
                           void* acts, void* stream) {
    if (n < 0) { set_error("mfn_mlp_fwd: bad n"); return MFN_ERR_ARG; }
    if (n > 0 && (!in || !weights || !out)) { set_error("mfn_mlp_fwd: null pointer"); return MFN_ERR_ARG; }
    MlpFwdArgs a{};
    a.in = (const __half*)in; a.W = (const __half*)weights; a.n_hidden = n_hidden; a.out_act = out_act; a.n_max = n;
    a.out = (__half*)out; a.acts = (__half*)acts;
    return mlp_forward(a, in_dim, width, (cudaStream_t)stream);
}

extern "C" int mfn_mlp_bwd(const void* dL_dout, const void* in, const void* acts, const void* out, const void* weights, int in_dim, int width,
                           int n_hidden, int out_act, int64_t n, void* dL_din, float* dW, void* stream) {
    if (n < 0) { set_error("mfn_mlp_bwd: bad n"); return MFN_ERR_ARG; }
    if (n > 0 && (!dL_dout || !in || !acts || !weights || !dW || (out_act != MFN_ACT_NONE && !out))) { set_error("mfn_mlp_bwd: null pointer"); return MFN_ERR_ARG; }
    MlpBwdArgs a{};
    a.dOut = (const __half*)dL_dout; a.in = (const __half*)in; a.acts = (const __half*)acts; a.outv = (const __half*)out; a.W = (const __half*)weights;
    a.n_hidden = n_hidden; a.out_act = out_act; a.n_max = n; a.dIn = (__half*)dL_din; a.dW = dW;
    return mlp_backward(a, in_dim, width, (cudaStream_t)stream);
}
