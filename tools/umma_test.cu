// Validates the hand-built tcgen05 descriptors of mf-nerf_b200/csrc/umma.cuh on a B200: small GEMMs in every operand-major
// combination the fused field kernels use, compared against a CPU reference; also reports the TMEM lane mapping for M = 64.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 tools/umma_test.cu -o tools/bin/umma_test
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../mf-nerf_b200/csrc/umma.cuh"

using namespace mfn::umma;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

// A: logical [M x K], B: logical [N x K] (both given row-major in global, fp16).  D[lane][n] for all 128 TMEM lanes is returned.
__global__ void __launch_bounds__(128) gemm_kernel(const __half* A, const __half* B, int M, int N, int K, int a_mn, int b_mn, float* D) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    unsigned char* sA = smem;
    unsigned char* sB = smem + 32768;
    const int tid = threadIdx.x, warp = tid >> 5;
    // stage: K-major operand = tile [MN rows x K cols]; MN-major operand = tile [K rows x MN cols]
    for (int i = tid; i < M * K; i += 128) {
        const int m = i / K, k = i % K;
        const int off = a_mn ? tile_off(k, m, M) : tile_off(m, k, K);
        *reinterpret_cast<__half*>(sA + off) = A[i];
    }
    for (int i = tid; i < N * K; i += 128) {
        const int n = i / K, k = i % K;
        const int off = b_mn ? tile_off(k, n, N) : tile_off(n, k, K);
        *reinterpret_cast<__half*>(sB + off) = B[i];
    }
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tmem_base_s, 128);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = idesc_f16(M, N, a_mn, b_mn);
        for (int k0 = 0; k0 < K; k0 += 16) {
            const uint64_t ad = a_mn ? desc_mnmajor(smem_u32(sA), M, k0) : desc_kmajor(smem_u32(sA), K, k0);
            const uint64_t bd = b_mn ? desc_mnmajor(smem_u32(sB), N, k0) : desc_kmajor(smem_u32(sB), K, k0);
            mma_f16_ss(tbase, ad, bd, idesc, k0 > 0);
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t r[8];
        tmem_ld_x8(tmem_addr(tbase, warp * 32, c0), r);
        tmem_ld_wait();
        for (int j = 0; j < 8; ++j) D[tid * N + c0 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 128);
}

static bool run(int M, int N, int K, int a_mn, int b_mn) {
    std::vector<__half> A(M * K), B(N * K);
    std::vector<float> Af(M * K), Bf(N * K), ref(M * N), D(128 * N);
    uint32_t s = 12345u + M * 7 + N * 3 + K + a_mn * 2 + b_mn;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((float)(s >> 9) / 4194304.f - 1.f); };
    for (int i = 0; i < M * K; ++i) { A[i] = __float2half(rnd()); Af[i] = __half2float(A[i]); }
    for (int i = 0; i < N * K; ++i) { B[i] = __float2half(rnd()); Bf[i] = __half2float(B[i]); }
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double a = 0; for (int k = 0; k < K; ++k) a += (double)Af[m * K + k] * Bf[n * K + k]; ref[m * N + n] = (float)a; }
    __half *dA, *dB; float* dD;
    CK(cudaMalloc(&dA, M * K * 2)); CK(cudaMalloc(&dB, N * K * 2)); CK(cudaMalloc(&dD, 128 * N * 4));
    CK(cudaMemcpy(dA, A.data(), M * K * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), N * K * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0, 128 * N * 4));
    CK(cudaFuncSetAttribute(gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    gemm_kernel<<<1, 128, 65536>>>(dA, dB, M, N, K, a_mn, b_mn, dD);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, 128 * N * 4, cudaMemcpyDeviceToHost));
    // find, for each logical row m, the TMEM lane that holds it
    int bad = 0; std::vector<int> lane_of(M, -1);
    for (int m = 0; m < M; ++m) {
        for (int lane = 0; lane < 128 && lane_of[m] < 0; ++lane) {
            bool ok = true;
            for (int n = 0; n < N && ok; ++n) ok = fabsf(D[lane * N + n] - ref[m * N + n]) <= 1e-3f + 1e-3f * fabsf(ref[m * N + n]);
            if (ok) lane_of[m] = lane;
        }
        bad += lane_of[m] < 0;
    }
    bool identity = true; for (int m = 0; m < M; ++m) identity &= lane_of[m] == m;
    printf("M=%3d N=%3d K=%3d A:%s B:%s  rows matched %d/%d  lane map %s", M, N, K, a_mn ? "MN" : "K ", b_mn ? "MN" : "K ", M - bad, M, identity ? "identity" : "");
    if (!identity && bad == 0) { printf("["); for (int m = 0; m < M; m += 8) printf("%d->%d ", m, lane_of[m]); printf("]"); }
    if (bad) printf("  (D[0][0..3] = %g %g %g %g, ref %g %g %g %g)", D[0], D[1], D[2], D[3], ref[0], ref[1], ref[2], ref[3]);
    printf("\n");
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return bad == 0;
}

int main() {
    bool ok = true;
    ok &= run(128, 64, 32, 0, 0);   // forward layer 1: feats[128x32] . W1[64x32]^T
    ok &= run(128, 16, 64, 0, 0);   // forward output layer
    ok &= run(128, 64, 64, 0, 0);   // forward hidden layer
    ok &= run(128, 64, 16, 0, 1);   // dgrad through the output layer: dZo[128x16] . W[16x64]   (B = W^T logical [64 x 16], stored as W)
    ok &= run(128, 32, 64, 0, 1);   // dgrad to the input: dZ[128x64] . W1[64x32]
    ok &= run(64, 32, 128, 1, 1);   // wgrad: dZ^T[64 x 128] . X[128 x 32]
    ok &= run(64, 64, 128, 1, 1);   // wgrad hidden
    ok &= run(64, 16, 128, 1, 1);   // wgrad of the output layer, transposed: H^T[64 x 128] . dZo[128 x 16]
    ok &= run(128, 128, 64, 0, 0);  // 128-wide rgb net
    printf(ok ? "ALL OK\n" : "FAILURES\n");
    return ok ? 0 : 1;
}
