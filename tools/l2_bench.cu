// L2 microbenchmark ("for T <= 2^20 report achieved L2 GB/s against a builder-measured random-32-B-sector L2
// peak"): the denominators the hash-grid gather (field_fwd) and the gradient scatter (grid_scatter_pair) are graded against when
// their tables are L2-resident (fp16 table 21.8 MiB, fp32 gradient table 43.6 MiB at T = 2^19; B200 L2 = 126 MB).
//   gather : every thread issues U independent 4-byte ld.global.nc at uniformly random addresses of a B-byte buffer (one 32-B sector
//            each; "pair" variant: the two lanes of a pair hit the two halves of one 8-byte slot, i.e. one sector per pair, the
//            access shape of gather_level_pair)
//   red    : every thread issues U red.global.add.v2.f32 at uniformly random 8-byte slots ("pair": lanes 2p, 2p+1 on adjacent
//            slots of one 32-B sector, the shape of grid_scatter_pair)
// Prints one JSON object; run by tools/gpu_session.sh, result committed as profiles/l2_peaks_r02.json.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo tools/l2_bench.cu -o tools/bin/l2_bench
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x) {   // lowbias32
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

template <int U, bool PAIR>
__global__ void __launch_bounds__(256) gather_kernel(const uint32_t* __restrict__ buf, uint32_t n_words, int iters, uint32_t* __restrict__ sink) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t key = PAIR ? (tid >> 1) : tid;
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
        uint32_t v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            uint32_t w = mix(key * 0x9e3779b9u + (uint32_t)(it * U + u) * 0x85ebca6bu) % n_words;
            if (PAIR) w = (w & ~1u) | (tid & 1u);
            asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(v[u]) : "l"(buf + w));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc ^= v[u];
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

template <int U, bool PAIR>
__global__ void __launch_bounds__(256) red_kernel(float2* __restrict__ buf, uint32_t n_slots, int iters) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t key = PAIR ? (tid >> 1) : tid;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            uint32_t w = mix(key * 0x9e3779b9u + (uint32_t)(it * U + u) * 0x85ebca6bu) % n_slots;
            if (PAIR) w = (w & ~1u) | (tid & 1u);
            asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(buf + w), "f"(1.0f), "f"(0.5f) : "memory");
        }
    }
}

template <typename F>
static float time_ms(F launch, int reps) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    launch(); launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(a)); launch(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    return best;
}

int main(int argc, char** argv) {
    int sms = 148;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const size_t sizes[3] = {(size_t)22 << 20, (size_t)44 << 20, (size_t)512 << 20};     // fp16 table, fp32 gradient table, HBM-resident control
    const char* names[3] = {"22MiB", "44MiB", "512MiB"};
    void* buf; uint32_t* sink;
    CK(cudaMalloc(&buf, sizes[2])); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(buf, 0, sizes[2]));
    const int iters = 64, reps = 5;
    printf("{\"sms\": %d", sms);
    for (int s = 0; s < 3; ++s) {
        const uint32_t n_words = (uint32_t)(sizes[s] / 4), n_slots = (uint32_t)(sizes[s] / 8);
        for (int cps = 4; cps <= 8; cps += 4) {
            const unsigned grid = sms * cps;
            const double threads = (double)grid * 256;
            {
                constexpr int U = 8;
                float ms = time_ms([&] { gather_kernel<U, false><<<grid, 256>>>((const uint32_t*)buf, n_words, iters, sink); }, reps);
                printf(",\n \"gather_%s_cta%d\": {\"Gsectors_per_s\": %.1f, \"GBps_32B_sectors\": %.0f}", names[s], cps, threads * iters * U / ms / 1e6,
                       threads * iters * U * 32.0 / ms / 1e6);
                ms = time_ms([&] { gather_kernel<U, true><<<grid, 256>>>((const uint32_t*)buf, n_words, iters, sink); }, reps);
                printf(",\n \"gather_pair_%s_cta%d\": {\"Gloads_per_s\": %.1f, \"Gsectors_per_s\": %.1f, \"GBps_32B_sectors\": %.0f}", names[s], cps,
                       threads * iters * U / ms / 1e6, threads * iters * U / 2 / ms / 1e6, threads * iters * U / 2 * 32.0 / ms / 1e6);
            }
            {
                constexpr int U = 8;
                float ms = time_ms([&] { red_kernel<U, false><<<grid, 256>>>((float2*)buf, n_slots, iters); }, reps);
                printf(",\n \"red_v2_%s_cta%d\": {\"Gred_per_s\": %.1f, \"GBps_8B_payload\": %.0f}", names[s], cps, threads * iters * U / ms / 1e6,
                       threads * iters * U * 8.0 / ms / 1e6);
                ms = time_ms([&] { red_kernel<U, true><<<grid, 256>>>((float2*)buf, n_slots, iters); }, reps);
                printf(",\n \"red_v2_pair_%s_cta%d\": {\"Gred_per_s\": %.1f, \"Gsectors_per_s\": %.1f}", names[s], cps, threads * iters * U / ms / 1e6,
                       threads * iters * U / 2 / ms / 1e6);
            }
        }
    }
    printf("\n}\n");
    return 0;
}
