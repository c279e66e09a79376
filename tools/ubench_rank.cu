// Microbenchmark (development tool, not part of the product): how fast can a CTA rank the tuples of a
// tile inside their radix bucket?  Compares the shared-memory atomic of the scatter kernel with
// match-based warp-private counters.   nvcc -arch=sm_100a -O3 -o ubench_rank ubench_rank.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int kThreads = 256, kItems = 16, kTile = kThreads * kItems, kWarps = kThreads / 32;

__device__ __forceinline__ uint32_t fmix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16; return h;
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, 4) rank_kernel(const uint32_t* __restrict__ keys, uint64_t n, int bits, uint32_t* __restrict__ out) {
    __shared__ uint32_t s_count[kWarps][257];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt = (1u << lane) - 1, mask = (1u << bits) - 1;
    for (uint32_t i = tid; i < kWarps * 257; i += kThreads) (&s_count[0][0])[i] = 0;
    __syncthreads();
    uint32_t acc = 0;
    const uint64_t n_tiles = n / kTile;
    for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const uint32_t* tk = keys + t * kTile;
        uint32_t d[kItems];
#pragma unroll
        for (int k = 0; k < kItems; ++k) d[k] = fmix32(tk[k * kThreads + tid]) & mask;
#pragma unroll
        for (int k = 0; k < kItems; ++k) {
            uint32_t rank = 0;
            if (MODE == 1) rank = atomicAdd(&s_count[0][d[k]], 1u);
            if (MODE == 5) atomicAdd(&s_count[0][d[k]], 1u);
            if (MODE == 2) {
                const uint32_t peers = __match_any_sync(0xffffffffu, d[k]);
                const uint32_t leader = __ffs(peers) - 1;
                uint32_t base = 0;
                if (lane == leader) base = atomicAdd(&s_count[0][d[k]], (uint32_t)__popc(peers));
                rank = __shfl_sync(0xffffffffu, base, leader) + __popc(peers & lt);
            }
            if (MODE == 3) {
                const uint32_t peers = __match_any_sync(0xffffffffu, d[k]);
                const uint32_t leader = __ffs(peers) - 1;
                uint32_t base = 0;
                if (lane == leader) { base = s_count[warp][d[k]]; s_count[warp][d[k]] = base + __popc(peers); }
                __syncwarp();
                rank = __shfl_sync(0xffffffffu, base, leader) + __popc(peers & lt);
            }
            if (MODE == 4) {
                uint32_t peers = 0xffffffffu;
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const uint32_t bal = __ballot_sync(0xffffffffu, (d[k] >> b) & 1u);
                    peers &= ((d[k] >> b) & 1u) ? bal : ~bal;
                }
                const uint32_t leader = __ffs(peers) - 1;
                uint32_t base = 0;
                if (lane == leader) { base = s_count[warp][d[k]]; s_count[warp][d[k]] = base + __popc(peers); }
                __syncwarp();
                rank = __shfl_sync(0xffffffffu, base, leader) + __popc(peers & lt);
            }
            if (MODE == 6) { // warp-private counters, atomics (no match)
                rank = atomicAdd(&s_count[warp][d[k]], 1u);
            }
            acc += rank ^ d[k];
        }
        __syncthreads();
        for (uint32_t i = tid; i < kWarps * 257; i += kThreads) (&s_count[0][0])[i] = 0;
        __syncthreads();
    }
    if (acc == 0x12345678u) out[0] = acc;
}

__global__ void fill(uint32_t* k, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) k[i] = fmix32((uint32_t)i * 2654435761u + 17);
}

template <int MODE>
float run(const uint32_t* keys, uint64_t n, int bits, uint32_t* out, int blocks) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    rank_kernel<MODE><<<blocks, kThreads>>>(keys, n, bits, out);
    cudaEventRecord(a);
    for (int r = 0; r < 3; ++r) rank_kernel<MODE><<<blocks, kThreads>>>(keys, n, bits, out);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / 3;
}

int main(int argc, char** argv) {
    const uint64_t n = 1ull << 29;
    const int bits = argc > 1 ? atoi(argv[1]) : 8;
    uint32_t *keys, *out;
    cudaMalloc(&keys, n * 4); cudaMalloc(&out, 4);
    fill<<<(unsigned)((n + 255) / 256), 256>>>(keys, n);
    cudaDeviceSynchronize();
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 4;
    printf("n=%llu bits=%d blocks=%d\n", (unsigned long long)n, bits, blocks);
    printf("0 load+hash only            %.3f ms\n", run<0>(keys, n, bits, out, blocks));
    printf("1 ATOMS per tuple           %.3f ms\n", run<1>(keys, n, bits, out, blocks));
    printf("5 RED (no return)           %.3f ms\n", run<5>(keys, n, bits, out, blocks));
    printf("6 ATOMS warp-private        %.3f ms\n", run<6>(keys, n, bits, out, blocks));
    printf("2 match + leader ATOMS      %.3f ms\n", run<2>(keys, n, bits, out, blocks));
    printf("3 match + private LDS/STS   %.3f ms\n", run<3>(keys, n, bits, out, blocks));
    printf("4 8 ballots + private       %.3f ms\n", run<4>(keys, n, bits, out, blocks));
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
