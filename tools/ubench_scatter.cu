// Microbenchmark (development tool, not part of the product): phase ablation of the tile scatter.
// Which phase of the kernel costs the time -- ranking, staging, the copy-out arithmetic, or the
// scattered global stores?   nvcc -arch=sm_100a -O3 -o ubench_scatter ubench_scatter.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int kItems = 16;

__device__ __forceinline__ uint32_t fmix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16; return h;
}

// MODE 0: load+hash+rank   1: +scan/reserve   2: +staging   3: +copy-out without global stores
//      4: full             5: full, but sequential destinations (dd = lo + pos)
//      6: full, 8-byte payload gathered from the tile window     7: like 4 with depth 4
template <int MODE, int CTAS, int STRIDE = 1, bool NOHASH = false, int kThreads = 256>
__global__ void __launch_bounds__(kThreads, CTAS) scatter_kernel(const uint32_t* __restrict__ keys, uint64_t n, int bits, uint32_t* __restrict__ cursor,
                                                              uint32_t* __restrict__ keys_out, uint32_t* __restrict__ idx_out,
                                                              const uint64_t* __restrict__ pay, uint64_t* __restrict__ pay_out, uint32_t* __restrict__ sink) {
    constexpr int kTile = kThreads * kItems, kWarps = kThreads / 32;
    extern __shared__ __align__(16) uint32_t s_dyn[];
    uint32_t* s_keys = s_dyn;
    uint32_t* s_idx = s_dyn + kTile;
    __shared__ uint32_t s_count[257], s_start[257], s_gbase[256], s_warp_sums[kWarps], s_total;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t nb = 1u << bits, mask = nb - 1;
    for (uint32_t i = tid; i <= nb; i += kThreads) s_count[i] = 0;
    __syncthreads();
    uint32_t acc = 0;
    const uint64_t n_tiles = n / kTile;
    for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const uint64_t lo = t * kTile;
        const uint32_t* tk = keys + lo;
        uint32_t key[kItems], pr[kItems];
#pragma unroll
        for (int k = 0; k < kItems; ++k) key[k] = tk[k * kThreads + tid];
#pragma unroll
        for (int k = 0; k < kItems; ++k) {
            const uint32_t d = (NOHASH ? key[k] : fmix32(key[k])) & mask;
            pr[k] = (d << 16) | atomicAdd(&s_count[d], 1u);
        }
        __syncthreads();
        if (MODE == 0) {
#pragma unroll
            for (int k = 0; k < kItems; ++k) acc += pr[k];
            if (tid <= nb) s_count[tid] = 0;
            __syncthreads();
            continue;
        }
        {
            const uint32_t c = tid < nb ? s_count[tid] : 0u;
            uint32_t inc = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
            if (lane == 31) s_warp_sums[warp] = inc;
            __syncthreads();
            uint32_t prefix = 0;
#pragma unroll
            for (uint32_t w = 0; w < kWarps; ++w) prefix += w < warp ? s_warp_sums[w] : 0u;
            const uint32_t start = prefix + inc - c;
            if (tid < nb) {
                s_start[tid] = start;
                uint32_t g = 0;
                if (c) g = atomicAdd(&cursor[tid * STRIDE], c);
                s_gbase[tid] = g - start;
                s_count[tid] = 0;
            }
            if (tid == kThreads - 1) { s_total = prefix + inc; s_start[nb] = prefix + inc; s_count[nb] = 0; }
        }
        __syncthreads();
        if (MODE == 1) {
#pragma unroll
            for (int k = 0; k < kItems; ++k) acc += pr[k] + s_start[pr[k] >> 16];
            __syncthreads();
            continue;
        }
#pragma unroll
        for (int k = 0; k < kItems; ++k) {
            const uint32_t pos = s_start[pr[k] >> 16] + (pr[k] & 0xffffu);
            s_keys[pos] = key[k];
            s_idx[pos]  = ((pr[k] >> 16) << 16) | (k * kThreads + tid);
        }
        __syncthreads();
        if (MODE == 2) { acc += s_keys[tid] + s_idx[tid]; __syncthreads(); continue; }
        const uint32_t total = s_total, lo32 = (uint32_t)lo;
        constexpr int kDepth = MODE == 7 ? 4 : 8;
        for (uint32_t base = 0; base < total; base += kDepth * kThreads) {
            uint32_t kk[kDepth], ww[kDepth], dd[kDepth], rr[kDepth];
#pragma unroll
            for (int j = 0; j < kDepth; ++j) {
                const uint32_t pos = base + j * kThreads + tid;
                kk[j] = s_keys[pos]; ww[j] = s_idx[pos];
                dd[j] = MODE == 5 ? lo32 + pos : s_gbase[(ww[j] >> 16) & 0xffu] + pos;
                rr[j] = lo32 + (ww[j] & (kTile - 1));
            }
            if (MODE == 3) {
#pragma unroll
                for (int j = 0; j < kDepth; ++j) acc += kk[j] ^ dd[j] ^ rr[j];
            } else {
#pragma unroll
                for (int j = 0; j < kDepth; ++j) { keys_out[dd[j]] = kk[j]; idx_out[dd[j]] = rr[j]; }
                if (MODE == 6) {
                    uint64_t v[kDepth];
#pragma unroll
                    for (int j = 0; j < kDepth; ++j) v[j] = pay[rr[j]];
#pragma unroll
                    for (int j = 0; j < kDepth; ++j) pay_out[dd[j]] = v[j];
                }
            }
        }
        __syncthreads();
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

__global__ void fill(uint32_t* k, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) k[i] = fmix32((uint32_t)i * 2654435761u + 17);
}
__global__ void fill_seq(uint32_t* k, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) k[i] = (uint32_t)i * 0x9E3779B1u;
}
__global__ void set_cursor(uint32_t* c, uint32_t nb, uint32_t cap, int stride) { if (threadIdx.x < nb) c[threadIdx.x * stride] = threadIdx.x * cap; }

struct Bufs { uint32_t *keys, *cursor, *ko, *io, *sink; uint64_t *pay, *po; };

template <int MODE, int CTAS, int STRIDE = 1, bool NOHASH = false, int kThreads = 256>
float run(const Bufs& b, uint64_t n, int bits, int sms) {
    const uint32_t nb = 1u << bits, cap = ((uint32_t)(n / nb + n / nb / 50 + 4096)) & ~31u;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int r = 0; r < 3; ++r) {
        cudaFuncSetAttribute(scatter_kernel<MODE, CTAS, STRIDE, NOHASH, kThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, kThreads * kItems * 8);
        set_cursor<<<1, 256>>>(b.cursor, nb, cap, STRIDE);
        cudaEventRecord(e0);
        scatter_kernel<MODE, CTAS, STRIDE, NOHASH, kThreads><<<sms * CTAS, kThreads, kThreads * kItems * 8>>>(b.keys, n, bits, b.cursor, b.ko, b.io, b.pay, b.po, b.sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;
    }
    return best;
}

int main(int argc, char** argv) {
    const uint64_t n = 1ull << 29;
    const int bits = argc > 1 ? atoi(argv[1]) : 8;
    Bufs b;
    const uint64_t slack = n + n / 40 + (1u << 21);
    cudaMalloc(&b.keys, n * 4); cudaMalloc(&b.cursor, 1024 * 64); cudaMalloc(&b.ko, slack * 4); cudaMalloc(&b.io, slack * 4);
    cudaMalloc(&b.sink, 4); cudaMalloc(&b.pay, n * 8); cudaMalloc(&b.po, slack * 8);
    fill<<<(unsigned)((n + 255) / 256), 256>>>(b.keys, n);
    cudaMemset(b.pay, 1, n * 8);
    cudaDeviceSynchronize();
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("n=%llu bits=%d\n", (unsigned long long)n, bits);
    printf("0 load+hash+rank              %.3f ms\n", run<0, 4>(b, n, bits, sms));
    printf("1 +scan/reserve               %.3f ms\n", run<1, 4>(b, n, bits, sms));
    printf("2 +staging                    %.3f ms\n", run<2, 4>(b, n, bits, sms));
    printf("3 +copy-out, no stores        %.3f ms\n", run<3, 4>(b, n, bits, sms));
    printf("4 full (keys+idx stores)      %.3f ms\n", run<4, 4>(b, n, bits, sms));
    printf("7 full, depth 4               %.3f ms\n", run<7, 4>(b, n, bits, sms));
    printf("5 full, sequential dest       %.3f ms\n", run<5, 4>(b, n, bits, sms));
    printf("6 full + 8B payload gather    %.3f ms\n", run<6, 4>(b, n, bits, sms));
    printf("1 +scan/reserve, stride 32B   %.3f ms\n", run<1, 4, 8>(b, n, bits, sms));
    printf("1 +scan/reserve, stride 128B  %.3f ms\n", run<1, 4, 32>(b, n, bits, sms));
    printf("1 +scan/reserve, stride 256B  %.3f ms\n", run<1, 4, 64>(b, n, bits, sms));
    printf("4 full, stride 128B           %.3f ms\n", run<4, 4, 32>(b, n, bits, sms));
    printf("6 full+payload, stride 128B   %.3f ms\n", run<6, 4, 32>(b, n, bits, sms));
    fill_seq<<<(unsigned)((n + 255) / 256), 256>>>(b.keys, n);
    cudaDeviceSynchronize();
    printf("4 full, aligned 64B runs      %.3f ms\n", run<4, 4, 1, true>(b, n, bits, sms));
    printf("6 +payload, aligned runs      %.3f ms\n", run<6, 4, 1, true>(b, n, bits, sms));
    fill<<<(unsigned)((n + 255) / 256), 256>>>(b.keys, n);
    cudaDeviceSynchronize();
    printf("4 full, tile 8192 (512 thr x2) %.3f ms\n", run<4, 2, 1, false, 512>(b, n, bits, sms));
    printf("4 full, tile 16384 (1024 x1)  %.3f ms\n", run<4, 1, 1, false, 1024>(b, n, bits, sms));
    printf("6 +payload, tile 8192         %.3f ms\n", run<6, 2, 1, false, 512>(b, n, bits, sms));
    printf("6 +payload, tile 16384        %.3f ms\n", run<6, 1, 1, false, 1024>(b, n, bits, sms));
    printf("4 full, 3 CTAs/SM             %.3f ms\n", run<4, 3>(b, n, bits, sms));
    printf("4 full, 2 CTAs/SM             %.3f ms\n", run<4, 2>(b, n, bits, sms));
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
