// Microbenchmark (development tool, not part of the product): random 8-byte gathers from a large
// array, the access pattern of late materialisation.  Compares load flavours and the
// cudaLimitMaxL2FetchGranularity hint.   nvcc -arch=sm_100a -O3 -o ubench_gather ubench_gather.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint64_t ld_plain(const uint64_t* p) { return *p; }
__device__ __forceinline__ uint64_t ld_nc(const uint64_t* p) {
    uint64_t v; asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p)); return v;
}
__device__ __forceinline__ uint64_t ld_l2_64(const uint64_t* p) {
    uint64_t v; asm volatile("ld.global.nc.L1::no_allocate.L2::64B.u64 %0, [%1];" : "=l"(v) : "l"(p)); return v;
}
__device__ __forceinline__ uint64_t ld_cv(const uint64_t* p) {
    uint64_t v; asm volatile("ld.global.cv.u64 %0, [%1];" : "=l"(v) : "l"(p)); return v;
}
__device__ __forceinline__ uint64_t ld_cg(const uint64_t* p) {
    uint64_t v; asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p)); return v;
}

template <int MODE>
__global__ void gather(const uint64_t* __restrict__ src, const uint32_t* __restrict__ idx, uint64_t n, uint64_t* __restrict__ out) {
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i + 3 * stride < n; i += 4 * stride) {
        uint32_t r0 = idx[i], r1 = idx[i + stride], r2 = idx[i + 2 * stride], r3 = idx[i + 3 * stride];
        uint64_t v0, v1, v2, v3;
        if (MODE == 0) { v0 = ld_plain(src + r0); v1 = ld_plain(src + r1); v2 = ld_plain(src + r2); v3 = ld_plain(src + r3); }
        if (MODE == 1) { v0 = ld_nc(src + r0); v1 = ld_nc(src + r1); v2 = ld_nc(src + r2); v3 = ld_nc(src + r3); }
        if (MODE == 2) { v0 = ld_l2_64(src + r0); v1 = ld_l2_64(src + r1); v2 = ld_l2_64(src + r2); v3 = ld_l2_64(src + r3); }
        if (MODE == 3) { v0 = ld_cv(src + r0); v1 = ld_cv(src + r1); v2 = ld_cv(src + r2); v3 = ld_cv(src + r3); }
        if (MODE == 4) { v0 = ld_cg(src + r0); v1 = ld_cg(src + r1); v2 = ld_cg(src + r2); v3 = ld_cg(src + r3); }
        out[i] = v0; out[i + stride] = v1; out[i + 2 * stride] = v2; out[i + 3 * stride] = v3;
    }
}

__global__ void fill_idx(uint32_t* idx, uint64_t n, uint32_t range) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        uint64_t z = i * 0x9E3779B97F4A7C15ull + 12345;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
        idx[i] = (uint32_t)(z % range);
    }
}

template <int MODE>
float run(const uint64_t* src, const uint32_t* idx, uint64_t n, uint64_t* out, int blocks) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    gather<MODE><<<blocks, 256>>>(src, idx, n, out);
    cudaEventRecord(a);
    for (int r = 0; r < 3; ++r) gather<MODE><<<blocks, 256>>>(src, idx, n, out);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / 3;
}

int main(int argc, char** argv) {
    size_t gran = argc > 1 ? atoi(argv[1]) : 0;
    if (gran) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran); printf("set limit %zu -> %s\n", gran, cudaGetErrorString(e)); }
    size_t got = 0; cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity); printf("L2 fetch granularity limit = %zu\n", got);
    const uint64_t n = 1ull << 27, range = 1u << 29; // 128 Mi gathers from a 4 GB array
    uint64_t *src, *out; uint32_t* idx;
    cudaMalloc(&src, range * 8); cudaMalloc(&out, n * 8); cudaMalloc(&idx, n * 4);
    cudaMemset(src, 1, range * 8);
    fill_idx<<<(unsigned)((n + 255) / 256), 256>>>(idx, n, (uint32_t)range);
    cudaDeviceSynchronize();
    const char* names[] = {"plain", "nc.no_allocate", "nc.L2::64B", "cv", "cg"};
    for (int blocks : {148 * 8, 148 * 32}) {
        float t[5] = {run<0>(src, idx, n, out, blocks), run<1>(src, idx, n, out, blocks), run<2>(src, idx, n, out, blocks),
                      run<3>(src, idx, n, out, blocks), run<4>(src, idx, n, out, blocks)};
        for (int m = 0; m < 5; ++m) printf("blocks=%5d %-16s %7.3f ms  %6.1f Grows/s\n", blocks, names[m], t[m], n / t[m] / 1e6);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
