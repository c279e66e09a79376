// Microbenchmark (development tool, not part of the product; not yet run): random 8-byte probes into
// hash tables spread over the shared memory of a thread-block cluster (distributed shared memory).
// Question for the next round (profiles/r1_notes.md): can a 16-CTA cluster hold the 128 sub-partition
// tables of one pass-1 region, so that the probe side skips scatter pass 2?  That pays only if remote
// probes run at >= ~0.5 per SM and cycle.   nvcc -arch=sm_100a -O3 -o ubench_dsmem ubench_dsmem.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cg = cooperative_groups;

constexpr int      kThreads = 512;
constexpr uint32_t kSlots   = 16384; // 128 KB of 8-byte slots per CTA
constexpr int      kItems   = 8;     // independent probes in flight per thread

__device__ __forceinline__ uint32_t fmix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16; return h;
}

// every thread issues `rounds` batches of kItems probes; a probe picks a CTA of the cluster and a slot
__global__ void __launch_bounds__(kThreads, 1) probe_kernel(uint32_t rounds, uint32_t cluster_size, unsigned long long* sink) {
    extern __shared__ __align__(16) unsigned long long table[];
    cg::cluster_group cluster = cg::this_cluster();
    for (uint32_t s = threadIdx.x; s < kSlots; s += kThreads) table[s] = (static_cast<unsigned long long>(blockIdx.x) << 32) | s;
    cluster.sync();
    const uint32_t my = cluster.block_rank();
    unsigned long long acc = 0;
    uint32_t x = fmix32(blockIdx.x * kThreads + threadIdx.x + 1);
    for (uint32_t r = 0; r < rounds; ++r) {
        unsigned long long v[kItems];
#pragma unroll
        for (int k = 0; k < kItems; ++k) {
            x = fmix32(x + k + 1);
            const uint32_t cta = cluster_size > 1 ? (x >> 20) % cluster_size : my;
            const unsigned long long* remote = cluster.map_shared_rank(table, cta);
            v[k] = remote[x & (kSlots - 1)];
        }
#pragma unroll
        for (int k = 0; k < kItems; ++k) acc += v[k];
    }
    cluster.sync(); // nobody leaves while its shared memory may still be read
    if (acc == 0x1234567ull) sink[0] = acc;
}

static float run(int cluster_size, uint32_t rounds, int sms, unsigned long long* sink) {
    cudaLaunchConfig_t cfg = {};
    const int grid = (sms / cluster_size) * cluster_size;
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSlots * 8;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster_size;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSlots * 8));
    if (cluster_size > 8) cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaLaunchKernelEx(&cfg, probe_kernel, rounds, (uint32_t)cluster_size, sink); // warm-up
    cudaEventRecord(a);
    cudaLaunchKernelEx(&cfg, probe_kernel, rounds, (uint32_t)cluster_size, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double probes = double(grid) * kThreads * double(rounds) * kItems;
    printf("cluster %2d: grid %3d  %.3f ms  %.1f G probes/s  (%.2f per SM and cycle at 1.9 GHz)  %s\n", cluster_size, grid, ms,
           probes / ms / 1e6, probes / ms / 1e6 / grid / 1.9, cudaGetErrorString(cudaGetLastError()));
    return ms;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned long long* sink;
    cudaMalloc(&sink, 8);
    for (int cs: {1, 2, 4, 8, 16}) run(cs, 2048, sms, sink);
    return 0;
}
