// Device-wide prefix scans (reduce / scan-of-partials / scan-and-add), used for page row offsets,
// VARCHAR page layout and pair compaction.  They move a few bytes per element and are never the
// bottleneck of the join; the hot kernels are in k_partition.cu and k_join.cu.
#include "rj_common.cuh"
#include "rj_internal.h"

namespace rj {
namespace {

constexpr int kScanThreads = 256;
constexpr int kScanItems   = 8;
constexpr int kScanTile    = kScanThreads * kScanItems;

struct SumOp {
    __device__ __forceinline__ uint64_t operator()(uint64_t a, uint64_t b) const { return a + b; }
    static constexpr uint64_t identity = 0;
};
struct MaxOp {
    __device__ __forceinline__ uint64_t operator()(uint64_t a, uint64_t b) const { return a > b ? a : b; }
    static constexpr uint64_t identity = 0;
};

template <class Op>
__device__ __forceinline__ uint64_t warp_inclusive(uint64_t v, Op op) {
    const uint32_t lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint64_t o = __shfl_up_sync(RJ_FULL_MASK, v, d);
        if (lane >= d) v = op(v, o);
    }
    return v;
}

// inclusive scan of one value per thread across the block; returns the inclusive value and the
// block total through *total
template <class Op>
__device__ __forceinline__ uint64_t block_inclusive(uint64_t v, Op op, uint64_t* total) {
    __shared__ uint64_t warp_sums[kScanThreads / 32];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t inc = warp_inclusive(v, op);
    __syncthreads(); // protect warp_sums reuse across calls
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    uint64_t prefix = Op::identity, tot = Op::identity;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w) {
        uint64_t s = warp_sums[w];
        if (w < warp) prefix = op(prefix, s);
        tot = op(tot, s);
    }
    *total = tot;
    return op(prefix, inc);
}

template <class InT, class Op>
__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const InT* __restrict__ in, uint64_t n,
                                                                   uint64_t* __restrict__ partial) {
    Op op;
    uint64_t base = static_cast<uint64_t>(blockIdx.x) * kScanTile;
    uint64_t acc = Op::identity;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        uint64_t i = base + static_cast<uint64_t>(k) * kScanThreads + threadIdx.x;
        if (i < n) acc = op(acc, static_cast<uint64_t>(in[i]));
    }
    uint64_t total;
    block_inclusive(acc, op, &total);
    if (threadIdx.x == 0) partial[blockIdx.x] = total;
}

// single block: INCLUSIVE scan of the per-tile partials in place; partial[n_tiles] = grand total
template <class Op>
__global__ void __launch_bounds__(kScanThreads) scan_partials_kernel(uint64_t* partial, uint64_t n_tiles) {
    Op op;
    uint64_t carry = Op::identity;
    for (uint64_t base = 0; base < n_tiles; base += kScanThreads) {
        uint64_t i = base + threadIdx.x;
        uint64_t v = i < n_tiles ? partial[i] : Op::identity;
        uint64_t total;
        uint64_t inc = block_inclusive(v, op, &total);
        if (i < n_tiles) partial[i] = op(carry, inc);
        carry = op(carry, total);
    }
    if (threadIdx.x == 0) partial[n_tiles] = carry;
}

// The partials hold INCLUSIVE prefixes after scan_partials_kernel; tile b uses partial[b-1].
template <class InT, class Op, bool kExclusiveOut>
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const InT* __restrict__ in, uint64_t n,
                                                                  const uint64_t* __restrict__ partial,
                                                                  uint64_t* __restrict__ out) {
    Op op;
    uint64_t base = static_cast<uint64_t>(blockIdx.x) * kScanTile + static_cast<uint64_t>(threadIdx.x) * kScanItems;
    uint64_t v[kScanItems];
    uint64_t acc = Op::identity;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        uint64_t i = base + k;
        v[k] = i < n ? static_cast<uint64_t>(in[i]) : Op::identity;
        acc = op(acc, v[k]);
    }
    uint64_t total;
    uint64_t inc = block_inclusive(acc, op, &total);
    // prefix of everything before this thread's first item
    uint64_t tile_prefix = blockIdx.x == 0 ? Op::identity : partial[blockIdx.x - 1];
    // exclusive-of-thread = inclusive minus own contribution is not defined for max; recompute by shuffle
    __shared__ uint64_t thread_inc[kScanThreads];
    thread_inc[threadIdx.x] = inc;
    __syncthreads();
    uint64_t run = threadIdx.x == 0 ? tile_prefix : op(tile_prefix, thread_inc[threadIdx.x - 1]);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        uint64_t i = base + k;
        if (kExclusiveOut) {
            if (i < n) out[i] = run;
            run = op(run, v[k]);
        } else {
            run = op(run, v[k]);
            if (i < n) out[i] = run;
        }
    }
    if (kExclusiveOut && blockIdx.x == gridDim.x - 1) {
        // last element + 1 holds the grand total
        uint64_t last = n - 1;
        if (last >= base && last < base + kScanItems) out[n] = run;
    }
}

template <class InT, class Op, bool kExclusiveOut>
void run_scan(const InT* in, uint64_t* out, uint64_t n, void* tmp, cudaStream_t s) {
    if (n == 0) {
        if (kExclusiveOut) RJ_CUDA(cudaMemsetAsync(out, 0, sizeof(uint64_t), s));
        return;
    }
    uint64_t  n_tiles = (n + kScanTile - 1) / kScanTile;
    uint64_t* partial = static_cast<uint64_t*>(tmp);
    scan_reduce_kernel<InT, Op><<<static_cast<unsigned>(n_tiles), kScanThreads, 0, s>>>(in, n, partial);
    RJ_LAUNCH_CHECK();
    scan_partials_kernel<Op><<<1, kScanThreads, 0, s>>>(partial, n_tiles);
    RJ_LAUNCH_CHECK();
    scan_apply_kernel<InT, Op, kExclusiveOut><<<static_cast<unsigned>(n_tiles), kScanThreads, 0, s>>>(in, n, partial, out);
    RJ_LAUNCH_CHECK();
}

} // namespace

size_t scan_tmp_bytes(uint64_t n) { return ((n + kScanTile - 1) / kScanTile + 2) * sizeof(uint64_t); }

void launch_exclusive_scan_u32_u64(const uint32_t* in, uint64_t* out, uint64_t n, void* tmp, cudaStream_t s) {
    run_scan<uint32_t, SumOp, true>(in, out, n, tmp, s);
}

void launch_inclusive_sum_u64(const uint64_t* in, uint64_t* out, uint64_t n, void* tmp, cudaStream_t s) {
    run_scan<uint64_t, SumOp, false>(in, out, n, tmp, s);
}

void launch_inclusive_max_u64(const uint64_t* in, uint64_t* out, uint64_t n, void* tmp, cudaStream_t s) {
    run_scan<uint64_t, MaxOp, false>(in, out, n, tmp, s);
}

} // namespace rj
