// Result validation on the device -- the sorted-multiset comparison of the reference's harness
// (tests/read_sql.cpp:1159-1222: decode both tables, sort the rows, compare pairwise) without the minutes
// of CPU it costs on IMDB-sized results.
//
// Rows are variable-width (VARCHAR cells), so they are not sorted themselves: every row gets a 64-bit
// hash of all its cells (NULL markers included), each table's (hash, row id) pairs are sorted by an LSD
// radix sort, and the rows at equal rank are then compared IN FULL -- every validity bit, every value bit
// pattern, every string byte.  "Equal" is therefore exact: a bijection between the tables' rows has been
// exhibited in which every pair of rows is identical.  (Two different rows with one hash could only turn
// an equal pair of tables into a reported mismatch, never the other way round.)
//
// Radix sort: 8 passes of 8 bits over (uint64 key, uint32 value).  A pass = per-block digit histograms
// (digit-major), the engine's exclusive scan, and a stable scatter: a block walks its 4096 elements in
// order, warp by warp, ranking equal digits inside a 32-element step with __match_any_sync and across
// steps / warps with shared-memory counters.
#include "rj_common.cuh"
#include "rj_internal.h"

namespace rj {
namespace {

constexpr int      kSortThreads = 256;
constexpr int      kSortWarps   = kSortThreads / 32;
constexpr uint32_t kSortChunk   = 4096; // elements per block
constexpr uint32_t kPerWarp     = kSortChunk / kSortWarps;

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}

// h[i] = mix(h[i] * K + cell hash): a fixed-width cell hashes to its value bits, a NULL to a marker
__global__ void __launch_bounds__(256)
    hash_fixed_cells_kernel(const void* __restrict__ values, const uint32_t* __restrict__ valid, uint64_t n, int width,
                            uint64_t* __restrict__ h) {
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        uint64_t c = 0x9e3779b97f4a7c15ull; // NULL
        if (valid == nullptr || test_bit(valid, i))
            c = mix64((width == 4 ? static_cast<uint64_t>(static_cast<const uint32_t*>(values)[i]) : static_cast<const uint64_t*>(values)[i]) + 1);
        h[i] = mix64(h[i] * 0x100000001b3ull + c);
    }
}

// the same for a string column whose per-row hashes are already computed (launch_varchar_hash: NULL rows = 0)
__global__ void __launch_bounds__(256)
    hash_combine_kernel(const uint64_t* __restrict__ cell, const uint32_t* __restrict__ valid, uint64_t n, uint64_t* __restrict__ h) {
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint64_t c = (valid == nullptr || test_bit(valid, i)) ? mix64(cell[i] + 2) : 0x9e3779b97f4a7c15ull;
        h[i] = mix64(h[i] * 0x100000001b3ull + c);
    }
}

__global__ void __launch_bounds__(256) iota_kernel(uint32_t* __restrict__ out, uint64_t n) {
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<uint64_t>(gridDim.x) * blockDim.x)
        out[i] = static_cast<uint32_t>(i);
}

// per-block digit counts, digit-major: counts[d * n_blocks + block]
__global__ void __launch_bounds__(kSortThreads)
    sort_hist_kernel(const uint64_t* __restrict__ keys, uint64_t n, int shift, uint32_t n_blocks, uint32_t* __restrict__ counts) {
    __shared__ uint32_t s_cnt[256];
    s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t lo = static_cast<uint64_t>(blockIdx.x) * kSortChunk;
    for (uint32_t i = threadIdx.x; i < kSortChunk && lo + i < n; i += kSortThreads) atomicAdd(&s_cnt[(keys[lo + i] >> shift) & 0xffu], 1u);
    __syncthreads();
    counts[static_cast<uint64_t>(threadIdx.x) * n_blocks + blockIdx.x] = s_cnt[threadIdx.x];
}

// stable scatter of one pass: element i of the block goes to base[digit][block] + (elements of that digit
// earlier in the block)
__global__ void __launch_bounds__(kSortThreads)
    sort_scatter_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint64_t n, int shift, uint32_t n_blocks,
                        const uint64_t* __restrict__ base, uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
    __shared__ uint32_t s_warp[kSortWarps][256]; // phase 1: digit counts per warp; phase 2: running offsets
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, lt = lanemask_lt();
    for (uint32_t i = tid; i < kSortWarps * 256; i += kSortThreads) (&s_warp[0][0])[i] = 0;
    __syncthreads();
    const uint64_t lo = static_cast<uint64_t>(blockIdx.x) * kSortChunk + warp * kPerWarp;
    // phase 1: counts per (warp, digit)
    for (uint32_t it = 0; it < kPerWarp / 32; ++it) {
        const uint64_t i = lo + it * 32 + lane;
        const bool     in = i < n;
        const uint32_t d = in ? static_cast<uint32_t>((keys[i] >> shift) & 0xffu) : 256u + lane; // absent lanes: unique pseudo-digits
        const uint32_t peers = __match_any_sync(RJ_FULL_MASK, d);
        if (in && (peers & lt) == 0) s_warp[warp][d] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // exclusive prefix over the warps, per digit; thread d owns digit d
    {
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
            const uint32_t c = s_warp[w][tid];
            s_warp[w][tid] = run;
            run += c;
        }
    }
    __syncthreads();
    // phase 2: the same walk, now placing
    for (uint32_t it = 0; it < kPerWarp / 32; ++it) {
        const uint64_t i = lo + it * 32 + lane;
        const bool     in = i < n;
        const uint64_t k = in ? keys[i] : 0ull;
        const uint32_t d = in ? static_cast<uint32_t>((k >> shift) & 0xffu) : 256u + lane;
        const uint32_t peers = __match_any_sync(RJ_FULL_MASK, d);
        if (in) {
            const uint32_t rank = s_warp[warp][d] + __popc(peers & lt);
            const uint64_t dst = base[static_cast<uint64_t>(d) * n_blocks + blockIdx.x] + rank;
            keys_out[dst] = k;
            vals_out[dst] = vals[i];
        }
        __syncwarp();
        if (in && (peers & lt) == 0) s_warp[warp][d] += __popc(peers);
        __syncwarp();
    }
}

// rows at equal rank must be identical in a fixed-width column: same validity, same value bits where valid
__global__ void __launch_bounds__(256)
    pairs_equal_fixed_kernel(const void* __restrict__ va, const uint32_t* __restrict__ valid_a, const uint32_t* __restrict__ idx_a,
                             const void* __restrict__ vb, const uint32_t* __restrict__ valid_b, const uint32_t* __restrict__ idx_b,
                             uint64_t n, int width, unsigned long long* __restrict__ mismatches) {
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint32_t ra = idx_a[i], rb = idx_b[i];
        const bool oa = valid_a == nullptr || test_bit(valid_a, ra), ob = valid_b == nullptr || test_bit(valid_b, rb);
        bool eq = oa == ob;
        if (eq && oa) {
            eq = width == 4 ? static_cast<const uint32_t*>(va)[ra] == static_cast<const uint32_t*>(vb)[rb]
                            : static_cast<const uint64_t*>(va)[ra] == static_cast<const uint64_t*>(vb)[rb];
        }
        if (!eq) atomicAdd(mismatches, 1ull);
    }
}

// string columns: validity here, bytes by launch_varchar_pairs_equal (keep[i] = strings equal)
__global__ void __launch_bounds__(256)
    pairs_equal_varchar_finish_kernel(const uint32_t* __restrict__ valid_a, const uint32_t* __restrict__ idx_a,
                                      const uint32_t* __restrict__ valid_b, const uint32_t* __restrict__ idx_b,
                                      const uint32_t* __restrict__ keep, uint64_t n, unsigned long long* __restrict__ mismatches) {
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const bool oa = valid_a == nullptr || test_bit(valid_a, idx_a[i]), ob = valid_b == nullptr || test_bit(valid_b, idx_b[i]);
        if (oa != ob || (oa && keep[i] == 0)) atomicAdd(mismatches, 1ull);
    }
}

__global__ void __launch_bounds__(256)
    keys_differ_kernel(const uint64_t* __restrict__ a, const uint64_t* __restrict__ b, uint64_t n, unsigned long long* __restrict__ mismatches) {
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<uint64_t>(gridDim.x) * blockDim.x)
        if (a[i] != b[i]) atomicAdd(mismatches, 1ull);
}

unsigned grid1d(uint64_t n, int sm_count) {
    const uint64_t want = (n + 255) / 256, cap = static_cast<uint64_t>(sm_count) * 8;
    return static_cast<unsigned>(want < 1 ? 1 : (want < cap ? want : cap));
}

} // namespace

void launch_hash_fixed_cells(const void* values, const uint32_t* valid, uint64_t n, int width, uint64_t* h, int sm_count, cudaStream_t s) {
    if (n == 0) return;
    hash_fixed_cells_kernel<<<grid1d(n, sm_count), 256, 0, s>>>(values, valid, n, width, h);
    RJ_LAUNCH_CHECK();
}

void launch_hash_combine(const uint64_t* cell, const uint32_t* valid, uint64_t n, uint64_t* h, int sm_count, cudaStream_t s) {
    if (n == 0) return;
    hash_combine_kernel<<<grid1d(n, sm_count), 256, 0, s>>>(cell, valid, n, h);
    RJ_LAUNCH_CHECK();
}

uint64_t sort_tmp_words(uint64_t n) { return ((n + kSortChunk - 1) / kSortChunk) * 256 + 256; }

// sorts (keys, vals) ascending by key, stably; the result is back in keys / vals (8 passes: an even number of
// ping-pongs).  alt_keys / alt_vals: scratch of n elements; counts [sort_tmp_words] u32; base [sort_tmp_words + 1] u64
void launch_radix_sort_u64(uint64_t* keys, uint32_t* vals, uint64_t* alt_keys, uint32_t* alt_vals, uint64_t n, uint32_t* counts,
                           uint64_t* base, void* scan_tmp, cudaStream_t s) {
    if (n < 2) return;
    const uint32_t n_blocks = static_cast<uint32_t>((n + kSortChunk - 1) / kSortChunk);
    const uint64_t entries = static_cast<uint64_t>(n_blocks) * 256;
    uint64_t* ka = keys; uint32_t* va = vals; uint64_t* kb = alt_keys; uint32_t* vb = alt_vals;
    for (int pass = 0; pass < 8; ++pass) {
        sort_hist_kernel<<<n_blocks, kSortThreads, 0, s>>>(ka, n, pass * 8, n_blocks, counts);
        RJ_LAUNCH_CHECK();
        launch_exclusive_scan_u32_u64(counts, base, entries, scan_tmp, s);
        sort_scatter_kernel<<<n_blocks, kSortThreads, 0, s>>>(ka, va, n, pass * 8, n_blocks, base, kb, vb);
        RJ_LAUNCH_CHECK();
        uint64_t* tk = ka; ka = kb; kb = tk;
        uint32_t* tv = va; va = vb; vb = tv;
    }
}

void launch_iota_u32(uint32_t* out, uint64_t n, int sm_count, cudaStream_t s) {
    if (n == 0) return;
    iota_kernel<<<grid1d(n, sm_count), 256, 0, s>>>(out, n);
    RJ_LAUNCH_CHECK();
}

void launch_pairs_equal_fixed(const void* va, const uint32_t* valid_a, const uint32_t* idx_a, const void* vb, const uint32_t* valid_b,
                              const uint32_t* idx_b, uint64_t n, int width, unsigned long long* mismatches, int sm_count, cudaStream_t s) {
    if (n == 0) return;
    pairs_equal_fixed_kernel<<<grid1d(n, sm_count), 256, 0, s>>>(va, valid_a, idx_a, vb, valid_b, idx_b, n, width, mismatches);
    RJ_LAUNCH_CHECK();
}

void launch_pairs_equal_varchar_finish(const uint32_t* valid_a, const uint32_t* idx_a, const uint32_t* valid_b, const uint32_t* idx_b,
                                       const uint32_t* keep, uint64_t n, unsigned long long* mismatches, int sm_count, cudaStream_t s) {
    if (n == 0) return;
    pairs_equal_varchar_finish_kernel<<<grid1d(n, sm_count), 256, 0, s>>>(valid_a, idx_a, valid_b, idx_b, keep, n, mismatches);
    RJ_LAUNCH_CHECK();
}

void launch_keys_differ(const uint64_t* a, const uint64_t* b, uint64_t n, unsigned long long* mismatches, int sm_count, cudaStream_t s) {
    if (n == 0) return;
    keys_differ_kernel<<<grid1d(n, sm_count), 256, 0, s>>>(a, b, n, mismatches);
    RJ_LAUNCH_CHECK();
}

} // namespace rj
