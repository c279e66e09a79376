// Late materialisation and page output.
//
//   gather        : out[i] = src[idx[i]] (+ validity bits) -- join keys and row-id lists of intermediates
//   encode_fixed  : the GPU replacement of Table::to_columnar for INT32/INT64/FP64
//                   (reference src/build_table.cpp:466-594).  The reference fills pages greedily
//                   (:488,:495); only the decoded multiset is observable (tests/read_sql.cpp:1206-1221),
//                   and under-filling a page is always legal, so every page takes a FIXED number of rows
//                   (1984 x 4 B or 1007 x 8 B: 4|8 + rows*width + ceil(rows/8) <= 8192) and pages become
//                   independent: one CTA gathers the rows of a page through the row-id list, compacts the
//                   non-NULL values with a block-wide prefix sum, assembles header + values + bitmap in
//                   shared memory and writes the page with ONE 8 KB TMA bulk store (cp.async.bulk, double
//                   buffered so the next page is assembled while the previous one drains).
#include "rj_common.cuh"
#include "rj_internal.h"

namespace rj {
namespace {

constexpr int kEncThreads = 256;

template <typename T>
__global__ void __launch_bounds__(256)
    gather_kernel(const T* __restrict__ src, const uint32_t* __restrict__ src_valid, const uint32_t* __restrict__ idx,
                  uint64_t n, T* __restrict__ out, uint32_t* __restrict__ out_valid, uint32_t idx_mask) {
    // each warp owns 32-row groups so that it can assemble whole validity words
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t gw = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nw = (static_cast<uint64_t>(gridDim.x) * blockDim.x) >> 5;
    const uint64_t n_groups = (n + 31) >> 5;
    for (uint64_t g = gw; g < n_groups; g += nw) {
        const uint64_t i  = (g << 5) + lane;
        const bool     in = i < n;
        uint32_t r = in ? (idx[i] & idx_mask) : 0u;
        bool     v = in;
        if (in && src_valid != nullptr) v = test_bit(src_valid, r);
        if (in) out[i] = v ? src[r] : T(0);
        if (out_valid != nullptr) {
            const uint32_t word = __ballot_sync(RJ_FULL_MASK, v);
            if (lane == 0) out_valid[g] = word;
        }
    }
}

__host__ __device__ constexpr uint32_t rows_per_page(int width) { return width == 4 ? 1984u : 1007u; }

// One CTA of 256 threads per output page.  Thread t owns the kPer consecutive rows
// [t*kPer, (t+1)*kPer) of the page (8 rows of a 4-byte column = exactly one bitmap byte, 4 rows of an
// 8-byte column = one nibble), so ALL row-id loads of the page are issued at once, then all value /
// validity gathers: 1000-2000 independent gathers in flight per CTA and up to 8 CTAs per SM, instead
// of 128 per warp with 12 warps per SM in the first (warp-per-page) version, which ncu showed to be
// latency bound (18 % active warps, 12 % issue slots, DRAM at 19 %).
template <typename T>
__global__ void __launch_bounds__(kEncThreads)
    encode_fixed_kernel(const T* __restrict__ values, const uint32_t* __restrict__ valid,
                        const uint8_t* __restrict__ valid_bytes, const uint32_t* __restrict__ idx,
                        const uint32_t* __restrict__ vidx, uint64_t n, uint8_t* __restrict__ pages_out,
                        uint32_t idx_mask, int valid_bit) {
    constexpr uint32_t kRows  = rows_per_page(sizeof(T));
    constexpr uint32_t kBegin = sizeof(T) == 4 ? 4 : 8;
    constexpr int      kPer   = (kRows + kEncThreads - 1) / kEncThreads; // 8 or 4
    __shared__ __align__(128) uint8_t buf[RJ_PAGE];
    __shared__ uint32_t warp_sums[kEncThreads / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // ONE page per CTA, grid = pages: the hardware hands CTAs out in page order, so the pages in flight
    // are always a contiguous window of the output (persistent CTAs with a static stride drift apart,
    // and the gathers of a wide window no longer fit L2 -- measured 86 B of DRAM reads per row).
    {
        const uint64_t p = blockIdx.x;
        uint4* b4 = reinterpret_cast<uint4*>(buf);
#pragma unroll
        for (int k = 0; k < RJ_PAGE / 16 / kEncThreads; ++k) b4[k * kEncThreads + tid] = make_uint4(0, 0, 0, 0);
        const uint64_t j0  = p * kRows;
        const uint32_t cnt = (n - j0 < kRows) ? static_cast<uint32_t>(n - j0) : kRows;
        const uint32_t i0  = tid * kPer;
        uint32_t r[kPer], vr[kPer];
        bool     ok[kPer];
        T        v[kPer];
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            const bool in = i0 + k < cnt;
            const uint32_t raw = in ? (idx != nullptr ? idx[j0 + i0 + k] : static_cast<uint32_t>(j0 + i0 + k)) : 0u;
            r[k]  = raw & idx_mask;
            vr[k] = valid_bit >= 0 ? ((raw >> valid_bit) & 1u) : r[k]; // flag bit, or the index of the validity lookup
            if (valid_bit < 0 && valid != nullptr && valid_bytes == nullptr && vidx != idx && in)
                vr[k] = vidx != nullptr ? vidx[j0 + i0 + k] : static_cast<uint32_t>(j0 + i0 + k);
        }
        uint32_t mine = 0, bits = 0;
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            ok[k] = i0 + k < cnt;
            if (valid_bit >= 0) {
                ok[k] = ok[k] && vr[k] != 0;
            } else if (valid_bytes != nullptr) {
                ok[k] = ok[k] && valid_bytes[r[k]] != 0;
            } else if (valid != nullptr) {
                ok[k] = ok[k] && test_bit(valid, vr[k]);
            }
            v[k] = ok[k] ? values[r[k]] : T(0);
            mine += ok[k] ? 1u : 0u;
            bits |= (ok[k] ? 1u : 0u) << k;
        }
        // exclusive scan of the per-thread non-NULL counts across the CTA
        uint32_t inc = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(RJ_FULL_MASK, inc, d);
            if (lane >= d) inc += o;
        }
        if (lane == 31) warp_sums[warp] = inc;
        __syncthreads(); // also: the zero fill above is complete
        uint32_t base = inc - mine, total = 0;
#pragma unroll
        for (uint32_t w = 0; w < kEncThreads / 32; ++w) {
            const uint32_t sw = warp_sums[w];
            if (w < warp) base += sw;
            total += sw;
        }
        T* vals = reinterpret_cast<T*>(buf + kBegin);
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            if (ok[k]) vals[base++] = v[k];
        }
        uint8_t* bm = buf + RJ_PAGE - ((cnt + 7) >> 3);
        if (kPer == 8) {
            if (i0 < cnt) bm[tid] = static_cast<uint8_t>(bits);
        } else {
            // two threads share a bitmap byte
            const uint32_t hi = __shfl_down_sync(RJ_FULL_MASK, bits, 1);
            if ((tid & 1) == 0 && i0 < cnt) bm[tid >> 1] = static_cast<uint8_t>(bits | (hi << 4));
        }
        if (tid == 0) *reinterpret_cast<uint32_t*>(buf) = cnt | (total << 16); // n_r @0, n_v @2
        // generic-proxy writes -> async proxy: fence by every writer, then one thread issues the store
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
            tma_store_1d(pages_out + p * RJ_PAGE, buf, RJ_PAGE);
            tma_store_commit();
        }
    }
    if (tid == 0) tma_store_wait_read<0>(); // the page must be read out of shared memory before the CTA retires
}

__global__ void __launch_bounds__(256) bitmap_to_bytes_kernel(const uint32_t* __restrict__ bits, uint64_t n, uint8_t* __restrict__ out) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = test_bit(bits, i) ? 1 : 0;
}

__global__ void __launch_bounds__(256) bytes_to_bitmap_kernel(const uint8_t* __restrict__ bytes, uint64_t n, uint32_t* __restrict__ out) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t gw = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nw = (static_cast<uint64_t>(gridDim.x) * blockDim.x) >> 5;
    const uint64_t n_groups = (n + 31) >> 5;
    for (uint64_t g = gw; g < n_groups; g += nw) {
        const uint64_t i = (g << 5) + lane;
        const uint32_t word = __ballot_sync(RJ_FULL_MASK, i < n && bytes[i] != 0);
        if (lane == 0) out[g] = word;
    }
}

__global__ void __launch_bounds__(256) count_nulls_kernel(const uint32_t* __restrict__ valid, const uint32_t* __restrict__ idx, uint64_t n,
                                                          uint32_t idx_mask, unsigned long long* __restrict__ out) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    uint32_t mine = 0;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t r = idx != nullptr ? (idx[i] & idx_mask) : i;
        mine += test_bit(valid, r) ? 0u : 1u;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mine += __shfl_xor_sync(RJ_FULL_MASK, mine, d);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(out, static_cast<unsigned long long>(mine));
}

__global__ void fill_u32_kernel(uint32_t* p, uint32_t v, uint64_t n) {
    uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

} // namespace

void launch_gather(const void* src, const uint32_t* src_valid, const uint32_t* idx, uint64_t n, int elem_bytes,
                   void* out, uint32_t* out_valid, int sm_count, cudaStream_t s, uint32_t idx_mask) {
    if (n == 0) return;
    uint64_t want = (n + 255) / 256;
    unsigned blocks = static_cast<unsigned>(want < static_cast<uint64_t>(sm_count) * 16 ? want : static_cast<uint64_t>(sm_count) * 16);
    if (elem_bytes == 4) {
        gather_kernel<uint32_t><<<blocks, 256, 0, s>>>(static_cast<const uint32_t*>(src), src_valid, idx, n,
                                                       static_cast<uint32_t*>(out), out_valid, idx_mask);
    } else {
        gather_kernel<uint64_t><<<blocks, 256, 0, s>>>(static_cast<const uint64_t*>(src), src_valid, idx, n,
                                                       static_cast<uint64_t*>(out), out_valid, idx_mask);
    }
    RJ_LAUNCH_CHECK();
}

void launch_encode_fixed(const void* values, const uint32_t* valid, const uint8_t* valid_bytes, const uint32_t* idx,
                         const uint32_t* vidx, uint64_t n, int type, void* pages_out, int sm_count, cudaStream_t s,
                         uint32_t idx_mask, int valid_bit) {
    if (n == 0) return;
    const uint32_t rows = type == RJ_INT32 ? rows_per_page(4) : rows_per_page(8);
    const uint64_t n_pages = (n + rows - 1) / rows;
    (void)sm_count;
    if (n_pages > 0x7fffffffull) throw CudaError("encode: too many pages for one launch");
    const unsigned blocks = static_cast<unsigned>(n_pages);
    if (type == RJ_INT32) {
        encode_fixed_kernel<uint32_t><<<blocks, kEncThreads, 0, s>>>(
            static_cast<const uint32_t*>(values), valid, valid_bytes, idx, vidx, n, static_cast<uint8_t*>(pages_out), idx_mask, valid_bit);
    } else {
        encode_fixed_kernel<uint64_t><<<blocks, kEncThreads, 0, s>>>(
            static_cast<const uint64_t*>(values), valid, valid_bytes, idx, vidx, n, static_cast<uint8_t*>(pages_out), idx_mask, valid_bit);
    }
    RJ_LAUNCH_CHECK();
}

void launch_bitmap_to_bytes(const uint32_t* bits, uint64_t n, uint8_t* out, int sm_count, cudaStream_t s) {
    if (n == 0) return;
    uint64_t want = (n + 255) / 256;
    unsigned blocks = static_cast<unsigned>(want < static_cast<uint64_t>(sm_count) * 16 ? want : static_cast<uint64_t>(sm_count) * 16);
    bitmap_to_bytes_kernel<<<blocks, 256, 0, s>>>(bits, n, out);
    RJ_LAUNCH_CHECK();
}

void launch_bytes_to_bitmap(const uint8_t* bytes, uint64_t n, uint32_t* out, int sm_count, cudaStream_t s) {
    if (n == 0) return;
    uint64_t want = (n + 255) / 256;
    unsigned blocks = static_cast<unsigned>(want < static_cast<uint64_t>(sm_count) * 16 ? want : static_cast<uint64_t>(sm_count) * 16);
    bytes_to_bitmap_kernel<<<blocks, 256, 0, s>>>(bytes, n, out);
    RJ_LAUNCH_CHECK();
}

void launch_count_nulls(const uint32_t* valid, const uint32_t* idx, uint64_t n, uint32_t idx_mask, unsigned long long* out,
                        int sm_count, cudaStream_t s) {
    if (n == 0 || valid == nullptr) return;
    uint64_t want = (n + 255) / 256;
    unsigned blocks = static_cast<unsigned>(want < static_cast<uint64_t>(sm_count) * 8 ? want : static_cast<uint64_t>(sm_count) * 8);
    count_nulls_kernel<<<blocks, 256, 0, s>>>(valid, idx, n, idx_mask, out);
    RJ_LAUNCH_CHECK();
}

void launch_fill_u32(uint32_t* p, uint32_t v, uint64_t n, cudaStream_t s) {
    if (n == 0) return;
    fill_u32_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(p, v, n);
    RJ_LAUNCH_CHECK();
}

} // namespace rj
