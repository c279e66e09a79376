// Late materialisation and page output.
//
//   gather        : out[i] = src[idx[i]] (+ validity bits) -- join keys and row-id lists of intermediates
//   encode_fixed  : the GPU replacement of Table::to_columnar for INT32/INT64/FP64
//                   (reference src/build_table.cpp:466-594).  The reference fills pages greedily
//                   (:488,:495); only the decoded multiset is observable (tests/read_sql.cpp:1206-1221),
//                   and under-filling a page is always legal, so every page takes a FIXED number of rows
//                   (1984 x 4 B or 1007 x 8 B: 4|8 + rows*width + ceil(rows/8) <= 8192) and pages become
//                   independent: one warp gathers the rows of a page through the row-id list, compacts the
//                   non-NULL values with ballot/popcount, assembles header + values + bitmap in shared
//                   memory and writes the page with ONE 8 KB TMA bulk store (cp.async.bulk, double
//                   buffered so the next page is assembled while the previous one drains).
#include "rj_common.cuh"
#include "rj_internal.h"

namespace rj {
namespace {

constexpr int kEncWarps = 4;
constexpr int kEncBufs  = 2;

template <typename T>
__global__ void __launch_bounds__(256)
    gather_kernel(const T* __restrict__ src, const uint32_t* __restrict__ src_valid, const uint32_t* __restrict__ idx,
                  uint64_t n, T* __restrict__ out, uint32_t* __restrict__ out_valid) {
    // each warp owns 32-row groups so that it can assemble whole validity words
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t gw = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nw = (static_cast<uint64_t>(gridDim.x) * blockDim.x) >> 5;
    const uint64_t n_groups = (n + 31) >> 5;
    for (uint64_t g = gw; g < n_groups; g += nw) {
        const uint64_t i  = (g << 5) + lane;
        const bool     in = i < n;
        uint32_t r = in ? idx[i] : 0u;
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

template <typename T>
__global__ void __launch_bounds__(kEncWarps * 32)
    encode_fixed_kernel(const T* __restrict__ values, const uint32_t* __restrict__ valid,
                        const uint8_t* __restrict__ valid_bytes, const uint32_t* __restrict__ idx,
                        const uint32_t* __restrict__ vidx, uint64_t n, uint8_t* __restrict__ pages_out) {
    constexpr uint32_t kRows  = rows_per_page(sizeof(T));
    constexpr uint32_t kBegin = sizeof(T) == 4 ? 4 : 8;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t lt = lanemask_lt();
    uint8_t* mybuf = smem + warp * kEncBufs * RJ_PAGE;
    const uint64_t n_pages = (n + kRows - 1) / kRows;
    const uint64_t gw = static_cast<uint64_t>(blockIdx.x) * kEncWarps + warp;
    const uint64_t nw = static_cast<uint64_t>(gridDim.x) * kEncWarps;
    uint32_t it = 0;
    for (uint64_t p = gw; p < n_pages; p += nw, ++it) {
        uint8_t* buf = mybuf + (it & 1) * RJ_PAGE;
        // the bulk store issued from this buffer two pages ago must have finished reading it
        if (lane == 0) tma_store_wait_read<kEncBufs - 1>();
        __syncwarp();
        // zero the page (header, value gap and bitmap tail must be deterministic)
        uint4* b4 = reinterpret_cast<uint4*>(buf);
#pragma unroll
        for (int k = 0; k < RJ_PAGE / 16 / 32; ++k) b4[k * 32 + lane] = make_uint4(0, 0, 0, 0);
        __syncwarp();
        const uint64_t j0  = p * kRows;
        const uint32_t cnt = (n - j0 < kRows) ? static_cast<uint32_t>(n - j0) : kRows;
        T*       vals = reinterpret_cast<T*>(buf + kBegin);
        uint8_t* bm   = buf + RJ_PAGE - ((cnt + 7) >> 3);
        uint32_t running = 0;
        // 4 groups of 32 rows per iteration: 4 independent row-id loads, then 4 independent gathers
        for (uint32_t base = 0; base < cnt; base += 128) {
            uint32_t r[4], vr[4];
            bool     in[4], ok[4];
            T        v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t i = base + u * 32 + lane;
                in[u] = i < cnt;
                r[u]  = in[u] ? (idx != nullptr ? idx[j0 + i] : static_cast<uint32_t>(j0 + i)) : 0u;
                vr[u] = r[u];
                if (valid != nullptr && vidx != idx && in[u]) vr[u] = vidx != nullptr ? vidx[j0 + i] : static_cast<uint32_t>(j0 + i);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                ok[u] = in[u];
                if (valid_bytes != nullptr) {
                    ok[u] = in[u] && valid_bytes[r[u]] != 0;
                } else if (valid != nullptr) {
                    ok[u] = in[u] && test_bit(valid, vr[u]);
                }
                v[u]  = ok[u] ? values[r[u]] : T(0);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t word = __ballot_sync(RJ_FULL_MASK, ok[u]);
                if (ok[u]) vals[running + __popc(word & lt)] = v[u];
                running += __popc(word);
                const uint32_t byte0 = (base + u * 32) >> 3; // first bitmap byte of this group
                if (lane < 4 && base + u * 32 + lane * 8 < cnt) bm[byte0 + lane] = static_cast<uint8_t>(word >> (8 * lane));
            }
        }
        if (lane == 0) *reinterpret_cast<uint32_t*>(buf) = cnt | (running << 16); // n_r @0, n_v @2
        // generic-proxy writes -> async proxy: fence by every writer, then one lane issues the store
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
            tma_store_1d(pages_out + p * RJ_PAGE, buf, RJ_PAGE);
            tma_store_commit();
        }
    }
    if (lane == 0) tma_store_wait_all<0>();
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

__global__ void fill_u32_kernel(uint32_t* p, uint32_t v, uint64_t n) {
    uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

} // namespace

void launch_gather(const void* src, const uint32_t* src_valid, const uint32_t* idx, uint64_t n, int elem_bytes,
                   void* out, uint32_t* out_valid, int sm_count, cudaStream_t s) {
    if (n == 0) return;
    uint64_t want = (n + 255) / 256;
    unsigned blocks = static_cast<unsigned>(want < static_cast<uint64_t>(sm_count) * 16 ? want : static_cast<uint64_t>(sm_count) * 16);
    if (elem_bytes == 4) {
        gather_kernel<uint32_t><<<blocks, 256, 0, s>>>(static_cast<const uint32_t*>(src), src_valid, idx, n,
                                                       static_cast<uint32_t*>(out), out_valid);
    } else {
        gather_kernel<uint64_t><<<blocks, 256, 0, s>>>(static_cast<const uint64_t*>(src), src_valid, idx, n,
                                                       static_cast<uint64_t*>(out), out_valid);
    }
    RJ_LAUNCH_CHECK();
}

void launch_encode_fixed(const void* values, const uint32_t* valid, const uint8_t* valid_bytes, const uint32_t* idx,
                         const uint32_t* vidx, uint64_t n, int type, void* pages_out, int sm_count, cudaStream_t s) {
    if (n == 0) return;
    const size_t smem = kEncWarps * kEncBufs * RJ_PAGE;
    const uint32_t rows = type == RJ_INT32 ? rows_per_page(4) : rows_per_page(8);
    uint64_t n_pages = (n + rows - 1) / rows;
    uint64_t want = (n_pages + kEncWarps - 1) / kEncWarps;
    unsigned blocks = static_cast<unsigned>(want < static_cast<uint64_t>(sm_count) * 3 ? want : static_cast<uint64_t>(sm_count) * 3);
    if (type == RJ_INT32) {
        static bool configured = false;
        if (!configured) {
            RJ_CUDA(cudaFuncSetAttribute(encode_fixed_kernel<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured = true;
        }
        encode_fixed_kernel<uint32_t><<<blocks, kEncWarps * 32, smem, s>>>(
            static_cast<const uint32_t*>(values), valid, valid_bytes, idx, vidx, n, static_cast<uint8_t*>(pages_out));
    } else {
        static bool configured = false;
        if (!configured) {
            RJ_CUDA(cudaFuncSetAttribute(encode_fixed_kernel<uint64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured = true;
        }
        encode_fixed_kernel<uint64_t><<<blocks, kEncWarps * 32, smem, s>>>(
            static_cast<const uint64_t*>(values), valid, valid_bytes, idx, vidx, n, static_cast<uint8_t*>(pages_out));
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

void launch_fill_u32(uint32_t* p, uint32_t v, uint64_t n, cudaStream_t s) {
    if (n == 0) return;
    fill_u32_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(p, v, n);
    RJ_LAUNCH_CHECK();
}

} // namespace rj
