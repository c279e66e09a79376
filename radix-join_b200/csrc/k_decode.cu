// Page ingest -- the GPU replacement of Table::from_columnar (reference src/build_table.cpp:312-436).
//
// Page layout (include/plan.h:54-58, build_table.cpp:322-430): u16 n_r @0, u16 n_v @2; the NULL
// bitmap is the LAST (n_r+7)/8 bytes of the 8 KB page, bit i = byte i/8, LSB first; INT32 values
// start @4, INT64/FP64 @8; VARCHAR: n_v u16 END offsets @4, characters @4+2*n_v; n_r == 0xffff is the
// first page of a long string (n_v characters @4), n_r == 0xfffe a continuation.
//
// Fixed-width decode is a warp-per-page kernel: each warp owns a 2-deep ring of 8 KB shared-memory
// buffers that one lane fills with 1-D TMA bulk loads (cp.async.bulk, completion on an mbarrier), so a
// whole page arrives as one asynchronous 8 KB transaction while the warp unpacks the previous page:
// header -> 32 rows at a time: bitmap bit, ballot, popcount prefix = index of the row's value among
// the page's non-NULL values -> dense value + validity word.  Output stores are fully coalesced.
#include "rj_common.cuh"
#include "rj_internal.h"

namespace rj {
namespace {

constexpr int kDecWarps  = 4;
constexpr int kDecStages = 2;

__global__ void __launch_bounds__(256) page_rows_kernel(const uint8_t* __restrict__ pages, uint64_t n_pages,
                                                        int type, uint32_t* __restrict__ rows,
                                                        uint64_t* __restrict__ totals) {
    uint64_t p = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    uint32_t r = 0, v = 0;
    if (p < n_pages) {
        uint32_t hdr = *reinterpret_cast<const uint32_t*>(pages + p * RJ_PAGE);
        uint32_t n_r = hdr & 0xffffu, n_v = hdr >> 16;
        if (type == RJ_VARCHAR && n_r >= 0xfffeu) {
            r = v = (n_r == 0xffffu) ? 1u : 0u; // build_table.cpp:384-405
        } else {
            r = n_r;
            v = n_v;
        }
        rows[p] = r;
    }
    if (totals) {
        // block reduction, then one atomic per block
        __shared__ unsigned long long s_r, s_v;
        if (threadIdx.x == 0) s_r = s_v = 0;
        __syncthreads();
        uint32_t wr = r, wv = v;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            wr += __shfl_xor_sync(RJ_FULL_MASK, wr, d);
            wv += __shfl_xor_sync(RJ_FULL_MASK, wv, d);
        }
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(&s_r, static_cast<unsigned long long>(wr));
            atomicAdd(&s_v, static_cast<unsigned long long>(wv));
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            atomicAdd(reinterpret_cast<unsigned long long*>(totals), s_r);
            atomicAdd(reinterpret_cast<unsigned long long*>(totals) + 1, s_v);
        }
    }
}

// OR a 32-row validity word that starts at global row g0 into the output bitmap
__device__ __forceinline__ void or_valid_word(uint32_t* __restrict__ out_valid, uint64_t g0, uint32_t word) {
    uint32_t sh = static_cast<uint32_t>(g0 & 31);
    uint64_t wi = g0 >> 5;
    atomicOr(&out_valid[wi], word << sh);
    if (sh) {
        uint32_t hi = word >> (32 - sh);
        if (hi) atomicOr(&out_valid[wi + 1], hi);
    }
}

template <typename T>
__device__ __forceinline__ void decode_fixed_page(const uint8_t* pg, uint64_t r0, T* __restrict__ out,
                                                  uint32_t* __restrict__ out_valid, uint32_t lane) {
    constexpr uint32_t kBegin   = sizeof(T) == 4 ? 4 : 8;
    constexpr uint32_t kMaxVals = (RJ_PAGE - kBegin) / sizeof(T);
    const uint32_t hdr = *reinterpret_cast<const uint32_t*>(pg);
    const uint32_t n_r = hdr & 0xffffu, n_v = hdr >> 16;
    const uint8_t* bm   = pg + RJ_PAGE - ((n_r + 7) >> 3);
    const T*       vals = reinterpret_cast<const T*>(pg + kBegin);
    const bool     dense = (n_v == n_r); // no NULL on this page: skip the bitmap
    const uint32_t lt = lanemask_lt();
    if (dense) {
        // no NULL on this page: a straight copy, 4 x 32 rows per step (4 independent LDS -> STG)
        const uint32_t n = n_r < kMaxVals ? n_r : kMaxVals;
        for (uint32_t base = 0; base < n; base += 128) {
            T v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t i = base + u * 32 + lane;
                v[u] = i < n ? vals[i] : T(0);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t i = base + u * 32 + lane;
                if (i < n) out[r0 + i] = v[u];
            }
        }
        if (out_valid != nullptr) {
            for (uint32_t base = lane * 32; base < n; base += 32 * 32) {
                const uint32_t left = n - base;
                or_valid_word(out_valid, r0 + base, left >= 32 ? 0xffffffffu : ((1u << left) - 1u));
            }
        }
        return;
    }
    uint32_t running = 0;
    for (uint32_t base = 0; base < n_r; base += 32) {
        const uint32_t i  = base + lane;
        const bool     in = i < n_r;
        uint32_t bit = in ? 1u : 0u;
        if (!dense && in) bit = (bm[i >> 3] >> (i & 7)) & 1u;
        const uint32_t word = __ballot_sync(RJ_FULL_MASK, bit);
        uint32_t rank = running + __popc(word & lt);
        rank = rank < kMaxVals ? rank : kMaxVals - 1; // malformed page: stay inside the buffer
        T v = bit ? vals[rank] : T(0);
        if (in) out[r0 + i] = v;
        running += __popc(word);
        if (out_valid != nullptr && lane == 0 && word) or_valid_word(out_valid, r0 + base, word);
    }
}

template <typename T>
__global__ void __launch_bounds__(kDecWarps * 32)
    decode_fixed_kernel(const uint8_t* __restrict__ pages, uint64_t n_pages, const uint64_t* __restrict__ row_start,
                        T* __restrict__ out, uint32_t* __restrict__ out_valid) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kDecWarps * kDecStages * RJ_PAGE);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t*  buf = smem + warp * kDecStages * RJ_PAGE;
    uint64_t* bar = bars + warp * kDecStages;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kDecStages; ++s) mbar_init(&bar[s], 1);
        fence_mbar_init();
    }
    __syncwarp();
    const uint64_t gw = static_cast<uint64_t>(blockIdx.x) * kDecWarps + warp;
    const uint64_t nw = static_cast<uint64_t>(gridDim.x) * kDecWarps;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kDecStages; ++s) {
            uint64_t p = gw + s * nw;
            if (p < n_pages) {
                mbar_arrive_expect_tx(&bar[s], RJ_PAGE);
                tma_load_1d(buf + s * RJ_PAGE, pages + p * RJ_PAGE, RJ_PAGE, &bar[s]);
            }
        }
    }
    uint32_t it = 0;
    for (uint64_t p = gw; p < n_pages; p += nw, ++it) {
        const uint32_t s = it % kDecStages, parity = (it / kDecStages) & 1u;
        mbar_wait(&bar[s], parity);
        decode_fixed_page<T>(buf + s * RJ_PAGE, row_start[p], out, out_valid, lane);
        __syncwarp(); // every lane is done reading the buffer before it is refilled
        const uint64_t pn = p + kDecStages * nw;
        if (lane == 0 && pn < n_pages) {
            mbar_arrive_expect_tx(&bar[s], RJ_PAGE);
            tma_load_1d(buf + s * RJ_PAGE, pages + pn * RJ_PAGE, RJ_PAGE, &bar[s]);
        }
    }
}

// VARCHAR: only the header, the offsets and the bitmap are read (a few % of the page); characters
// stay in the page buffer and are addressed through the descriptor (late materialisation).
__global__ void __launch_bounds__(128)
    decode_varchar_kernel(const uint8_t* __restrict__ pages, uint64_t n_pages, const uint64_t* __restrict__ row_start,
                          uint64_t* __restrict__ desc, uint32_t* __restrict__ out_valid, uint32_t* __restrict__ err_flags) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t gw = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nw = (static_cast<uint64_t>(gridDim.x) * blockDim.x) >> 5;
    const uint32_t lt = lanemask_lt();
    for (uint64_t p = gw; p < n_pages; p += nw) {
        const uint8_t* pg  = pages + p * RJ_PAGE;
        const uint32_t hdr = *reinterpret_cast<const uint32_t*>(pg);
        const uint32_t n_r = hdr & 0xffffu, n_v = hdr >> 16;
        const uint64_t r0  = row_start[p];
        if (n_r == 0xfffeu) {
            // continuation: its chars belong to the row BEFORE it (build_table.cpp:392-405).  After a
            // 0xffff / 0xfffe page that is the long string being assembled; after a regular page the
            // reference appends to that page's last row if it is a string (handled below, where that
            // page is decoded) and throws "long string page 0xfffe must follows a string" if it is NULL.
            if (lane == 0 && err_flags != nullptr) {
                bool ok = p > 0;
                if (ok) {
                    const uint8_t* pv  = pg - RJ_PAGE;
                    const uint32_t ph  = *reinterpret_cast<const uint32_t*>(pv);
                    const uint32_t pnr = ph & 0xffffu, pnv = ph >> 16;
                    if (pnr < 0xfffeu) {
                        const uint8_t* pbm = pv + RJ_PAGE - ((pnr + 7) >> 3);
                        ok = pnr > 0 && pnv > 0 && ((pbm[(pnr - 1) >> 3] >> ((pnr - 1) & 7)) & 1u);
                    }
                }
                if (!ok) atomicOr(err_flags, RJ_ERR_ORPHAN_LONG_PAGE);
            }
            continue;
        }
        if (n_r == 0xffffu) {
            if (lane == 0) {
                // total length = this page's chars + every following 0xfffe page (build_table.cpp:392-405)
                uint64_t total = n_v;
                for (uint64_t q = p + 1; q < n_pages; ++q) {
                    uint32_t h2 = *reinterpret_cast<const uint32_t*>(pages + q * RJ_PAGE);
                    if ((h2 & 0xffffu) != 0xfffeu) break;
                    total += h2 >> 16;
                }
                if (total > RJ_DESC_LEN_MASK) total = RJ_DESC_LEN_MASK;
                desc[r0] = (p * RJ_PAGE + 4) | (total << RJ_DESC_LEN_SHIFT) | RJ_DESC_LONG;
                if (out_valid) atomicOr(&out_valid[r0 >> 5], 1u << (r0 & 31));
            }
            continue;
        }
        const uint16_t* offs = reinterpret_cast<const uint16_t*>(pg + 4);
        const uint64_t  data = p * RJ_PAGE + 4 + 2ull * n_v;
        const uint8_t*  bm   = pg + RJ_PAGE - ((n_r + 7) >> 3);
        const bool      dense = (n_v == n_r);
        uint32_t running = 0;
        for (uint32_t base = 0; base < n_r; base += 32) {
            const uint32_t i  = base + lane;
            const bool     in = i < n_r;
            uint32_t bit = in ? 1u : 0u;
            if (!dense && in) bit = (bm[i >> 3] >> (i & 7)) & 1u;
            const uint32_t word = __ballot_sync(RJ_FULL_MASK, bit);
            uint32_t rank = running + __popc(word & lt);
            rank = rank < 4094 ? rank : 4093;
            uint64_t d = 0;
            if (bit) {
                const uint32_t end   = offs[rank];
                const uint32_t start = rank ? offs[rank - 1] : 0u;
                d = (data + start) | (static_cast<uint64_t>(end - start) << RJ_DESC_LEN_SHIFT);
            }
            if (in) desc[r0 + i] = d;
            running += __popc(word);
            if (out_valid != nullptr && lane == 0 && word) or_valid_word(out_valid, r0 + base, word);
        }
        // a 0xfffe page right behind this one continues its LAST row (see above): that row becomes a chain
        __syncwarp();
        if (lane == 0 && n_r > 0 && n_v > 0 && n_v <= 4094 && p + 1 < n_pages) {
            const uint32_t h2 = *reinterpret_cast<const uint32_t*>(pg + RJ_PAGE);
            if ((h2 & 0xffffu) == 0xfffeu && ((bm[(n_r - 1) >> 3] >> ((n_r - 1) & 7)) & 1u)) {
                const uint32_t end   = offs[n_v - 1];
                const uint32_t start = n_v > 1 ? offs[n_v - 2] : 0u;
                uint64_t total = end - start;
                for (uint64_t q = p + 1; q < n_pages; ++q) {
                    const uint32_t hq = *reinterpret_cast<const uint32_t*>(pages + q * RJ_PAGE);
                    if ((hq & 0xffffu) != 0xfffeu) break;
                    total += hq >> 16;
                }
                if (total > RJ_DESC_LEN_MASK) total = RJ_DESC_LEN_MASK;
                desc[r0 + n_r - 1] = (data + start) | (total << RJ_DESC_LEN_SHIFT) | RJ_DESC_LONG;
            }
        }
    }
}

} // namespace

void launch_page_rows(const void* pages, uint64_t n_pages, int type, uint32_t* rows, uint64_t* totals,
                      cudaStream_t s) {
    if (n_pages == 0) return;
    unsigned blocks = static_cast<unsigned>((n_pages + 255) / 256);
    page_rows_kernel<<<blocks, 256, 0, s>>>(static_cast<const uint8_t*>(pages), n_pages, type, rows, totals);
    RJ_LAUNCH_CHECK();
}

void launch_decode_fixed(const void* pages, uint64_t n_pages, int type, const uint64_t* row_start, void* values,
                         uint32_t* valid, int sm_count, cudaStream_t s) {
    if (n_pages == 0) return;
    const size_t smem = kDecWarps * kDecStages * RJ_PAGE + kDecWarps * kDecStages * sizeof(uint64_t);
    // 3 CTAs of 64 KB per SM; never more warps than pages
    uint64_t want = (n_pages + kDecWarps - 1) / kDecWarps;
    unsigned blocks = static_cast<unsigned>(want < static_cast<uint64_t>(sm_count) * 3 ? want : static_cast<uint64_t>(sm_count) * 3);
    const uint8_t* pg = static_cast<const uint8_t*>(pages);
    if (type == RJ_INT32) {
        static SmemConfigured cfg;
        cfg.ensure(decode_fixed_kernel<uint32_t>, smem);
        decode_fixed_kernel<uint32_t><<<blocks, kDecWarps * 32, smem, s>>>(pg, n_pages, row_start,
                                                                          static_cast<uint32_t*>(values), valid);
    } else {
        static SmemConfigured cfg;
        cfg.ensure(decode_fixed_kernel<uint64_t>, smem);
        decode_fixed_kernel<uint64_t><<<blocks, kDecWarps * 32, smem, s>>>(pg, n_pages, row_start,
                                                                          static_cast<uint64_t*>(values), valid);
    }
    RJ_LAUNCH_CHECK();
}

void launch_decode_varchar(const void* pages, uint64_t n_pages, const uint64_t* row_start, uint64_t* desc,
                           uint32_t* valid, int sm_count, cudaStream_t s, uint32_t* err_flags) {
    if (n_pages == 0) return;
    uint64_t want = (n_pages + 3) / 4;
    unsigned blocks = static_cast<unsigned>(want < static_cast<uint64_t>(sm_count) * 8 ? want : static_cast<uint64_t>(sm_count) * 8);
    decode_varchar_kernel<<<blocks, 128, 0, s>>>(static_cast<const uint8_t*>(pages), n_pages, row_start, desc, valid, err_flags);
    RJ_LAUNCH_CHECK();
}

} // namespace rj
