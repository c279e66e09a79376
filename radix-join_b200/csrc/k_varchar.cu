// VARCHAR support: string join keys and VARCHAR page output.
//
// Strings are never copied before the root: a decoded VARCHAR column is a descriptor per row
// (byte address inside the column's page buffer, length, "long" flag for 0xffff/0xfffe page chains,
// see rj_decode_varchar) and the characters stay in the uploaded pages.
//
// Output (replaces the VARCHAR branch of Table::to_columnar, reference src/build_table.cpp:595-677):
// the reference packs rows greedily (:650-653).  Only the decoded rows are observable and an
// under-filled page is legal, so the layout is made parallel: every row gets a weight in BITS
// (16 + 8*len + 1 for a string, 1 for a NULL), a prefix sum places rows on a bit axis, and row j goes
// to page floor(start_j / C) of its segment with C = 65497 - Wmax + 1, which guarantees
// 4 + 2*n_v + chars + ceil(n_r/8) <= 8192 for every page.  Strings longer than 1020 bytes break the
// axis into segments: up to 8185 bytes they get a page of their own, above that the reference's
// long-string chain (:603-619,:644-648: 0xffff page + 0xfffe pages of <= 8188 chars).
#include "rj_common.cuh"
#include "rj_internal.h"

namespace rj {
namespace {

constexpr uint32_t kSoloLen   = 1020;              // longer strings get their own page(s)
constexpr uint32_t kLongLen   = RJ_PAGE - 7;       // > 8185: long-string chain (build_table.cpp:644)
constexpr uint32_t kChunk     = RJ_PAGE - 4;       // chars per long-string page (:614)
constexpr uint64_t kCapBits   = 8ull * (RJ_PAGE - 4) - 7;
constexpr int      kWrThreads = 256;

struct Str {
    uint64_t addr;
    uint32_t len;
    bool     is_long;
};

__device__ __forceinline__ Str unpack(uint64_t d) {
    Str s;
    s.addr    = d & RJ_DESC_ADDR_MASK;
    s.len     = static_cast<uint32_t>((d >> RJ_DESC_LEN_SHIFT) & RJ_DESC_LEN_MASK);
    s.is_long = (d & RJ_DESC_LONG) != 0;
    return s;
}

// Calls f(ptr, n, pos) for every contiguous piece of chars [a, b) of the string.
template <class F>
__device__ __forceinline__ void for_each_piece(const uint8_t* __restrict__ pages, const Str& s, uint32_t a,
                                               uint32_t b, F f) {
    if (!s.is_long) {
        if (b > a) f(pages + s.addr + a, b - a, a);
        return;
    }
    // chain of pages.  Its head is a 0xffff page (n_v chars at +4) or -- malformed input the reference
    // nevertheless decodes by appending (build_table.cpp:392-405) -- the last string of a regular page,
    // which runs from its address to the end of that page's character data; every following 0xfffe page
    // adds its n_v chars at +4.
    uint64_t q   = s.addr / RJ_PAGE;
    uint32_t pos = 0;
    bool     head = true;
    while (pos < b) {
        const uint8_t* pg  = pages + q * RJ_PAGE;
        const uint32_t hdr = *reinterpret_cast<const uint32_t*>(pg);
        const uint8_t* src = pg + 4;
        uint32_t       n   = hdr >> 16;
        if (head && (hdr & 0xffffu) != 0xffffu) {
            const uint32_t n_v = hdr >> 16;
            const uint32_t end = 4 + 2 * n_v + (n_v ? reinterpret_cast<const uint16_t*>(pg + 4)[n_v - 1] : 0u);
            src = pages + s.addr;
            n   = end - static_cast<uint32_t>(s.addr - q * RJ_PAGE);
        }
        head = false;
        const uint32_t lo = pos > a ? pos : a;
        const uint32_t hi = pos + n < b ? pos + n : b;
        if (hi > lo) f(src + (lo - pos), hi - lo, lo);
        pos += n;
        ++q;
        if (n == 0) break; // malformed chain
    }
}

__global__ void __launch_bounds__(256)
    varchar_hash_kernel(const uint8_t* __restrict__ pages, const uint64_t* __restrict__ desc,
                        const uint32_t* __restrict__ valid, uint64_t n, uint64_t* __restrict__ out) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t h = 0;
        if (valid == nullptr || test_bit(valid, i)) {
            const Str s = unpack(desc[i]);
            h = 14695981039346656037ull ^ s.len;
            for_each_piece(pages, s, 0, s.len, [&](const uint8_t* p, uint32_t m, uint32_t) {
                for (uint32_t k = 0; k < m; ++k) {
                    h ^= p[k];
                    h *= 1099511628211ull;
                }
            });
        }
        out[i] = h;
    }
}

__device__ uint8_t char_at(const uint8_t* __restrict__ pages, const Str& s, uint32_t pos) {
    uint8_t c = 0;
    for_each_piece(pages, s, pos, pos + 1, [&](const uint8_t* p, uint32_t, uint32_t) { c = p[0]; });
    return c;
}

__global__ void __launch_bounds__(256)
    varchar_pairs_equal_kernel(const uint8_t* __restrict__ pages_a, const uint64_t* __restrict__ desc_a,
                               const uint32_t* __restrict__ idx_a, const uint8_t* __restrict__ pages_b,
                               const uint64_t* __restrict__ desc_b, const uint32_t* __restrict__ idx_b, uint64_t n,
                               uint32_t* __restrict__ keep) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const Str a = unpack(desc_a[idx_a[i]]);
        const Str b = unpack(desc_b[idx_b[i]]);
        bool eq = a.len == b.len;
        if (eq) {
            if (!a.is_long && !b.is_long) {
                const uint8_t* pa = pages_a + a.addr;
                const uint8_t* pb = pages_b + b.addr;
                for (uint32_t k = 0; k < a.len && eq; ++k) eq = pa[k] == pb[k];
            } else {
                for (uint32_t k = 0; k < a.len && eq; ++k) eq = char_at(pages_a, a, k) == char_at(pages_b, b, k);
            }
        }
        keep[i] = eq ? 1u : 0u;
    }
}

__global__ void __launch_bounds__(256)
    compact_pairs_kernel(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, const uint32_t* __restrict__ keep,
                         const uint64_t* __restrict__ pos, uint64_t n, uint32_t* __restrict__ out_a,
                         uint32_t* __restrict__ out_b) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (keep[i]) {
            out_a[pos[i]] = a[i];
            out_b[pos[i]] = b[i];
        }
    }
}

// ---- output layout ---------------------------------------------------------------------------------
// kind of output row j: 0 = NULL, 1 = packed string, 2 = solo page, 3 = long chain
struct RowInfo {
    Str      s;
    uint32_t kind;
    uint32_t weight; // bits on the packing axis (0 for breakers)
    uint32_t pages;  // pages started by this row if it is a breaker
};

__device__ __forceinline__ RowInfo row_info(const VarcharLayoutDev& L, uint64_t j) {
    RowInfo        r;
    const uint32_t src = L.idx != nullptr ? L.idx[j] : static_cast<uint32_t>(j);
    const bool     v   = L.valid == nullptr || test_bit(L.valid, src);
    r.s = unpack(v ? L.desc[src] : 0ull);
    if (!v) {
        r.kind = 0; r.weight = 1; r.pages = 0;
    } else if (r.s.len <= kSoloLen) {
        r.kind = 1; r.weight = 8 * (2 + r.s.len) + 1; r.pages = 0;
    } else if (r.s.len <= kLongLen) {
        r.kind = 2; r.weight = 0; r.pages = 1;
    } else {
        r.kind = 3; r.weight = 0; r.pages = (r.s.len + kChunk - 1) / kChunk;
    }
    return r;
}

__global__ void __launch_bounds__(256)
    varchar_weights_kernel(VarcharLayoutDev L, uint64_t* __restrict__ weights) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    uint32_t wmax = 0;
    for (uint64_t j = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; j < L.n; j += stride) {
        const RowInfo r = row_info(L, j);
        weights[j]      = r.weight;
        L.head_pages[j] = r.pages; // breakers: final value; packed rows are decided by varchar_heads_kernel
        if (r.kind == 1 && r.weight > wmax) wmax = r.weight;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        uint32_t o = __shfl_xor_sync(RJ_FULL_MASK, wmax, d);
        wmax = o > wmax ? o : wmax;
    }
    if ((threadIdx.x & 31) == 0 && wmax) atomicMax(reinterpret_cast<unsigned long long*>(L.scalars), static_cast<unsigned long long>(wmax));
}

// marks[j] = weight prefix at row j if j is a breaker, else 0 (input of the max-scan)
__global__ void __launch_bounds__(256) varchar_marks_kernel(VarcharLayoutDev L, uint64_t* __restrict__ marks) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t j = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; j < L.n; j += stride) {
        marks[j] = L.head_pages[j] ? L.weight_scan[j] : 0ull;
    }
}

__device__ __forceinline__ uint64_t page_capacity_bits(const VarcharLayoutDev& L) {
    const uint64_t wmax = L.scalars[0];
    return wmax ? kCapBits - wmax + 1 : kCapBits;
}

__global__ void __launch_bounds__(256)
    varchar_heads_kernel(VarcharLayoutDev L, const uint64_t* __restrict__ weights) {
    const uint64_t cap = page_capacity_bits(L);
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t j = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; j < L.n; j += stride) {
        if (L.head_pages[j]) continue; // breaker: already a head
        bool head = true;
        if (j > 0 && weights[j - 1] != 0) {
            // predecessor is a packed/NULL row of the same segment (breakers have weight 0, packed rows >= 1)
            const uint64_t base = L.base_scan[j];
            const uint64_t q    = (L.weight_scan[j] - weights[j] - base) / cap;
            const uint64_t qp   = (L.weight_scan[j - 1] - weights[j - 1] - base) / cap;
            head = q != qp;
        }
        L.head_pages[j] = head ? 1u : 0u;
    }
}

__global__ void __launch_bounds__(256) varchar_page_rows_kernel(VarcharLayoutDev L) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t j = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; j < L.n; j += stride) {
        const uint32_t k = L.head_pages[j];
        const uint64_t p = L.page_of[j];
        for (uint32_t c = 0; c < k; ++c) L.page_row[p + c] = static_cast<uint32_t>(j);
    }
}

// One CTA per output page.  The page is assembled in shared memory and leaves with one TMA bulk store.
__global__ void __launch_bounds__(kWrThreads) varchar_write_kernel(VarcharLayoutDev L, uint8_t* __restrict__ pages_out) {
    __shared__ __align__(128) uint8_t buf[RJ_PAGE];
    __shared__ uint64_t s_warp[kWrThreads / 32];
    __shared__ uint64_t s_carry;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint8_t* __restrict__ src_pages = L.src_pages;

    for (uint64_t pg = blockIdx.x; pg < L.n_pages; pg += gridDim.x) {
        // the previous page's bulk store must have finished reading buf
        if (threadIdx.x == 0) tma_store_wait_read<0>();
        __syncthreads();
        uint4* b4 = reinterpret_cast<uint4*>(buf);
        for (uint32_t k = threadIdx.x; k < RJ_PAGE / 16; k += kWrThreads) b4[k] = make_uint4(0, 0, 0, 0);
        __syncthreads();

        const uint64_t j0 = L.page_row[pg];
        const RowInfo  r0 = row_info(L, j0);
        if (r0.kind >= 2) {
            // ---- solo page or one page of a long chain ---------------------------------------------
            uint32_t a = 0, b = r0.s.len, dst = 4;
            if (r0.kind == 3) {
                const uint32_t c = static_cast<uint32_t>(pg - L.page_of[j0]);
                a = c * kChunk;
                b = a + kChunk < r0.s.len ? a + kChunk : r0.s.len;
                if (threadIdx.x == 0) *reinterpret_cast<uint32_t*>(buf) = (c == 0 ? 0xffffu : 0xfffeu) | ((b - a) << 16);
            } else {
                dst = 6; // one offset
                if (threadIdx.x == 0) {
                    *reinterpret_cast<uint32_t*>(buf) = 1u | (1u << 16);
                    *reinterpret_cast<uint16_t*>(buf + 4) = static_cast<uint16_t>(r0.s.len);
                    buf[RJ_PAGE - 1] = 1;
                }
            }
            for_each_piece(src_pages, r0.s, a, b, [&](const uint8_t* p, uint32_t m, uint32_t pos) {
                for (uint32_t k = threadIdx.x; k < m; k += kWrThreads) buf[dst + (pos - a) + k] = p[k];
            });
        } else {
            // ---- packed page: rows [j0, j1) ------------------------------------------------------------
            const uint64_t j1  = pg + 1 < L.n_pages ? L.page_row[pg + 1] : L.n;
            const uint32_t n_r = static_cast<uint32_t>(j1 - j0);
            uint8_t*       bm  = buf + RJ_PAGE - ((n_r + 7) >> 3);
            uint32_t       n_v = 0;
            // two sweeps over the rows: sweep 0 writes offsets + bitmap and counts n_v,
            // sweep 1 copies the characters behind the offset array
            for (int sweep = 0; sweep < 2; ++sweep) {
                if (threadIdx.x == 0) s_carry = 0;
                __syncthreads();
                for (uint32_t base = 0; base < n_r; base += kWrThreads) {
                    const uint32_t i  = base + threadIdx.x;
                    const bool     in = i < n_r;
                    RowInfo r;
                    r.kind = 0; r.s.len = 0; r.s.addr = 0; r.s.is_long = false;
                    if (in) r = row_info(L, j0 + i);
                    const bool     v = in && r.kind == 1;
                    // packed (count << 32 | chars) inclusive scan across the block
                    uint64_t x = v ? ((1ull << 32) | r.s.len) : 0ull;
                    uint64_t inc = x;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        uint64_t o = __shfl_up_sync(RJ_FULL_MASK, inc, d);
                        if (lane >= d) inc += o;
                    }
                    if (lane == 31) s_warp[warp] = inc;
                    __syncthreads();
                    uint64_t prefix = s_carry;
                    for (uint32_t w = 0; w < warp; ++w) prefix += s_warp[w];
                    const uint64_t mine_inc = prefix + inc;
                    const uint32_t vrank = static_cast<uint32_t>((mine_inc - x) >> 32);
                    const uint32_t cend  = static_cast<uint32_t>(mine_inc & 0xffffffffu);
                    if (sweep == 0) {
                        if (v) *reinterpret_cast<uint16_t*>(buf + 4 + 2 * vrank) = static_cast<uint16_t>(cend);
                        const uint32_t word = __ballot_sync(RJ_FULL_MASK, v);
                        if (lane < 4 && base + warp * 32 + lane * 8 < n_r) bm[((base + warp * 32) >> 3) + lane] = static_cast<uint8_t>(word >> (8 * lane));
                    } else {
                        // warp-cooperative character copy: one string at a time, lanes stride its bytes
                        const uint32_t cstart = cend - r.s.len;
                        const uint32_t vmask = __ballot_sync(RJ_FULL_MASK, v && r.s.len > 0);
                        uint32_t m = vmask;
                        while (m) {
                            const int      src_lane = __ffs(m) - 1;
                            m &= m - 1;
                            const uint64_t addr = __shfl_sync(RJ_FULL_MASK, r.s.addr, src_lane);
                            const uint32_t len  = __shfl_sync(RJ_FULL_MASK, r.s.len, src_lane);
                            const uint32_t lng  = __shfl_sync(RJ_FULL_MASK, r.s.is_long ? 1u : 0u, src_lane);
                            const uint32_t at   = __shfl_sync(RJ_FULL_MASK, cstart, src_lane);
                            uint8_t* d = buf + 4 + 2 * n_v + at;
                            if (!lng) {
                                const uint8_t* p = src_pages + addr;
                                for (uint32_t k = lane; k < len; k += 32) d[k] = p[k];
                            } else {
                                Str s; s.addr = addr; s.len = len; s.is_long = true;
                                for_each_piece(src_pages, s, 0, len, [&](const uint8_t* p, uint32_t cnt, uint32_t pos) {
                                    for (uint32_t k = lane; k < cnt; k += 32) d[pos + k] = p[k];
                                });
                            }
                        }
                    }
                    __syncthreads();
                    if (threadIdx.x == kWrThreads - 1) s_carry = mine_inc;
                    __syncthreads();
                }
                if (sweep == 0) {
                    n_v = static_cast<uint32_t>(s_carry >> 32);
                    if (threadIdx.x == 0) *reinterpret_cast<uint32_t*>(buf) = n_r | (n_v << 16);
                }
                __syncthreads();
            }
        }
        fence_proxy_async_smem();
        __syncthreads();
        if (threadIdx.x == 0) {
            tma_store_1d(pages_out + pg * RJ_PAGE, buf, RJ_PAGE);
            tma_store_commit();
        }
    }
    if (threadIdx.x == 0) tma_store_wait_all<0>();
}

unsigned grid_for(uint64_t n, int per_block, int sm_count, int waves) {
    uint64_t want = (n + per_block - 1) / per_block;
    uint64_t cap  = static_cast<uint64_t>(sm_count) * waves;
    if (want < 1) want = 1;
    return static_cast<unsigned>(want < cap ? want : cap);
}

} // namespace

void launch_varchar_hash(const uint8_t* pages, const uint64_t* desc, const uint32_t* valid, uint64_t n,
                         uint64_t* out_hash, int sm_count, cudaStream_t s) {
    if (n == 0) return;
    varchar_hash_kernel<<<grid_for(n, 256, sm_count, 16), 256, 0, s>>>(pages, desc, valid, n, out_hash);
    RJ_LAUNCH_CHECK();
}

void launch_varchar_pairs_equal(const uint8_t* pages_a, const uint64_t* desc_a, const uint32_t* idx_a,
                                const uint8_t* pages_b, const uint64_t* desc_b, const uint32_t* idx_b, uint64_t n,
                                uint32_t* keep, int sm_count, cudaStream_t s) {
    if (n == 0) return;
    varchar_pairs_equal_kernel<<<grid_for(n, 256, sm_count, 16), 256, 0, s>>>(pages_a, desc_a, idx_a, pages_b, desc_b,
                                                                              idx_b, n, keep);
    RJ_LAUNCH_CHECK();
}

void launch_compact_pairs(const uint32_t* a, const uint32_t* b, const uint32_t* keep, const uint64_t* pos, uint64_t n,
                          uint32_t* out_a, uint32_t* out_b, cudaStream_t s) {
    if (n == 0) return;
    uint64_t want = (n + 255) / 256;
    compact_pairs_kernel<<<static_cast<unsigned>(want < 65535 * 8 ? want : 65535 * 8), 256, 0, s>>>(a, b, keep, pos, n, out_a, out_b);
    RJ_LAUNCH_CHECK();
}

void launch_varchar_weights(const VarcharLayoutDev& L, uint64_t* weights, int sm_count, cudaStream_t s) {
    if (L.n == 0) return;
    varchar_weights_kernel<<<grid_for(L.n, 256, sm_count, 16), 256, 0, s>>>(L, weights);
    RJ_LAUNCH_CHECK();
}

void launch_varchar_marks(const VarcharLayoutDev& L, uint64_t* marks, int sm_count, cudaStream_t s) {
    if (L.n == 0) return;
    varchar_marks_kernel<<<grid_for(L.n, 256, sm_count, 16), 256, 0, s>>>(L, marks);
    RJ_LAUNCH_CHECK();
}

void launch_varchar_heads(const VarcharLayoutDev& L, const uint64_t* weights, int sm_count, cudaStream_t s) {
    if (L.n == 0) return;
    varchar_heads_kernel<<<grid_for(L.n, 256, sm_count, 16), 256, 0, s>>>(L, weights);
    RJ_LAUNCH_CHECK();
}

void launch_varchar_page_rows(const VarcharLayoutDev& L, int sm_count, cudaStream_t s) {
    if (L.n == 0) return;
    varchar_page_rows_kernel<<<grid_for(L.n, 256, sm_count, 16), 256, 0, s>>>(L);
    RJ_LAUNCH_CHECK();
}

void launch_varchar_write(const VarcharLayoutDev& L, uint8_t* pages_out, int sm_count, cudaStream_t s) {
    if (L.n_pages == 0) return;
    uint64_t cap = static_cast<uint64_t>(sm_count) * 8;
    unsigned blocks = static_cast<unsigned>(L.n_pages < cap ? L.n_pages : cap);
    varchar_write_kernel<<<blocks, kWrThreads, 0, s>>>(L, pages_out);
    RJ_LAUNCH_CHECK();
}


namespace {
__global__ void __launch_bounds__(256)
    varchar_desc_from_offsets_kernel(const uint64_t* __restrict__ off, uint64_t n, uint64_t* __restrict__ desc) {
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint64_t a = off[i], len = off[i + 1] - a;
        desc[i] = (a & RJ_DESC_ADDR_MASK) | ((len & RJ_DESC_LEN_MASK) << RJ_DESC_LEN_SHIFT);
    }
}
} // namespace

// Descriptors of strings that lie back to back in one buffer (row i = bytes [off[i], off[i+1])): what the page
// writer (launch_varchar_plan / write) takes, so dense strings on the device become pages without a host pass --
// the role of ColumnInserter<std::string> (reference include/plan.h:230-335).
void launch_varchar_desc_from_offsets(const uint64_t* off, uint64_t n, uint64_t* desc, int sm_count, cudaStream_t s) {
    if (n == 0) return;
    const uint64_t want = (n + 255) / 256, cap = static_cast<uint64_t>(sm_count) * 8;
    varchar_desc_from_offsets_kernel<<<static_cast<unsigned>(want < cap ? want : cap), 256, 0, s>>>(off, n, desc);
    RJ_LAUNCH_CHECK();
}

} // namespace rj
