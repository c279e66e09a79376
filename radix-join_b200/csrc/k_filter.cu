// Pre-filter evaluation on decoded device columns -- the step immediately before the hot path in the
// reference's harness: Comparison::eval / LogicalOperation::eval over InnerColumns
// (reference src/statement.cpp:46-133,186-200; include/inner_column.h:170-325 fixed-width predicates,
// :386-562 string predicates; include/statement.h:118-161 LIKE) and the selection of the rows that pass
// (src/build_table.cpp:94-119 walks the result bitmap row by row).
//
// Same result layout as the reference: one bit per row, LSB first (a uint32 word here = four of its
// bytes), bit = "row is not NULL and the comparison holds"; IS NULL / IS NOT NULL are the validity bitmap
// itself; AND / OR / NOT are word-wise, and NOT flips NULL rows to true exactly like bitmap_not
// (statement.cpp:8-16).  Bits past the last row are kept zero.
//
// Kernels: a warp evaluates 32 rows and writes the ballot as one output word (coalesced value loads, one
// validity word per warp); strings are read in place in the page buffer through their descriptors (long
// strings are walked across their 0xffff / 0xfffe page chain); the row-id list of a bitmap comes from a
// per-word popcount, the engine's exclusive scan, and a warp-per-word expansion.
#include "rj_common.cuh"
#include "rj_internal.h"

namespace rj {
namespace {

constexpr int kFilterThreads = 256;

enum : int { OP_EQ = 0, OP_NEQ, OP_LT, OP_GT, OP_LEQ, OP_GEQ, OP_LIKE, OP_NOT_LIKE, OP_IS_NULL, OP_IS_NOT_NULL };

template <typename T>
__device__ __forceinline__ bool compare(T v, T rhs, int op) {
    switch (op) {
    case OP_EQ: return v == rhs;
    case OP_NEQ: return v != rhs;
    case OP_LT: return v < rhs;
    case OP_GT: return v > rhs;
    case OP_LEQ: return v <= rhs;
    default: return v >= rhs;
    }
}

template <typename T>
__global__ void __launch_bounds__(kFilterThreads)
    filter_compare_kernel(const T* __restrict__ values, const uint32_t* __restrict__ valid, uint64_t n, int op, T rhs,
                          uint32_t* __restrict__ out) {
    const uint64_t n_words = (n + 31) / 32;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warps = static_cast<uint64_t>(gridDim.x) * (kFilterThreads / 32);
    for (uint64_t w = static_cast<uint64_t>(blockIdx.x) * (kFilterThreads / 32) + (threadIdx.x >> 5); w < n_words; w += warps) {
        const uint64_t i = w * 32 + lane;
        bool ok = i < n;
        if (ok && valid != nullptr) ok = (valid[w] >> lane) & 1u;
        if (ok) ok = compare<T>(values[i], rhs, op);
        const uint32_t word = __ballot_sync(RJ_FULL_MASK, ok);
        if (lane == 0) out[w] = word;
    }
}

// IS NOT NULL: the validity bitmap (all ones when the column holds no NULL); IS NULL: its complement
__global__ void __launch_bounds__(kFilterThreads)
    filter_null_kernel(const uint32_t* __restrict__ valid, uint64_t n, int want_null, uint32_t* __restrict__ out) {
    const uint64_t n_words = (n + 31) / 32;
    for (uint64_t w = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; w < n_words; w += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        uint32_t v = valid ? valid[w] : 0xffffffffu;
        if (want_null) v = ~v;
        const uint64_t left = n - w * 32;
        if (left < 32) v &= (1u << left) - 1u;
        out[w] = v;
    }
}

// op: 0 = AND, 1 = OR, 2 = NOT (b ignored)
__global__ void __launch_bounds__(kFilterThreads)
    bitmap_logic_kernel(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, uint64_t n, int op, uint32_t* __restrict__ out) {
    const uint64_t n_words = (n + 31) / 32;
    for (uint64_t w = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; w < n_words; w += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        uint32_t v = op == 0 ? (a[w] & b[w]) : (op == 1 ? (a[w] | b[w]) : ~a[w]);
        const uint64_t left = n - w * 32;
        if (left < 32) v &= (1u << left) - 1u;
        out[w] = v;
    }
}

// ---- strings -------------------------------------------------------------------------------------------
// Sequential reader of one string in the page buffer.  Short strings are `len` contiguous bytes at
// `addr`; a long string starts in a 0xffff page (chars from byte 4, count at byte 2) and continues in
// the 0xfffe pages behind it (src/build_table.cpp:384-405).
struct StrReader {
    const uint8_t* pages;
    uint64_t       at;        // address of the next byte
    uint32_t       left;      // bytes left in the string
    uint32_t       page_left; // bytes left in the current page (long strings)
    bool           is_long;
    __device__ StrReader(const uint8_t* pg, uint64_t desc): pages(pg) {
        at = desc & RJ_DESC_ADDR_MASK;
        left = static_cast<uint32_t>((desc >> RJ_DESC_LEN_SHIFT) & RJ_DESC_LEN_MASK);
        is_long = (desc & RJ_DESC_LONG) != 0;
        page_left = left;
        if (is_long) {
            const uint64_t page = at / RJ_PAGE * RJ_PAGE;
            const uint32_t in_page = ld_u16_unaligned(pages + page + 2);
            page_left = in_page - static_cast<uint32_t>(at - page - 4);
        }
    }
    __device__ bool end() const { return left == 0; }
    __device__ uint8_t next() {
        if (is_long && page_left == 0) {
            const uint64_t page = (at - 1) / RJ_PAGE * RJ_PAGE + RJ_PAGE; // the continuation page
            page_left = ld_u16_unaligned(pages + page + 2);
            at = page + 4;
        }
        const uint8_t c = pages[at];
        ++at;
        --left;
        --page_left;
        return c;
    }
};

// <0, 0, >0 like std::string_view::compare (bytes as unsigned char, then length)
__device__ int str_compare(const uint8_t* pages, uint64_t desc, const uint8_t* rhs, uint32_t rhs_len) {
    StrReader r(pages, desc);
    uint32_t  i = 0;
    while (!r.end() && i < rhs_len) {
        const uint8_t c = r.next();
        if (c != rhs[i]) return c < rhs[i] ? -1 : 1;
        ++i;
    }
    if (r.end() && i == rhs_len) return 0;
    return r.end() ? -1 : 1;
}

// SQL LIKE as the reference defines it (statement.h:118-161): '%' -> ".*", '_' -> ".", everything else
// literal, FULL match under RE2's defaults -- '.' is any character except '\n', and a character is a
// UTF-8 code point.  Iterative wildcard matching with one backtrack point (the last '%'); the string is
// addressed by byte index, so a long string is materialised byte by byte through `byte_at`.
struct StrIndex {
    const uint8_t* pages;
    uint64_t       addr;
    uint32_t       len;
    bool           is_long;
    __device__ StrIndex(const uint8_t* pg, uint64_t desc): pages(pg) {
        addr = desc & RJ_DESC_ADDR_MASK;
        len = static_cast<uint32_t>((desc >> RJ_DESC_LEN_SHIFT) & RJ_DESC_LEN_MASK);
        is_long = (desc & RJ_DESC_LONG) != 0;
    }
    __device__ uint8_t byte_at(uint32_t i) const {
        if (!is_long) return pages[addr + i];
        uint64_t page = addr / RJ_PAGE * RJ_PAGE; // walk the chain: long strings are rare
        uint32_t skip = i;
        for (;;) {
            const uint32_t in_page = ld_u16_unaligned(pages + page + 2);
            if (skip < in_page) return pages[page + 4 + skip];
            skip -= in_page;
            page += RJ_PAGE;
        }
    }
};

__device__ __forceinline__ uint32_t utf8_len(uint8_t lead) { return lead < 0x80 ? 1u : (lead < 0xe0 ? 2u : (lead < 0xf0 ? 3u : 4u)); }

__device__ bool like_match(const StrIndex& s, const uint8_t* pat, uint32_t pat_len) {
    uint32_t si = 0, pi = 0;
    uint32_t star_p = 0xffffffffu, star_s = 0; // pattern index after the last '%', string index it was tried at
    while (si < s.len) {
        if (pi < pat_len && pat[pi] == '%') {
            star_p = ++pi;
            star_s = si;
            continue;
        }
        const uint8_t c = s.byte_at(si);
        bool step = false;
        uint32_t adv = 1;
        if (pi < pat_len) {
            if (pat[pi] == '_') {
                step = c != '\n';
                adv = utf8_len(c);
                if (si + adv > s.len) adv = s.len - si;
            } else {
                step = c == pat[pi];
            }
        }
        if (step) {
            si += adv;
            ++pi;
            continue;
        }
        // mismatch: let the last '%' swallow one more character (never a newline)
        if (star_p == 0xffffffffu) return false;
        const uint8_t sc = s.byte_at(star_s);
        if (sc == '\n') return false;
        star_s += utf8_len(sc);
        if (star_s > s.len) star_s = s.len;
        si = star_s;
        pi = star_p;
    }
    while (pi < pat_len && pat[pi] == '%') ++pi;
    return pi == pat_len;
}

__global__ void __launch_bounds__(kFilterThreads)
    filter_varchar_kernel(const uint8_t* __restrict__ pages, const uint64_t* __restrict__ desc, const uint32_t* __restrict__ valid,
                          uint64_t n, int op, const uint8_t* __restrict__ rhs, uint32_t rhs_len, uint32_t* __restrict__ out) {
    const uint64_t n_words = (n + 31) / 32;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warps = static_cast<uint64_t>(gridDim.x) * (kFilterThreads / 32);
    for (uint64_t w = static_cast<uint64_t>(blockIdx.x) * (kFilterThreads / 32) + (threadIdx.x >> 5); w < n_words; w += warps) {
        const uint64_t i = w * 32 + lane;
        bool ok = i < n;
        if (ok && valid != nullptr) ok = (valid[w] >> lane) & 1u;
        if (ok) {
            const uint64_t d = desc[i];
            if (op == OP_LIKE || op == OP_NOT_LIKE) {
                const bool m = like_match(StrIndex(pages, d), rhs, rhs_len);
                ok = op == OP_LIKE ? m : !m;
            } else {
                ok = compare<int>(str_compare(pages, d, rhs, rhs_len), 0, op);
            }
        }
        const uint32_t word = __ballot_sync(RJ_FULL_MASK, ok);
        if (lane == 0) out[w] = word;
    }
}

// ---- bitmap -> row ids -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFilterThreads)
    bitmap_popc_kernel(const uint32_t* __restrict__ bits, uint64_t n_words, uint32_t* __restrict__ counts) {
    for (uint64_t w = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; w < n_words; w += static_cast<uint64_t>(gridDim.x) * blockDim.x)
        counts[w] = __popc(bits[w]);
}

__global__ void __launch_bounds__(kFilterThreads)
    bitmap_expand_kernel(const uint32_t* __restrict__ bits, const uint64_t* __restrict__ start, uint64_t n_words, uint32_t* __restrict__ ids) {
    const uint32_t lane = threadIdx.x & 31, lt = lanemask_lt();
    const uint64_t warps = static_cast<uint64_t>(gridDim.x) * (kFilterThreads / 32);
    for (uint64_t w = static_cast<uint64_t>(blockIdx.x) * (kFilterThreads / 32) + (threadIdx.x >> 5); w < n_words; w += warps) {
        const uint32_t word = bits[w];
        if ((word >> lane) & 1u) ids[start[w] + __popc(word & lt)] = static_cast<uint32_t>(w * 32 + lane);
    }
}

unsigned grid_for(uint64_t items, int per_block, int sm_count) {
    const uint64_t want = (items + per_block - 1) / per_block;
    const uint64_t cap = static_cast<uint64_t>(sm_count) * 8;
    return static_cast<unsigned>(want < 1 ? 1 : (want < cap ? want : cap));
}

} // namespace

void launch_filter_compare(const void* values, const uint32_t* valid, uint64_t n, int type, int op, int64_t rhs_i, double rhs_d,
                           uint32_t* out, int sm_count, cudaStream_t s) {
    if (n == 0) return;
    if (op < OP_EQ || op > OP_GEQ) throw CudaError("filter: comparison operator out of range for a fixed-width column");
    const unsigned grid = grid_for((n + 31) / 32, kFilterThreads / 32, sm_count);
    if (type == RJ_INT32) {
        // the literal is narrowed exactly like the reference does (statement.cpp:55: static_cast<int32_t>)
        filter_compare_kernel<int32_t><<<grid, kFilterThreads, 0, s>>>(static_cast<const int32_t*>(values), valid, n, op, static_cast<int32_t>(rhs_i), out);
    } else if (type == RJ_INT64) {
        filter_compare_kernel<int64_t><<<grid, kFilterThreads, 0, s>>>(static_cast<const int64_t*>(values), valid, n, op, rhs_i, out);
    } else if (type == RJ_FP64) {
        filter_compare_kernel<double><<<grid, kFilterThreads, 0, s>>>(static_cast<const double*>(values), valid, n, op, rhs_d, out);
    } else {
        throw CudaError("filter: not a fixed-width type");
    }
    RJ_LAUNCH_CHECK();
}

void launch_filter_null(const uint32_t* valid, uint64_t n, bool want_null, uint32_t* out, int sm_count, cudaStream_t s) {
    if (n == 0) return;
    filter_null_kernel<<<grid_for((n + 31) / 32, kFilterThreads, sm_count), kFilterThreads, 0, s>>>(valid, n, want_null ? 1 : 0, out);
    RJ_LAUNCH_CHECK();
}

void launch_bitmap_logic(const uint32_t* a, const uint32_t* b, uint64_t n, int op, uint32_t* out, int sm_count, cudaStream_t s) {
    if (n == 0) return;
    if (op < 0 || op > 2) throw CudaError("filter: logical operator out of range");
    bitmap_logic_kernel<<<grid_for((n + 31) / 32, kFilterThreads, sm_count), kFilterThreads, 0, s>>>(a, b ? b : a, n, op, out);
    RJ_LAUNCH_CHECK();
}

void launch_filter_varchar(const void* pages, const uint64_t* desc, const uint32_t* valid, uint64_t n, int op, const uint8_t* d_rhs,
                           uint32_t rhs_len, uint32_t* out, int sm_count, cudaStream_t s) {
    if (n == 0) return;
    if (op < OP_EQ || op > OP_NOT_LIKE) throw CudaError("filter: comparison operator out of range for a VARCHAR column");
    filter_varchar_kernel<<<grid_for((n + 31) / 32, kFilterThreads / 32, sm_count), kFilterThreads, 0, s>>>(
        static_cast<const uint8_t*>(pages), desc, valid, n, op, d_rhs, rhs_len, out);
    RJ_LAUNCH_CHECK();
}

void launch_bitmap_popc(const uint32_t* bits, uint64_t n_words, uint32_t* counts, int sm_count, cudaStream_t s) {
    if (n_words == 0) return;
    bitmap_popc_kernel<<<grid_for(n_words, kFilterThreads, sm_count), kFilterThreads, 0, s>>>(bits, n_words, counts);
    RJ_LAUNCH_CHECK();
}

void launch_bitmap_expand(const uint32_t* bits, const uint64_t* start, uint64_t n_words, uint32_t* ids, int sm_count, cudaStream_t s) {
    if (n_words == 0) return;
    bitmap_expand_kernel<<<grid_for(n_words, kFilterThreads / 32, sm_count), kFilterThreads, 0, s>>>(bits, start, n_words, ids);
    RJ_LAUNCH_CHECK();
}

} // namespace rj

// ---- layout of the multi-GPU pull exchange, from the all-gathered histograms (one small kernel instead of ~60
//      tensor operations; see radix_join_b200/dist_join.py: pull_layout, which this mirrors) ------------------
namespace rj {
namespace {

constexpr int kLayoutThreads = 1024;

// one block per side (blockIdx.x: 0 build, 1 probe).  H[q][side][f]: tuples of rank q in final partition f.
__global__ void __launch_bounds__(kLayoutThreads)
    dist_layout_kernel(const uint32_t* __restrict__ H, int me, int g, int bits, int p1, const uint64_t* __restrict__ ptrs /* [2][5][8] */,
                       const int* __restrict__ widths /* [2][5] */, uint32_t* __restrict__ cursor /* [2][ndig] */,
                       uint64_t* __restrict__ table /* [2][ndig][5] */, uint32_t* __restrict__ start /* [2][ndig+1] */,
                       uint32_t* __restrict__ tile /* [2][ndig+1] */, uint32_t* __restrict__ group /* [2][ndig] */,
                       uint32_t* __restrict__ local_hist /* [2][nfin >> g] */, unsigned long long* __restrict__ scalars /* [2][2]: owned, sent */) {
    __shared__ uint32_t s_cnt[8 * 256];   // [q][digit]
    __shared__ uint32_t s_run[8 * 256];   // sender-local start of every digit run
    __shared__ uint32_t s_sub_start[257], s_sub_tile[257];
    __shared__ unsigned long long s_owned;
    const int side = blockIdx.x, G = 1 << g, tid = threadIdx.x;
    const uint32_t nfin = 1u << bits, ndig = 1u << p1, per = ndig >> g, fpd = nfin / ndig;
    if (tid == 0) s_owned = 0;
    // 1. tuples per (sender, pass-1 digit)
    for (uint32_t i = tid; i < static_cast<uint32_t>(G) * ndig; i += kLayoutThreads) {
        const uint32_t q = i / ndig, d = i % ndig;
        const uint32_t* h = H + (static_cast<size_t>(q) * 2 + side) * nfin + static_cast<size_t>(d) * fpd;
        uint32_t sum = 0;
        for (uint32_t f = 0; f < fpd; ++f) sum += h[f];
        s_cnt[q * 256 + d] = sum;
    }
    __syncthreads();
    // 2. every sender's run starts (exclusive prefix over its digits); 3. the sub-regions this rank owns
    if (tid < G) {
        uint32_t run = 0;
        for (uint32_t d = 0; d < ndig; ++d) {
            s_run[tid * 256 + d] = run;
            run += s_cnt[tid * 256 + d];
        }
    }
    if (tid == 32) { // another warp: sub-region x = j * G + q
        uint32_t pos = 0, tiles = 0;
        for (uint32_t x = 0; x < per * G; ++x) {
            const uint32_t c = s_cnt[(x % G) * 256 + me * per + x / G];
            s_sub_start[x] = pos;
            s_sub_tile[x] = tiles;
            pos += c;
            tiles += (c + 4095u) / 4096u;
        }
        s_sub_start[per * G] = pos;
        s_sub_tile[per * G] = tiles;
    }
    __syncthreads();
    for (uint32_t d = tid; d < ndig; d += kLayoutThreads) cursor[side * ndig + d] = s_run[me * 256 + d];
    const uint32_t n_sub = per * G;
    for (uint32_t x = tid; x <= n_sub; x += kLayoutThreads) {
        start[side * (ndig + 1) + x] = s_sub_start[x];
        tile[side * (ndig + 1) + x] = s_sub_tile[x];
        if (x < n_sub) {
            const uint32_t q = x % G, d = me * per + x / G;
            group[side * ndig + x] = x / G;
            const long long delta = static_cast<long long>(s_run[q * 256 + d]) - static_cast<long long>(s_sub_start[x]);
#pragma unroll
            for (int a = 0; a < 5; ++a)
                table[(static_cast<size_t>(side) * ndig + x) * 5 + a] =
                    ptrs[(side * 5 + a) * 8 + q] + static_cast<uint64_t>(delta * widths[side * 5 + a]);
        }
    }
    // 4. tuples per final partition of the range this rank owns; what it owns / sends in total
    const uint32_t nloc = nfin >> g;
    unsigned long long mine = 0;
    for (uint32_t f = tid; f < nloc; f += kLayoutThreads) {
        uint32_t sum = 0;
        for (int q = 0; q < G; ++q) sum += H[(static_cast<size_t>(q) * 2 + side) * nfin + static_cast<size_t>(me) * nloc + f];
        local_hist[side * nloc + f] = sum;
        mine += sum;
    }
    atomicAdd(&s_owned, mine);
    __syncthreads();
    if (tid == 0) {
        unsigned long long sent = 0;
        for (uint32_t d = 0; d < ndig; ++d)
            if (d / per != static_cast<uint32_t>(me)) sent += s_cnt[me * 256 + d];
        scalars[side * 2 + 0] = s_owned;
        scalars[side * 2 + 1] = sent;
    }
}

} // namespace

void launch_dist_layout(const uint32_t* H, int me, int g, int bits, int p1, const uint64_t* ptrs, const int* widths, uint32_t* cursor,
                        uint64_t* table, uint32_t* start, uint32_t* tile, uint32_t* group, uint32_t* local_hist,
                        unsigned long long* scalars, cudaStream_t s) {
    if (g < 1 || g > 3 || p1 > 8 || p1 <= g || bits < p1 || bits > 15) throw CudaError("dist_layout: unsupported geometry");
    dist_layout_kernel<<<2, kLayoutThreads, 0, s>>>(H, me, g, bits, p1, ptrs, widths, cursor, table, start, tile, group, local_hist, scalars);
    RJ_LAUNCH_CHECK();
}

} // namespace rj
