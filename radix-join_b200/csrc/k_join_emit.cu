// Root join fused with page output: build + probe in shared memory, result PAGES written straight from
// the join kernel.
//
// Replaces, for the root of the plan, hash_join_omp's per-bucket build / probe / row emission
// (reference src/execute.cpp:196-261) together with Table::to_columnar of the result
// (src/build_table.cpp:456-594).  The general path (k_join.cu + k_gather_encode.cu) emits (build, probe)
// position pairs and gathers every output column through them: on config 2 that is 1.6 G gathers of 4-8
// bytes, each its own 128-byte L1 wavefront, and the LSU -- not DRAM -- bounds the encode.  Here both
// sides arrive FULLY partitioned with their output columns travelling beside the keys (scatter passes 1
// and 2 carry them), so a work unit reads its probe tuples and their payload sequentially, looks the
// build payload up in shared memory, and appends finished rows to a shared-memory image of the result
// pages.  No pair list, no gather, no separate encode pass.
//
// Row alignment across columns (include/plan.h:102-105: columns are row-aligned by cumulative row index,
// page boundaries are free) is kept by emitting CHUNKS of kChunkRows = 1984 rows: one page of every 4-byte
// column (1984 rows, the engine's fixed fill) and two pages of 992 rows of every 8-byte column.  A CTA
// reserves chunk c with one global atomic and owns pages c / 2c, 2c+1 of every column, so the columns'
// page lists enumerate the same rows in the same order.  (8-byte pages hold 992 instead of 1007 rows:
// 1.5 % more pages, the price of page-aligned chunks.)
//
// Within a chunk, rows are staged UNCOMPACTED at their row slot; when the chunk closes, every nullable
// page is compacted in place (values of non-NULL rows move to the front, in row order: registers are
// the double buffer, a block-wide prefix sum gives the slots) and gets its bitmap and header; then one
// thread writes the pages with 1-D TMA bulk stores.
//
// Build keys must be unique inside every table (every key / foreign-key join): the 64-bit CAS insert sees
// an equal key for free, raises a global flag and the whole launch is abandoned -- the engine then runs
// the general path, which handles duplicates with chains.
#include "rj_common.cuh"
#include "rj_internal.h"

namespace rj {
namespace {

constexpr int      kThreads   = 512;
constexpr int      kWarps     = kThreads / 32;
constexpr uint32_t kSlots     = kEmitSlots;       // 4096 x 64-bit (key | local build index << 32)
constexpr uint32_t kSlotMask  = kSlots - 1;
constexpr uint32_t kCap       = kEmitBuildCap;    // 3072 build tuples per table (75 % fill)
constexpr int      kBuildItems = kCap / kThreads; // 6
constexpr int      kItems     = 4;                // probe tuples per thread and batch
constexpr uint32_t kBatch     = kItems * kThreads;
constexpr uint32_t kChunkRows = kEmitChunkRows;   // 1984
constexpr uint32_t kHalfRows  = kChunkRows / 2;   // 992 rows per 8-byte page
constexpr uint32_t kNone      = 0xffffffffu;

__device__ __forceinline__ uint32_t probe_step(uint32_t k) { return ((k * 0x9E3779B1u) >> 20) | 1u; }

struct EmitArgs {
    const uint32_t* bkeys;
    const uint32_t* pkeys;
    const uint32_t* off_b;
    const uint32_t* off_p;
    const uint32_t* unit_start;
    uint32_t*       unit_cursor;
    uint32_t        nparts;
    int             part_bits;
    // carried columns, in final partition order beside the keys
    int            n_bpay, n_ppay;
    const void*    bpay[kEmitMaxPay];
    const uint8_t* bvalid[kEmitMaxPay]; // one byte per tuple, NULL = column has no NULL
    int            bwidth[kEmitMaxPay];
    const void*    ppay[kEmitMaxPay];
    const uint8_t* pvalid[kEmitMaxPay];
    int            pwidth[kEmitMaxPay];
    // output columns
    int      n_out;
    int      out_src[kEmitMaxOut];   // 0 = join key, 1 = build payload, 2 = probe payload
    int      out_idx[kEmitMaxOut];
    int      out_width[kEmitMaxOut]; // 4 or 8
    int      out_nullable[kEmitMaxOut];
    uint8_t* out_pages[kEmitMaxOut];
    // shared-memory layout (byte offsets into the dynamic segment, computed by the launcher)
    uint32_t sm_bpay[kEmitMaxPay], sm_bvalid[kEmitMaxPay];
    uint32_t sm_page[kEmitMaxOut];  // first page image of the column (8-byte columns: two consecutive images)
    uint32_t sm_valid[kEmitMaxOut]; // validity byte per staged row (nullable columns)
    // results
    uint32_t*           chunk_counter;
    unsigned long long* row_counter;
    uint32_t*           abort_flag; // set when a table meets a duplicate build key
};

// exclusive prefix of a packed 64-bit count vector over the CTA; *total = sum over all threads
__device__ __forceinline__ uint64_t block_scan_u64(uint64_t v, uint64_t* s_warp /* [kWarps] */, uint64_t* total) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint64_t o = __shfl_up_sync(RJ_FULL_MASK, inc, d);
        if (lane >= d) inc += o;
    }
    __syncthreads(); // s_warp may still be read by the previous scan
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint64_t before = 0, all = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
        const uint64_t s = s_warp[w];
        if (w < static_cast<int>(warp)) before += s;
        all += s;
    }
    *total = all;
    return before + inc - v;
}

// Close one page image: `n` staged rows of width W sit uncompacted at their row slots.  Nullable pages are
// compacted in place; every page gets its bitmap (last ceil(n/8) bytes) and header (n_r @0, n_v @2).
// Called by the whole CTA.
template <typename T>
__device__ __forceinline__ void close_page(uint8_t* pg, uint32_t n, const uint8_t* valid /* NULL: no NULLs */, uint64_t* s_warp) {
    constexpr uint32_t kBegin = sizeof(T) == 4 ? 4 : 8;
    constexpr int      kPer   = sizeof(T) == 4 ? (kChunkRows + kThreads - 1) / kThreads : (kHalfRows + kThreads - 1) / kThreads; // 4 or 2
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    T*             vals = reinterpret_cast<T*>(pg + kBegin);
    uint8_t*       bm   = pg + RJ_PAGE - ((n + 7) >> 3);
    const uint32_t bm_bytes = (n + 7) >> 3;
    uint32_t n_v = n;
    if (valid != nullptr) {
        // thread t owns rows t, t + 512, ...: a warp's 32 rows of one segment are consecutive, so a ballot
        // is a bitmap word; counts of the segments ride in 16-bit fields of one scan
        T        v[kPer];
        bool     ok[kPer];
        uint64_t cnt = 0;
#pragma unroll
        for (int q = 0; q < kPer; ++q) {
            const uint32_t r = q * kThreads + tid;
            ok[q] = r < n && valid[r] != 0;
            v[q]  = ok[q] ? vals[r] : T(0);
            cnt |= static_cast<uint64_t>(ok[q] ? 1u : 0u) << (16 * q);
        }
        uint64_t total;
        const uint64_t excl = block_scan_u64(cnt, s_warp, &total); // its barriers also order the reads above before the writes below
        uint32_t base = 0;
#pragma unroll
        for (int q = 0; q < kPer; ++q) {
            if (ok[q]) vals[base + static_cast<uint32_t>((excl >> (16 * q)) & 0xffffu)] = v[q];
            base += static_cast<uint32_t>((total >> (16 * q)) & 0xffffu);
            const uint32_t word = __ballot_sync(RJ_FULL_MASK, ok[q]);
            const uint32_t b0   = (q * kThreads + warp * 32) >> 3; // first bitmap byte of this warp's 32 rows
            if (lane < 4 && b0 + lane < bm_bytes) bm[b0 + lane] = static_cast<uint8_t>(word >> (8 * lane));
        }
        n_v = base;
    } else {
        // no NULLs: the staged order is the page order; the bitmap is all ones up to row n
        for (uint32_t b = tid; b < bm_bytes; b += kThreads) bm[b] = (b == (n >> 3)) ? static_cast<uint8_t>((1u << (n & 7)) - 1u) : 0xffu;
    }
    if (tid == 0) *reinterpret_cast<uint32_t*>(pg) = n | (n_v << 16);
}

template <int NP>
__global__ void __launch_bounds__(kThreads, 2) join_emit_kernel(const EmitArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    unsigned long long* slots = reinterpret_cast<unsigned long long*>(smem); // table first: 32 KB
    __shared__ uint64_t s_warp[kWarps];
    __shared__ uint32_t s_rows;   // rows reserved in the open chunk (may overshoot kChunkRows)
    __shared__ uint32_t s_unit, s_part, s_chunk;

    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t lt = lanemask_lt();
    const uint32_t n_units = a.unit_start[a.nparts];
    const int      part_bits = a.part_bits;
    if (tid == 0) s_rows = 0;

    // write the open chunk (n rows) as pages and start a new one; called by the whole CTA
    auto flush = [&](uint32_t n) {
        if (tid == 0) {
            s_chunk = atomicAdd(a.chunk_counter, 1u);
            atomicAdd(a.row_counter, static_cast<unsigned long long>(n));
        }
#pragma unroll
        for (int j = 0; j < kEmitMaxOut; ++j) {
            if (j < a.n_out) {
                uint8_t*       pg = smem + a.sm_page[j];
                const uint8_t* vb = a.out_nullable[j] ? smem + a.sm_valid[j] : nullptr;
                if (a.out_width[j] == 4) {
                    close_page<uint32_t>(pg, n, vb, s_warp);
                } else {
                    const uint32_t n0 = n < kHalfRows ? n : kHalfRows;
                    close_page<uint64_t>(pg, n0, vb, s_warp);
                    close_page<uint64_t>(pg + RJ_PAGE, n - n0, vb ? vb + kHalfRows : nullptr, s_warp);
                }
            }
        }
        fence_proxy_async_smem(); // generic-proxy writes -> async proxy, by every writer
        __syncthreads();
        if (tid == 0) {
            const uint64_t c = s_chunk;
#pragma unroll
            for (int j = 0; j < kEmitMaxOut; ++j) {
                if (j < a.n_out) {
                    if (a.out_width[j] == 4) {
                        tma_store_1d(a.out_pages[j] + c * RJ_PAGE, smem + a.sm_page[j], RJ_PAGE);
                    } else {
                        tma_store_1d(a.out_pages[j] + 2 * c * RJ_PAGE, smem + a.sm_page[j], 2 * RJ_PAGE);
                    }
                }
            }
            tma_store_commit();
            tma_store_wait_read<0>(); // the images are reused right away
            s_rows = 0;
        }
        __syncthreads();
    };

    for (;;) {
        // ---- next work unit, in global order (see k_join.cu) -------------------------------------------
        __syncthreads();
        if (tid < 32) {
            uint32_t u = 0;
            if (lane == 0) {
                u = atomicAdd(a.unit_cursor, 1u);
                if (*reinterpret_cast<volatile uint32_t*>(a.abort_flag)) u = 0xffffffffu; // somebody met a duplicate key
            }
            u = __shfl_sync(RJ_FULL_MASK, u, 0);
            uint32_t lo = 0, hi = a.nparts;
            if (u < n_units) {
                while (hi - lo > 1) {
                    const uint32_t span = hi - lo;
                    const uint32_t step = (span + 31) / 32;
                    const uint32_t probe = lo + (lane + 1) * step;
                    const bool     le = probe < hi && a.unit_start[probe] <= u;
                    const uint32_t k = __popc(__ballot_sync(RJ_FULL_MASK, le));
                    const uint32_t nlo = lo + k * step;
                    const uint32_t nhi = (k < 32 && lo + (k + 1) * step < hi) ? lo + (k + 1) * step : hi;
                    lo = nlo;
                    hi = nhi;
                }
            }
            if (lane == 0) {
                s_unit = u;
                s_part = lo;
            }
        }
        __syncthreads();
        const uint32_t u = s_unit;
        if (u >= n_units) break;
        const uint32_t part = s_part;
        const uint32_t local = u - a.unit_start[part];
        const uint32_t b_lo = a.off_b[part], b_hi = a.off_b[part + 1];
        const uint32_t p_lo = a.off_p[part], p_hi = a.off_p[part + 1];
        const uint32_t n_pchunks = (p_hi - p_lo + kJoinProbeChunk - 1) / kJoinProbeChunk;
        const uint32_t bc = local / n_pchunks, pc = local - bc * n_pchunks;
        const uint32_t bs = b_lo + bc * kCap;
        const uint32_t nb = (b_hi - bs > kCap) ? kCap : b_hi - bs;
        const uint32_t ps = p_lo + pc * kJoinProbeChunk;
        const uint32_t pe = (p_hi - ps > kJoinProbeChunk) ? ps + kJoinProbeChunk : p_hi;

        // ---- build: table of (key | index inside the chunk), the chunk's payload beside it ---------------
        uint32_t bkey[kBuildItems];
#pragma unroll
        for (int k = 0; k < kBuildItems; ++k) {
            const uint32_t i = k * kThreads + tid;
            bkey[k] = i < nb ? a.bkeys[bs + i] : 0u;
        }
        // (the barrier at the top of the loop ended every read of the previous table and payload)
        for (uint32_t s = tid; s < kSlots; s += kThreads) slots[s] = ~0ull;
#pragma unroll
        for (int c = 0; c < kEmitMaxPay; ++c) {
            if (c < a.n_bpay) {
                if (a.bwidth[c] == 8) {
                    uint64_t*       dst = reinterpret_cast<uint64_t*>(smem + a.sm_bpay[c]);
                    const uint64_t* src = static_cast<const uint64_t*>(a.bpay[c]) + bs;
                    for (uint32_t i = tid; i < nb; i += kThreads) dst[i] = src[i];
                } else {
                    uint32_t*       dst = reinterpret_cast<uint32_t*>(smem + a.sm_bpay[c]);
                    const uint32_t* src = static_cast<const uint32_t*>(a.bpay[c]) + bs;
                    for (uint32_t i = tid; i < nb; i += kThreads) dst[i] = src[i];
                }
                if (a.bvalid[c] != nullptr) {
                    uint8_t*       dst = smem + a.sm_bvalid[c];
                    const uint8_t* src = a.bvalid[c] + bs;
                    for (uint32_t i = tid; i < nb; i += kThreads) dst[i] = src[i];
                }
            }
        }
        __syncthreads();
        bool dup = false;
#pragma unroll
        for (int k = 0; k < kBuildItems; ++k) {
            const uint32_t i = k * kThreads + tid;
            if (i < nb) {
                const uint32_t           key  = bkey[k];
                const unsigned long long mine = static_cast<unsigned long long>(key) | (static_cast<unsigned long long>(i) << 32);
                uint32_t       sl   = (hash_key(key) >> part_bits) & kSlotMask;
                const uint32_t step = probe_step(key);
                for (;;) {
                    unsigned long long cur = slots[sl];
                    if (cur == ~0ull) cur = atomicCAS(&slots[sl], ~0ull, mine);
                    if (cur == ~0ull) break;
                    if (static_cast<uint32_t>(cur) == key) {
                        dup = true;
                        break;
                    }
                    sl = (sl + step) & kSlotMask;
                }
            }
            __syncwarp();
        }
        if (__syncthreads_or(dup ? 1 : 0)) {
            // not a key / foreign-key join: leave it to the general path
            if (tid == 0) atomicExch(a.abort_flag, 1u);
            return;
        }

        // ---- probe -----------------------------------------------------------------------------------------
        for (uint32_t base = ps; base < pe; base += kBatch) {
            uint32_t key[kItems], lidx[kItems];
            uint64_t pval[NP > 0 ? NP : 1][kItems];
            uint32_t pok = 0; // bit (c * kItems + k): probe payload c of item k is not NULL
#pragma unroll
            for (int k = 0; k < kItems; ++k) {
                const uint32_t i = base + k * kThreads + tid;
                key[k]  = i < pe ? a.pkeys[i] : 0u;
                lidx[k] = i < pe ? 0u : kNone; // kNone = no tuple here
            }
#pragma unroll
            for (int c = 0; c < NP; ++c) {
#pragma unroll
                for (int k = 0; k < kItems; ++k) {
                    const uint32_t i = base + k * kThreads + tid;
                    pval[c][k] = 0;
                    if (i < pe) {
                        pval[c][k] = a.pwidth[c] == 8 ? static_cast<const uint64_t*>(a.ppay[c])[i]
                                                      : static_cast<uint64_t>(static_cast<const uint32_t*>(a.ppay[c])[i]);
                        if (a.pvalid[c] == nullptr || a.pvalid[c][i] != 0) pok |= 1u << (c * kItems + k);
                    }
                }
            }
            // look every tuple up: at most one match, the table holds distinct keys
            uint32_t pending = 0;
#pragma unroll
            for (int k = 0; k < kItems; ++k) {
                if (lidx[k] != kNone) {
                    uint32_t       sl   = (hash_key(key[k]) >> part_bits) & kSlotMask;
                    const uint32_t step = probe_step(key[k]);
                    lidx[k] = kNone;
                    for (;;) {
                        const unsigned long long e = slots[sl];
                        if (e == ~0ull) break;
                        if (static_cast<uint32_t>(e) == key[k]) {
                            lidx[k] = static_cast<uint32_t>(e >> 32);
                            pending |= 1u << k;
                            break;
                        }
                        sl = (sl + step) & kSlotMask;
                    }
                }
                __syncwarp();
            }
            // stage the matches as rows of the open chunk; rows that do not fit wait for the next chunk
            for (;;) {
                uint32_t bal[kItems], total = 0;
#pragma unroll
                for (int k = 0; k < kItems; ++k) {
                    bal[k] = __ballot_sync(RJ_FULL_MASK, (pending >> k) & 1u);
                    total += __popc(bal[k]);
                }
                uint32_t off = 0;
                if (lane == 0 && total) off = atomicAdd(&s_rows, total);
                off = __shfl_sync(RJ_FULL_MASK, off, 0);
#pragma unroll
                for (int k = 0; k < kItems; ++k) {
                    const uint32_t pos = off + __popc(bal[k] & lt);
                    off += __popc(bal[k]);
                    if (((pending >> k) & 1u) && pos < kChunkRows) {
                        pending &= ~(1u << k);
#pragma unroll
                        for (int j = 0; j < kEmitMaxOut; ++j) {
                            if (j < a.n_out) {
                                uint64_t v  = key[k];
                                bool     ok = true;
                                const int c = a.out_idx[j];
                                if (a.out_src[j] == 1) {
                                    v = a.bwidth[c] == 8 ? reinterpret_cast<const uint64_t*>(smem + a.sm_bpay[c])[lidx[k]]
                                                         : static_cast<uint64_t>(reinterpret_cast<const uint32_t*>(smem + a.sm_bpay[c])[lidx[k]]);
                                    if (a.bvalid[c] != nullptr) ok = (smem + a.sm_bvalid[c])[lidx[k]] != 0;
                                } else if (a.out_src[j] == 2) {
                                    if (NP > 0) {
                                        // c is 0 or 1: select without dynamic register indexing
                                        v  = (NP > 1 && c == 1) ? pval[NP > 1 ? 1 : 0][k] : pval[0][k];
                                        ok = (pok >> (c * kItems + k)) & 1u;
                                    }
                                }
                                uint8_t* pg = smem + a.sm_page[j];
                                if (a.out_width[j] == 4) {
                                    reinterpret_cast<uint32_t*>(pg + 4)[pos] = static_cast<uint32_t>(v);
                                } else {
                                    const uint32_t h = pos >= kHalfRows ? 1u : 0u;
                                    reinterpret_cast<uint64_t*>(pg + h * RJ_PAGE + 8)[pos - h * kHalfRows] = v;
                                }
                                if (a.out_nullable[j]) (smem + a.sm_valid[j])[pos] = ok ? 1 : 0;
                            }
                        }
                    }
                }
                __syncthreads(); // every reservation of this round is in s_rows, every row in shared memory
                const uint32_t reserved = s_rows;
                // (second barrier: nobody adds to s_rows again before everybody has read it -- the decision
                // to flush must be the same in every thread)
                const int waiting = __syncthreads_or(pending != 0 ? 1 : 0);
                if (reserved >= kChunkRows) flush(kChunkRows); // resets s_rows behind a barrier
                if (!waiting) break;
            }
        }
    }
    // the partly filled last chunk of this CTA
    __syncthreads();
    const uint32_t left = s_rows;
    if (left > 0) flush(left < kChunkRows ? left : kChunkRows);
}

} // namespace

size_t join_emit_smem(const JoinEmitLaunch& L, uint32_t* sm_bpay, uint32_t* sm_bvalid, uint32_t* sm_page, uint32_t* sm_valid) {
    size_t off = sizeof(uint64_t) * kEmitSlots;
    auto   take = [&](size_t bytes, size_t align) {
        off = (off + align - 1) / align * align;
        const size_t at = off;
        off += bytes;
        return static_cast<uint32_t>(at);
    };
    for (int j = 0; j < L.n_out; ++j) sm_page[j] = take(L.out_width[j] == 4 ? RJ_PAGE : 2 * RJ_PAGE, 128);
    for (int c = 0; c < L.n_bpay; ++c) sm_bpay[c] = take(size_t(kEmitBuildCap) * L.bwidth[c], 16);
    for (int c = 0; c < L.n_bpay; ++c) sm_bvalid[c] = L.bvalid[c] ? take(kEmitBuildCap, 16) : 0;
    for (int j = 0; j < L.n_out; ++j) sm_valid[j] = L.out_nullable[j] ? take(kEmitChunkRows, 16) : 0;
    return off;
}

bool join_emit_fits(const JoinEmitLaunch& L) {
    uint32_t a[kEmitMaxPay], b[kEmitMaxPay], c[kEmitMaxOut], d[kEmitMaxOut];
    if (L.n_out < 1 || L.n_out > kEmitMaxOut || L.n_bpay > kEmitMaxPay || L.n_ppay > kEmitMaxPay) return false;
    return join_emit_smem(L, a, b, c, d) <= 112 * 1024; // two CTAs per SM
}

void launch_join_emit(const JoinEmitLaunch& L, int sm_count, cudaStream_t s) {
    EmitArgs a{};
    a.bkeys = L.bkeys; a.pkeys = L.pkeys; a.off_b = L.off_b; a.off_p = L.off_p;
    a.unit_start = L.unit_start; a.unit_cursor = L.unit_cursor; a.nparts = L.nparts; a.part_bits = L.part_bits;
    a.n_bpay = L.n_bpay; a.n_ppay = L.n_ppay; a.n_out = L.n_out;
    for (int c = 0; c < kEmitMaxPay; ++c) {
        a.bpay[c] = L.bpay[c]; a.bvalid[c] = L.bvalid[c]; a.bwidth[c] = L.bwidth[c];
        a.ppay[c] = L.ppay[c]; a.pvalid[c] = L.pvalid[c]; a.pwidth[c] = L.pwidth[c];
    }
    for (int j = 0; j < kEmitMaxOut; ++j) {
        a.out_src[j] = L.out_src[j]; a.out_idx[j] = L.out_idx[j]; a.out_width[j] = L.out_width[j];
        a.out_nullable[j] = L.out_nullable[j]; a.out_pages[j] = L.out_pages[j];
    }
    a.chunk_counter = L.chunk_counter; a.row_counter = L.row_counter; a.abort_flag = L.abort_flag;
    const size_t smem = join_emit_smem(L, a.sm_bpay, a.sm_bvalid, a.sm_page, a.sm_valid);
    if (smem > 112 * 1024) throw CudaError("join_emit: the columns do not fit shared memory");
    const unsigned grid = join_emit_grid(sm_count);
    switch (L.n_ppay) {
    case 0: {
        static SmemConfigured cfg;
        cfg.ensure(join_emit_kernel<0>, smem);
        join_emit_kernel<0><<<grid, kThreads, smem, s>>>(a);
        break;
    }
    case 1: {
        static SmemConfigured cfg;
        cfg.ensure(join_emit_kernel<1>, smem);
        join_emit_kernel<1><<<grid, kThreads, smem, s>>>(a);
        break;
    }
    default: {
        static SmemConfigured cfg;
        cfg.ensure(join_emit_kernel<2>, smem);
        join_emit_kernel<2><<<grid, kThreads, smem, s>>>(a);
        break;
    }
    }
    RJ_LAUNCH_CHECK();
}

} // namespace rj
