// Root join fused with page output: build + probe in shared memory, result PAGES written straight from
// the join kernel.
//
// Replaces, for the root of the plan, hash_join_omp's per-bucket build / probe / row emission
// (reference src/execute.cpp:196-261) together with Table::to_columnar of the result
// (src/build_table.cpp:456-594).  The general path (k_join.cu + k_gather_encode.cu) emits (build, probe)
// position pairs and gathers every output column through them: on config 2 that is 1.6 G gathers of 4-8
// bytes, each its own 128-byte L1 wavefront, and the LSU -- not DRAM -- bounds the encode.  Here both
// sides arrive FULLY partitioned with their output columns travelling beside the keys
// (k_scatter_carry.cu), so a work unit reads its probe tuples and their columns sequentially, looks the
// build columns up in shared memory, and stores finished rows into the result pages.  No pair list, no
// gather, no separate encode pass.
//
// Memory pipeline.  A unit's probe tuples are consumed in batches of 2048; the keys, carried values and
// validity bytes of a batch arrive by 1-D TMA bulk copies into one of two shared-memory buffers, and
// batch i+1 is requested before batch i is probed, so the loads of a CTA are in flight all the time and
// no thread waits on a global load in the probe loop.  The build side's carried columns arrive the same
// way while the hash table is being filled.
//
// Result pages.  Row alignment across columns (include/plan.h:102-105: columns are row-aligned by
// cumulative row index, page boundaries are free) is kept by emitting CHUNKS of 1984 rows: one page of
// every 4-byte column (1984 rows, the engine's fixed fill) and two pages of 992 rows of every 8-byte
// column.  A CTA reserves chunk c with one global atomic and owns pages c / 2c, 2c+1 of every column, so
// the columns' page lists enumerate the same rows in the same order.  (8-byte pages hold 992 instead of
// 1007 rows: 1.5 % more pages, the price of page-aligned chunks.)
//
// A page stores only its non-NULL values, packed (src/build_table.cpp:484-501).  Rows are therefore placed
// by ONE packed shared-memory atomic per warp and batch that hands out, together, the warp's row range and
// its range of value slots in every nullable column (12-bit fields: rows, then one count per nullable
// carried column).  Because all ranges come from the same atomic they are ordered alike, so the k-th
// non-NULL row of a page owns its k-th value slot and the values go straight from registers to their final
// place in global memory -- consecutive rows of a warp store to consecutive addresses.  A CTA holds TWO
// chunks open (positions 0..1983 = chunk A, 1984..3967 = chunk B), so a batch of up to 2048 matches is
// always placed in one round; the warps whose rows contain a page boundary (rows 991, 1983, 2975, 3967)
// publish the value counts up to it, which is all a later row needs to find its slot in its own page.
// Validity bitmaps are OR-ed together in shared memory (one run of ones per warp and item, REDUX when a
// NULL is among them) and written with the page headers when chunk A closes; B then becomes A.
//
// Build keys must be unique inside every table (every key / foreign-key join): the 64-bit CAS insert sees
// an equal key for free, raises a global flag and the whole launch is abandoned -- the engine then runs
// the general path, which handles duplicates with chains.
#include "rj_common.cuh"
#include "rj_internal.h"

#include <type_traits>

namespace rj {
namespace {

constexpr int      kThreads    = 512;
constexpr int      kWarps      = kThreads / 32;
constexpr uint32_t kSlots      = kEmitSlots;       // 4096 x 64-bit (key | local build index << 32)
constexpr uint32_t kSlotMask   = kSlots - 1;
constexpr uint32_t kCap        = kEmitBuildCap;    // 3072 build tuples per table (75 % fill)
constexpr int      kBuildItems = kCap / kThreads;  // 6
constexpr int      kItems      = 4;                // probe tuples per thread and batch
constexpr uint32_t kBatch      = kItems * kThreads;
constexpr uint32_t kChunkRows  = kEmitChunkRows;   // 1984
constexpr uint32_t kHalfRows   = kChunkRows / 2;   // 992 rows per 8-byte page = 31 bitmap words
constexpr uint32_t kNone       = 0xffffffffu;
constexpr int      kMaxNull    = 2 * kEmitMaxPay;  // nullable carried columns
constexpr int      kFieldBits  = 12;               // rows reserved before a close stay below 1984 + 2048 < 4096
constexpr uint32_t kFieldMask  = (1u << kFieldBits) - 1;
constexpr uint32_t kBitmapWords = kChunkRows / 32; // 62 per chunk; a CTA keeps two chunks' worth
constexpr uint32_t kOpenRows    = 2 * kChunkRows;  // rows a round may place (chunks A and B)
static_assert(kMaxNull <= 4, "five 12-bit fields fit one 64-bit counter");

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
    uint32_t        probe_chunk; // probe tuples per work unit
    // carried columns, in final partition order beside the keys
    const void*    bpay[kEmitMaxPay];
    const uint8_t* bvalid[kEmitMaxPay]; // one byte per tuple, NULL = the column holds no NULL
    int            bwidth[kEmitMaxPay];
    int            bnull[kEmitMaxPay];  // index among the nullable carried columns, or -1
    const void*    ppay[kEmitMaxPay];
    const uint8_t* pvalid[kEmitMaxPay];
    int            pwidth[kEmitMaxPay];
    int            pnull[kEmitMaxPay];
    int            n_null;
    // output columns
    int      n_out;
    int      out_src[kEmitMaxOut];   // 0 = join key, 1 = build column, 2 = probe column
    int      out_idx[kEmitMaxOut];
    int      out_width[kEmitMaxOut]; // 4 or 8
    int      out_null[kEmitMaxOut];  // index among the nullable carried columns, or -1
    uint8_t* out_pages[kEmitMaxOut];
    // shared-memory layout (byte offsets into the dynamic segment, computed by the launcher)
    uint32_t sm_bpay[kEmitMaxPay], sm_bvalid[kEmitMaxPay];
    uint32_t sm_pkeys[2], sm_ppay[2][kEmitMaxPay], sm_pvalid[2][kEmitMaxPay];
    uint32_t sm_bitmap[kMaxNull];
    // results
    uint32_t*           chunk_counter;
    unsigned long long* row_counter;
    uint32_t*           abort_flag; // set when a table meets a duplicate build key
};

__device__ __forceinline__ uint32_t round16(uint32_t b) { return (b + 15u) & ~15u; }

template <int NB, int NP, int NN>
__global__ void __launch_bounds__(kThreads, 2) join_emit_kernel(const EmitArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    unsigned long long* slots = reinterpret_cast<unsigned long long*>(smem); // table first: 32 KB
    __shared__ unsigned long long s_ctr; // rows | values of nullable column 0 << 12 | ... of the open chunk
    __shared__ uint32_t s_unit, s_part;
    __shared__ uint32_t s_chunk[2];           // chunks A and B
    __shared__ uint32_t s_nvb[5][kMaxNull];   // [h]: non-NULL values of column nn among rows [0, 992 h) of the open chunks
    __shared__ __align__(8) uint64_t s_bbar, s_pbar[2];

    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t lt = lanemask_lt();
    const uint32_t n_units = a.unit_start[a.nparts];
    const int      part_bits = a.part_bits;
    constexpr int  n_null = NN;
    if (tid == 0) {
        s_ctr = 0;
        mbar_init(&s_bbar, 1);
        mbar_init(&s_pbar[0], 1);
        mbar_init(&s_pbar[1], 1);
        fence_mbar_init();
        s_chunk[0] = atomicAdd(a.chunk_counter, 1u); // a CTA always owns its open chunks A and B and one more ahead
        s_chunk[1] = atomicAdd(a.chunk_counter, 1u);
    }
    if (tid < 5 * kMaxNull) s_nvb[tid / kMaxNull][tid % kMaxNull] = 0u;
    for (uint32_t w = tid; w < kMaxNull * 2 * kBitmapWords; w += kThreads)
        if (w / (2 * kBitmapWords) < static_cast<uint32_t>(n_null)) reinterpret_cast<uint32_t*>(smem + a.sm_bitmap[w / (2 * kBitmapWords)])[w % (2 * kBitmapWords)] = 0u;
    uint32_t unit_no = 0, batch_no = 0; // mbarrier phases
    uint32_t chunk_ahead = 0;           // thread 0: the chunk that becomes B at the next close
    if (tid == 0) chunk_ahead = atomicAdd(a.chunk_counter, 1u);

    // one batch of probe tuples [base, base + cnt) into buffer s (one thread)
    auto issue_probe = [&](int s, uint32_t base, uint32_t cnt) {
        uint32_t bytes = round16((cnt + (base & 3u)) * 4u);
        uint32_t pb[kEmitMaxPay] = {}, vb[kEmitMaxPay] = {};
#pragma unroll
        for (int c = 0; c < NP; ++c) {
            pb[c] = round16((cnt + (base & (16u / a.pwidth[c] - 1u))) * a.pwidth[c]);
            vb[c] = a.pvalid[c] != nullptr ? round16(cnt + (base & 15u)) : 0u;
            bytes += pb[c] + vb[c];
        }
        mbar_arrive_expect_tx(&s_pbar[s], bytes);
        tma_load_1d(smem + a.sm_pkeys[s], a.pkeys + (base - (base & 3u)), round16((cnt + (base & 3u)) * 4u), &s_pbar[s]);
#pragma unroll
        for (int c = 0; c < NP; ++c) {
            const uint32_t sk = base & (16u / a.pwidth[c] - 1u);
            tma_load_1d(smem + a.sm_ppay[s][c], static_cast<const char*>(a.ppay[c]) + static_cast<uint64_t>(base - sk) * a.pwidth[c], pb[c], &s_pbar[s]);
            if (vb[c]) tma_load_1d(smem + a.sm_pvalid[s][c], a.pvalid[c] + (base - (base & 15u)), vb[c], &s_pbar[s]);
        }
    };

    // Pages of one chunk: bitmaps and headers (the values are already in place).  `n` rows; nullable
    // columns take their bitmap from words [0, 62) of the shared bitmaps and their value counts from
    // nv_first (rows < 992) / nv_all.  Threads 0..61 write full chunks word by word.
    constexpr uint32_t kMetaThreads = 64; // 62 bitmap words per chunk: two warps do all of it, the others move on
    auto write_chunk_meta = [&](uint64_t c, uint32_t n, const uint32_t* nv_first, const uint32_t* nv_all) {
        if (tid >= kMetaThreads) return;
#pragma unroll
        for (int j = 0; j < kEmitMaxOut; ++j) {
            if (j >= a.n_out) continue;
            const int       nn = a.out_null[j];
            const uint32_t* bm = nn >= 0 ? reinterpret_cast<const uint32_t*>(smem + a.sm_bitmap[nn]) : nullptr;
            const uint32_t  va = nn >= 0 ? nv_all[nn] : n;
            const uint32_t  vf = nn >= 0 ? nv_first[nn] : (n < kHalfRows ? n : kHalfRows);
            const int       pages = a.out_width[j] == 4 ? 1 : 2;
            for (int h = 0; h < pages; ++h) {
                uint8_t* pg = a.out_pages[j] + (pages * c + h) * RJ_PAGE;
                const uint32_t first = h * kHalfRows; // first row of the page inside the chunk
                const uint32_t rows = pages == 1 ? n : (h == 0 ? (n < kHalfRows ? n : kHalfRows) : (n > kHalfRows ? n - kHalfRows : 0u));
                const uint32_t bytes = (rows + 7) >> 3;
                uint8_t* dst = pg + RJ_PAGE - bytes;
                const uint32_t* bmh = bm ? bm + first / 32 : nullptr;
                if ((rows & 31u) == 0) { // full pages (1984 / 992 rows): the bitmap is word-aligned
                    for (uint32_t w = tid; w < rows / 32; w += kMetaThreads) reinterpret_cast<uint32_t*>(dst)[w] = bmh ? bmh[w] : 0xffffffffu;
                } else {
                    for (uint32_t b = tid; b < bytes; b += kMetaThreads) {
                        const uint32_t word = bmh ? bmh[b >> 2] : ((b >> 2) < (rows >> 5) ? 0xffffffffu : ((1u << (rows & 31u)) - 1u));
                        dst[b] = static_cast<uint8_t>(word >> (8 * (b & 3u)));
                    }
                }
                if (tid == 0) *reinterpret_cast<uint32_t*>(pg) = rows | ((pages == 1 ? va : (h == 0 ? vf : va - vf)) << 16);
            }
        }
    };

    // Chunk A is full: write its bitmaps and headers, then B becomes A.  Runs behind the barrier that ended
    // the round's stores into the shared bitmaps.  Only the first two warps work (62 bitmap words per
    // chunk); they synchronise among themselves with a named barrier, and the rest of the CTA meets them at
    // the barrier in front of the next reservation -- probing the next batch in the meantime.
    // `reserved` = rows handed out so far (<= 3968 placed).
    auto close_full_chunk = [&](uint32_t reserved) {
        if (tid >= kMetaThreads) return;
        asm volatile("bar.sync 1, 64;" ::: "memory"); // an earlier close of this round is complete
        uint32_t nv1[kMaxNull], nv2[kMaxNull];
#pragma unroll
        for (int nn = 0; nn < kMaxNull; ++nn) {
            nv1[nn] = nn < NN ? s_nvb[1][nn] : 0u;
            nv2[nn] = nn < NN ? s_nvb[2][nn] : 0u;
        }
        const uint32_t nv3 = tid < NN ? s_nvb[3][tid] : 0u, nv4 = tid < NN ? s_nvb[4][tid] : 0u;
        write_chunk_meta(s_chunk[0], kChunkRows, nv1, nv2);
        asm volatile("bar.sync 1, 64;" ::: "memory"); // both warps have read the bitmaps, s_nvb and s_chunk
        // B's words move down
        if (tid < kBitmapWords) {
#pragma unroll
            for (int nn = 0; nn < NN; ++nn) {
                uint32_t* bm = reinterpret_cast<uint32_t*>(smem + a.sm_bitmap[nn]);
                bm[tid] = bm[kBitmapWords + tid];
                bm[kBitmapWords + tid] = 0u;
            }
        }
        if (tid == 0) {
            atomicAdd(a.row_counter, static_cast<unsigned long long>(kChunkRows));
            // rows and value counts become relative to the new chunk A
            unsigned long long ctr = s_ctr;
            if (reserved > kOpenRows) { // rows beyond B were not placed: they ask again
                ctr = kOpenRows;
#pragma unroll
                for (int nn = 0; nn < NN; ++nn) ctr |= static_cast<unsigned long long>(s_nvb[4][nn]) << (kFieldBits * (nn + 1));
            }
            unsigned long long sub = kChunkRows;
#pragma unroll
            for (int nn = 0; nn < NN; ++nn) sub |= static_cast<unsigned long long>(nv2[nn]) << (kFieldBits * (nn + 1));
            s_ctr = ctr - sub;
            s_chunk[0] = s_chunk[1];
            s_chunk[1] = chunk_ahead;
            chunk_ahead = atomicAdd(a.chunk_counter, 1u); // consumed at the next close
        }
        if (tid < NN) {
            const uint32_t mine2 = s_nvb[2][tid];
            s_nvb[1][tid] = nv3 - mine2; // meaningful once a row beyond the boundary exists
            s_nvb[2][tid] = nv4 - mine2;
        }
    };

    bool     close_pending = false; // a chunk was closed since the last CTA-wide barrier
    uint32_t u_next = 0;
    if (tid == 0) u_next = atomicAdd(a.unit_cursor, 1u);
    for (;;) {
        // ---- next work unit, in global order (see k_join.cu); the cursor was advanced one unit ahead --------
        __syncthreads();
        close_pending = false;
        if (tid < 32) {
            uint32_t u = __shfl_sync(RJ_FULL_MASK, u_next, 0);
            if (lane == 0 && *reinterpret_cast<volatile uint32_t*>(a.abort_flag)) u = 0xffffffffu; // somebody met a duplicate key
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
                if (lane == 0) u_next = atomicAdd(a.unit_cursor, 1u); // returns while this unit is being processed
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
        const uint32_t n_pchunks = (p_hi - p_lo + a.probe_chunk - 1) / a.probe_chunk;
        const uint32_t bc = local / n_pchunks, pc = local - bc * n_pchunks;
        const uint32_t bs = b_lo + bc * kCap;
        const uint32_t nb = (b_hi - bs > kCap) ? kCap : b_hi - bs;
        const uint32_t ps = p_lo + pc * a.probe_chunk;
        const uint32_t pe = (p_hi - ps > a.probe_chunk) ? ps + a.probe_chunk : p_hi;

        // ---- build: the carried columns and the first probe batch are requested, then the table is filled -----
        if (tid == 0) {
            // (the barrier at the top of the loop ended every read of the previous unit's buffers)
            uint32_t bytes = 0, pb[kEmitMaxPay] = {}, vb[kEmitMaxPay] = {};
#pragma unroll
            for (int c = 0; c < NB; ++c) {
                pb[c] = round16((nb + (bs & (16u / a.bwidth[c] - 1u))) * a.bwidth[c]);
                vb[c] = a.bvalid[c] != nullptr ? round16(nb + (bs & 15u)) : 0u;
                bytes += pb[c] + vb[c];
            }
            if (NB > 0) {
                mbar_arrive_expect_tx(&s_bbar, bytes);
#pragma unroll
                for (int c = 0; c < NB; ++c) {
                    const uint32_t sk = bs & (16u / a.bwidth[c] - 1u);
                    tma_load_1d(smem + a.sm_bpay[c], static_cast<const char*>(a.bpay[c]) + static_cast<uint64_t>(bs - sk) * a.bwidth[c], pb[c], &s_bbar);
                    if (vb[c]) tma_load_1d(smem + a.sm_bvalid[c], a.bvalid[c] + (bs - (bs & 15u)), vb[c], &s_bbar);
                }
            }
            const uint32_t first = pe - ps < kBatch ? pe - ps : kBatch;
            issue_probe(batch_no & 1, ps, first);
        }
        uint32_t bkey[kBuildItems];
#pragma unroll
        for (int k = 0; k < kBuildItems; ++k) {
            const uint32_t i = k * kThreads + tid;
            bkey[k] = i < nb ? a.bkeys[bs + i] : 0u;
        }
        for (uint32_t s = tid; s < kSlots; s += kThreads) slots[s] = ~0ull;
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
            // not a key / foreign-key join: leave it to the general path (outstanding bulk copies land in this
            // CTA's shared memory before it retires)
            if (tid == 0) atomicExch(a.abort_flag, 1u);
            if (NB > 0) mbar_wait(&s_bbar, unit_no & 1);
            mbar_wait(&s_pbar[batch_no & 1], (batch_no >> 1) & 1);
            return;
        }
        if (NB > 0) mbar_wait(&s_bbar, unit_no & 1);
        ++unit_no;
        const uint32_t bskew[kEmitMaxPay] = {NB > 0 ? (bs & (16u / a.bwidth[0] - 1u)) : 0u, NB > 1 ? (bs & (16u / a.bwidth[NB > 1 ? 1 : 0] - 1u)) : 0u};

        // ---- probe ------------------------------------------------------------------------------------------
        for (uint32_t base = ps; base < pe; base += kBatch, ++batch_no) {
            const int      sb = batch_no & 1;
            const uint32_t cnt = pe - base < kBatch ? pe - base : kBatch;
            if (tid == 0 && base + kBatch < pe) {
                // the other buffer's last reader was the previous batch, which ended behind a barrier
                const uint32_t nbase = base + kBatch;
                issue_probe(sb ^ 1, nbase, pe - nbase < kBatch ? pe - nbase : kBatch);
            }
            mbar_wait(&s_pbar[sb], (batch_no >> 1) & 1);
            const uint32_t* pk = reinterpret_cast<const uint32_t*>(smem + a.sm_pkeys[sb]) + (base & 3u);

            // look every tuple up: at most one match, the table holds distinct keys
            uint32_t key[kItems], lidx[kItems];
            uint32_t pending = 0;
#pragma unroll
            for (int k = 0; k < kItems; ++k) {
                const uint32_t i = k * kThreads + tid;
                key[k]  = pk[i]; // past cnt: stale bytes, never used
                lidx[k] = kNone;
                if (i < cnt) {
                    uint32_t       sl   = (hash_key(key[k]) >> part_bits) & kSlotMask;
                    const uint32_t step = probe_step(key[k]);
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
            // validity of the nullable carried columns for this thread's matches: bit (nn * kItems + k)
            uint32_t okbits = 0;
#pragma unroll
            for (int c = 0; c < NB; ++c) {
                if (a.bnull[c] >= 0) {
                    const uint8_t* bv = smem + a.sm_bvalid[c] + (bs & 15u);
#pragma unroll
                    for (int k = 0; k < kItems; ++k)
                        if (((pending >> k) & 1u) && bv[lidx[k]]) okbits |= 1u << (a.bnull[c] * kItems + k);
                }
            }
#pragma unroll
            for (int c = 0; c < NP; ++c) {
                if (a.pnull[c] >= 0) {
                    const uint8_t* pv = smem + a.sm_pvalid[sb][c] + (base & 15u);
#pragma unroll
                    for (int k = 0; k < kItems; ++k)
                        if (((pending >> k) & 1u) && pv[k * kThreads + tid]) okbits |= 1u << (a.pnull[c] * kItems + k);
                }
            }

            // place the matches as rows of the open chunks (A: positions 0..1983, B: 1984..3967)
            for (;;) {
                // -- phase A: row range and value-slot ranges of this warp, from one packed atomic
                if (close_pending) {
                    __syncthreads(); // the chunk closed after the last round has been written and B has become A
                    close_pending = false;
                }
                // (counts are summed over the warp with REDUX: this thread's matches, and how many of them are
                // non-NULL in each nullable column)
                const uint32_t total = __reduce_add_sync(RJ_FULL_MASK, static_cast<uint32_t>(__popc(pending)));
                unsigned long long add = total;
#pragma unroll
                for (int nn = 0; nn < NN; ++nn) {
                    const uint32_t cntv = __reduce_add_sync(RJ_FULL_MASK, static_cast<uint32_t>(__popc(pending & (okbits >> (nn * kItems)) & ((1u << kItems) - 1u))));
                    add |= static_cast<unsigned long long>(cntv) << (kFieldBits * (nn + 1));
                }
                unsigned long long old = 0;
                if (lane == 0 && total) old = atomicAdd(&s_ctr, add);
                old = __shfl_sync(RJ_FULL_MASK, old, 0);
                const uint32_t off = static_cast<uint32_t>(old) & kFieldMask;
                // the warps whose rows contain the last row of a page (991, 1983, 2975, 3967) publish the
                // value counts up to it
                if (NN > 0 && total && (off / kHalfRows != (off + total - 1) / kHalfRows || (off + total) % kHalfRows == 0)) {
#pragma unroll
                    for (int h = 1; h <= 4; ++h) {
                        const uint32_t last = h * kHalfRows - 1;
                        if (off <= last && off + total > last) {
                            uint32_t v[NN > 0 ? NN : 1];
#pragma unroll
                            for (int nn = 0; nn < NN; ++nn) v[nn] = static_cast<uint32_t>(old >> (kFieldBits * (nn + 1))) & kFieldMask;
                            uint32_t o = off;
#pragma unroll
                            for (int k = 0; k < kItems; ++k) {
                                const uint32_t bal  = __ballot_sync(RJ_FULL_MASK, (pending >> k) & 1u);
                                const uint32_t p    = o + __popc(bal & lt);
                                const uint32_t upto = __ballot_sync(RJ_FULL_MASK, ((pending >> k) & 1u) && p <= last);
#pragma unroll
                                for (int nn = 0; nn < NN; ++nn)
                                    v[nn] += __popc(__ballot_sync(RJ_FULL_MASK, (pending >> k) & (okbits >> (nn * kItems + k)) & 1u) & upto);
                                o += __popc(bal);
                            }
#pragma unroll
                            for (int nn = 0; nn < NN; ++nn)
                                if (lane == 0) s_nvb[h][nn] = v[nn];
                        }
                    }
                }
                __syncthreads(); // every reservation of this round is in s_ctr, the page boundaries are published
                uint32_t       reserved = static_cast<uint32_t>(s_ctr) & kFieldMask;
                const uint32_t cA = s_chunk[0], cB = s_chunk[1];

                // -- phase B: every row below position 3968 goes to its final place
                if (total && off < kOpenRows) {
                    uint32_t pos[kItems];               // kNone: not placed in this round
                    uint32_t slot[NN > 0 ? NN : 1][kItems];
                    const uint32_t sbase = smem_u32(smem);
                    if (total == kItems * 32u && off + total <= kOpenRows) {
                        // dense round (every lane matched in every item -- the rule in a foreign-key join -- and all
                        // of it fits): rows are lanes in order, a column's validity bits are its four ballots
                        // shifted to the warp's first row, OR-ed in by five lanes with one atomic each
#pragma unroll
                        for (int k = 0; k < kItems; ++k) pos[k] = off + 32u * k + lane;
#pragma unroll
                        for (int nn = 0; nn < NN; ++nn) {
                            uint32_t vo = static_cast<uint32_t>(old >> (kFieldBits * (nn + 1))) & kFieldMask;
                            uint32_t prev = 0u, mine_lo = 0u, mine_hi = 0u; // lane i ORs word (off >> 5) + i = vb[i-1] : vb[i] funnel-shifted
#pragma unroll
                            for (int k = 0; k < kItems; ++k) {
                                const uint32_t vb = __ballot_sync(RJ_FULL_MASK, (okbits >> (nn * kItems + k)) & 1u);
                                slot[nn][k] = vo + __popc(vb & lt);
                                vo += __popc(vb);
                                if (lane == static_cast<uint32_t>(k)) { mine_hi = vb; mine_lo = prev; }
                                prev = vb;
                            }
                            if (lane == kItems) mine_lo = prev; // the last word: only what the shift carries over
                            const uint32_t word = __funnelshift_l(mine_lo, mine_hi, off & 31u);
                            if (lane <= kItems && word)
                                asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(sbase + a.sm_bitmap[nn] + ((off >> 5) + lane) * 4u), "r"(word) : "memory");
                        }
                        pending = 0;
                    } else {
                        uint32_t ro = off;
                        uint32_t vo[NN > 0 ? NN : 1];
#pragma unroll
                        for (int nn = 0; nn < NN; ++nn) vo[nn] = static_cast<uint32_t>(old >> (kFieldBits * (nn + 1))) & kFieldMask;
#pragma unroll
                        for (int k = 0; k < kItems; ++k) {
                            const bool     mine = (pending >> k) & 1u;
                            const uint32_t bal  = __ballot_sync(RJ_FULL_MASK, mine);
                            const uint32_t cntk = __popc(bal);
                            const uint32_t p    = ro + __popc(bal & lt);
                            const bool     fits = mine && p < kOpenRows;
                            pos[k] = fits ? p : kNone;
#pragma unroll
                            for (int nn = 0; nn < NN; ++nn) {
                                const bool     okk = (okbits >> (nn * kItems + k)) & 1u;
                                const uint32_t vb  = __ballot_sync(RJ_FULL_MASK, mine && okk);
                                slot[nn][k] = vo[nn] + __popc(vb & lt);
                                vo[nn] += __popc(vb);
                                // validity bits of the rows placed now, OR-ed into the bitmap (two words at most)
                                if (cntk && ro < kOpenRows) {
                                    const uint32_t word0 = ro >> 5, sh = ro & 31u;
                                    uint32_t lo, hi;
                                    if (bal == RJ_FULL_MASK || vb == bal) {
                                        // every lane has a row (rows = lanes in order), or no NULL among the rows: no shuffling
                                        const uint32_t nfit = ro + cntk <= kOpenRows ? cntk : kOpenRows - ro;
                                        uint32_t bits = bal == RJ_FULL_MASK ? vb : 0xffffffffu;
                                        if (nfit < 32) bits &= (1u << nfit) - 1u;
                                        lo = bits << sh;
                                        hi = sh ? bits >> (32u - sh) : 0u;
                                    } else {
                                        const uint32_t bit = (fits && okk) ? (1u << (p & 31u)) : 0u;
                                        lo = __reduce_or_sync(RJ_FULL_MASK, (p >> 5) == word0 ? bit : 0u);
                                        hi = __reduce_or_sync(RJ_FULL_MASK, (p >> 5) == word0 ? 0u : bit);
                                    }
                                    const uint32_t w = lane == 0 ? lo : hi;
                                    if (lane < 2 && w)
                                        asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(sbase + a.sm_bitmap[nn] + (word0 + lane) * 4u), "r"(w) : "memory");
                                }
                            }
                            ro += cntk;
                            if (fits) pending &= ~(1u << k);
                        }
                    }

                    // A column's four values go to their pages.  N = index among the nullable carried columns
                    // (-1: the column holds no NULL), WIDE = 8-byte values: pages 2 cA, 2 cA + 1 for positions below
                    // 1984, then 2 cB, 2 cB + 1; the slot inside a page = the row's value slot minus the values
                    // that precede the page (s_nvb), or row position minus the rows that precede it.
                    auto emit = [&](auto n_c, auto wide_c, int j, const uint64_t (&v)[kItems]) {
                        constexpr int  N    = decltype(n_c)::value;
                        constexpr bool WIDE = decltype(wide_c)::value;
                        uint8_t* const pages = a.out_pages[j];
                        if constexpr (!WIDE) {
                            uint32_t* const pgA = reinterpret_cast<uint32_t*>(pages + static_cast<uint64_t>(cA) * RJ_PAGE + 4);
                            uint32_t* const pgB = reinterpret_cast<uint32_t*>(pages + static_cast<uint64_t>(cB) * RJ_PAGE + 4);
                            const uint32_t subB = N >= 0 ? s_nvb[2][N >= 0 ? N : 0] : kChunkRows;
#pragma unroll
                            for (int k = 0; k < kItems; ++k) {
                                const uint32_t p = pos[k];
                                if (p == kNone) continue;
                                if (N >= 0 && !((okbits >> ((N >= 0 ? N : 0) * kItems + k)) & 1u)) continue; // NULL: no value stored
                                const uint32_t sl  = N >= 0 ? slot[N >= 0 ? N : 0][k] : p;
                                const bool     inB = p >= kChunkRows;
                                (inB ? pgB : pgA)[sl - (inB ? subB : 0u)] = static_cast<uint32_t>(v[k]);
                            }
                        } else {
                            uint64_t* const pgA = reinterpret_cast<uint64_t*>(pages + 2ull * cA * RJ_PAGE + 8);
                            uint64_t* const pgB = reinterpret_cast<uint64_t*>(pages + 2ull * cB * RJ_PAGE + 8) - 2 * (RJ_PAGE / 8); // indexed with h = 2, 3
#pragma unroll
                            for (int k = 0; k < kItems; ++k) {
                                const uint32_t p = pos[k];
                                if (p == kNone) continue;
                                if (N >= 0 && !((okbits >> ((N >= 0 ? N : 0) * kItems + k)) & 1u)) continue;
                                const uint32_t sl  = N >= 0 ? slot[N >= 0 ? N : 0][k] : p;
                                const uint32_t h   = ((p >> 5) * 2115u) >> 16; // p / 992 for p < 4096
                                const uint32_t sub = N >= 0 ? s_nvb[h][N >= 0 ? N : 0] : h * kHalfRows;
                                (h < 2 ? pgA : pgB)[h * (RJ_PAGE / 8) + sl - sub] = v[k];
                            }
                        }
                    };
                    // every output column that shows source `src`, column `c`
                    auto emit_source = [&](int src, int c, const uint64_t (&v)[kItems]) {
#pragma unroll
                        for (int j = 0; j < kEmitMaxOut; ++j) {
                            if (j >= a.n_out || a.out_src[j] != src || (src != 0 && a.out_idx[j] != c)) continue;
                            const int  nn   = NN > 0 ? a.out_null[j] : -1;
                            const bool wide = a.out_width[j] == 8;
                            auto with_n = [&](auto n_c) {
                                if (wide) emit(n_c, std::true_type{}, j, v); else emit(n_c, std::false_type{}, j, v);
                            };
                            if (nn < 0) with_n(std::integral_constant<int, -1>{});
                            if constexpr (NN > 0) { if (nn == 0) with_n(std::integral_constant<int, 0>{}); }
                            if constexpr (NN > 1) { if (nn == 1) with_n(std::integral_constant<int, 1>{}); }
                            if constexpr (NN > 2) { if (nn == 2) with_n(std::integral_constant<int, 2>{}); }
                            if constexpr (NN > 3) { if (nn == 3) with_n(std::integral_constant<int, 3>{}); }
                        }
                    };
                    // the values are fetched once per source: the key, the build columns (by table index), the
                    // probe columns (by batch offset)
                    {
                        uint64_t v[kItems];
#pragma unroll
                        for (int k = 0; k < kItems; ++k) v[k] = key[k];
                        emit_source(0, 0, v);
                    }
#pragma unroll
                    for (int c = 0; c < NB; ++c) {
                        uint64_t v[kItems];
                        const uint8_t* col = smem + a.sm_bpay[c];
                        if (a.bwidth[c] == 8) {
#pragma unroll
                            for (int k = 0; k < kItems; ++k) v[k] = pos[k] != kNone ? (reinterpret_cast<const uint64_t*>(col) + bskew[c])[lidx[k]] : 0ull;
                        } else {
#pragma unroll
                            for (int k = 0; k < kItems; ++k) v[k] = pos[k] != kNone ? (reinterpret_cast<const uint32_t*>(col) + bskew[c])[lidx[k]] : 0u;
                        }
                        emit_source(1, c, v);
                    }
#pragma unroll
                    for (int c = 0; c < NP; ++c) {
                        uint64_t v[kItems];
                        const uint8_t* col = smem + a.sm_ppay[sb][c];
                        if (a.pwidth[c] == 8) {
#pragma unroll
                            for (int k = 0; k < kItems; ++k) v[k] = (reinterpret_cast<const uint64_t*>(col) + (base & 1u))[k * kThreads + tid];
                        } else {
#pragma unroll
                            for (int k = 0; k < kItems; ++k) v[k] = (reinterpret_cast<const uint32_t*>(col) + (base & 3u))[k * kThreads + tid];
                        }
                        emit_source(2, c, v);
                    }
                }
                __syncthreads(); // the bitmaps of this round are complete; s_ctr, s_chunk and s_nvb have been read
                const bool again = reserved > kOpenRows; // rows beyond chunk B (a full batch on top of a nearly full A): once more
                while (reserved >= kChunkRows) {
                    close_full_chunk(reserved);
                    reserved = (reserved > kOpenRows ? kOpenRows : reserved) - kChunkRows;
                    close_pending = true; // everybody meets the closing warps before the next reservation
                }
                if (!again) break;
            }
        }
    }
    // the partly filled chunk A of this CTA, and the chunks it held in reserve (B, one ahead): empty pages
    __syncthreads();
    {
        const unsigned long long ctr = s_ctr;
        const uint32_t n = static_cast<uint32_t>(ctr) & kFieldMask; // < 1984: full chunks were closed in the loop
        uint32_t nv1[kMaxNull], nva[kMaxNull], zero[kMaxNull];
#pragma unroll
        for (int nn = 0; nn < kMaxNull; ++nn) {
            nva[nn]  = nn < NN ? static_cast<uint32_t>(ctr >> (kFieldBits * (nn + 1))) & kFieldMask : 0u;
            nv1[nn]  = nn < NN ? (n > kHalfRows ? s_nvb[1][nn] : nva[nn]) : 0u;
            zero[nn] = 0u;
        }
        write_chunk_meta(s_chunk[0], n, nv1, nva);
        write_chunk_meta(s_chunk[1], 0u, zero, zero);
        uint32_t ahead = __shfl_sync(RJ_FULL_MASK, chunk_ahead, 0);
        __shared__ uint32_t s_ahead;
        if (tid == 0) {
            s_ahead = ahead;
            atomicAdd(a.row_counter, static_cast<unsigned long long>(n));
        }
        __syncthreads();
        write_chunk_meta(s_ahead, 0u, zero, zero);
    }
}

} // namespace

// one instantiation's launch: its attributes (dynamic shared-memory limit, carve-out) are set once per device.
// (The statics must be per INSTANTIATION: every join_emit_kernel<...> has the same function-pointer type, so a
// generic lambda taking the pointer would share one set of statics among all of them.)
template <int NB, int NP, int NN>
static void launch_emit_instance(const EmitArgs& a, size_t smem, unsigned grid, cudaStream_t s) {
    auto kern = join_emit_kernel<NB, NP, NN>;
    static SmemConfigured cfg;
    static bool carved[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!carved[dev & 63]) {
        // two CTAs of ~112 KB per SM need the whole shared-memory carve-out
        RJ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        carved[dev & 63] = true;
    }
    cfg.ensure(kern, smem);
    kern<<<grid, kThreads, smem, s>>>(a);
}

static size_t join_emit_layout(const JoinEmitLaunch& L, EmitArgs* a) {
    size_t off = sizeof(uint64_t) * kEmitSlots;
    auto   take = [&](size_t bytes) {
        off = (off + 15) / 16 * 16;
        const size_t at = off;
        off += bytes;
        return static_cast<uint32_t>(at);
    };
    EmitArgs tmp{};
    EmitArgs& r = a ? *a : tmp;
    for (int c = 0; c < L.n_bpay; ++c) r.sm_bpay[c] = take(size_t(kEmitBuildCap) * L.bwidth[c] + 16);
    for (int c = 0; c < L.n_bpay; ++c) r.sm_bvalid[c] = L.bvalid[c] ? take(kEmitBuildCap + 16) : 0;
    for (int s = 0; s < 2; ++s) {
        r.sm_pkeys[s] = take(size_t(kBatch) * 4 + 16);
        for (int c = 0; c < L.n_ppay; ++c) r.sm_ppay[s][c] = take(size_t(kBatch) * L.pwidth[c] + 16);
        for (int c = 0; c < L.n_ppay; ++c) r.sm_pvalid[s][c] = L.pvalid[c] ? take(kBatch + 16) : 0;
    }
    int n_null = 0;
    for (int c = 0; c < L.n_bpay; ++c) n_null += L.bvalid[c] ? 1 : 0;
    for (int c = 0; c < L.n_ppay; ++c) n_null += L.pvalid[c] ? 1 : 0;
    for (int i = 0; i < n_null; ++i) r.sm_bitmap[i] = take(2 * kBitmapWords * 4);
    return off;
}

// dynamic shared memory one CTA may use: two CTAs per SM when the columns allow it, else one
static size_t join_emit_smem_cap(size_t need) { return need <= 113 * 1024 ? 113 * 1024 : 226 * 1024; }

bool join_emit_fits(const JoinEmitLaunch& L) {
    if (L.n_out < 1 || L.n_out > kEmitMaxOut || L.n_bpay > kEmitMaxPay || L.n_ppay > kEmitMaxPay) return false;
    return join_emit_layout(L, nullptr) <= 226 * 1024;
}

void launch_join_emit(const JoinEmitLaunch& L, int sm_count, cudaStream_t s) {
    EmitArgs a{};
    a.bkeys = L.bkeys; a.pkeys = L.pkeys; a.off_b = L.off_b; a.off_p = L.off_p;
    a.unit_start = L.unit_start; a.unit_cursor = L.unit_cursor; a.nparts = L.nparts; a.part_bits = L.part_bits;
    a.probe_chunk = kEmitProbeChunk;
    a.n_out = L.n_out;
    int n_null = 0;
    for (int c = 0; c < kEmitMaxPay; ++c) {
        a.bpay[c] = L.bpay[c]; a.bvalid[c] = L.bvalid[c]; a.bwidth[c] = L.bwidth[c] ? L.bwidth[c] : 4;
        a.bnull[c] = (c < L.n_bpay && L.bvalid[c]) ? n_null++ : -1;
    }
    for (int c = 0; c < kEmitMaxPay; ++c) {
        a.ppay[c] = L.ppay[c]; a.pvalid[c] = L.pvalid[c]; a.pwidth[c] = L.pwidth[c] ? L.pwidth[c] : 4;
        a.pnull[c] = (c < L.n_ppay && L.pvalid[c]) ? n_null++ : -1;
    }
    a.n_null = n_null;
    for (int j = 0; j < kEmitMaxOut; ++j) {
        a.out_src[j] = L.out_src[j]; a.out_idx[j] = L.out_idx[j]; a.out_width[j] = L.out_width[j];
        a.out_pages[j] = L.out_pages[j];
        a.out_null[j] = -1;
        if (j < L.n_out && L.out_src[j] == 1) a.out_null[j] = a.bnull[L.out_idx[j]];
        if (j < L.n_out && L.out_src[j] == 2) a.out_null[j] = a.pnull[L.out_idx[j]];
    }
    a.chunk_counter = L.chunk_counter; a.row_counter = L.row_counter; a.abort_flag = L.abort_flag;
    const size_t smem = join_emit_layout(L, &a);
    if (smem > 226 * 1024) throw CudaError("join_emit: the columns do not fit shared memory");
    (void)join_emit_smem_cap;
    const unsigned grid = join_emit_grid(sm_count);
    // instantiated for every (build columns, probe columns, nullable ones among them)
    bool launched = false;
    auto try_launch = [&](auto nb_c, auto np_c, auto nn_c) {
        constexpr int B = decltype(nb_c)::value, P = decltype(np_c)::value, N = decltype(nn_c)::value;
        if constexpr (N <= B + P) {
            if (!launched && L.n_bpay == B && L.n_ppay == P && n_null == N) {
                launch_emit_instance<B, P, N>(a, smem, grid, s);
                launched = true;
            }
        }
    };
    auto for_nn = [&](auto nb_c, auto np_c) {
        try_launch(nb_c, np_c, std::integral_constant<int, 0>{});
        try_launch(nb_c, np_c, std::integral_constant<int, 1>{});
        try_launch(nb_c, np_c, std::integral_constant<int, 2>{});
        try_launch(nb_c, np_c, std::integral_constant<int, 3>{});
        try_launch(nb_c, np_c, std::integral_constant<int, 4>{});
    };
    auto for_np = [&](auto nb_c) {
        for_nn(nb_c, std::integral_constant<int, 0>{});
        for_nn(nb_c, std::integral_constant<int, 1>{});
        for_nn(nb_c, std::integral_constant<int, 2>{});
    };
    for_np(std::integral_constant<int, 0>{});
    for_np(std::integral_constant<int, 1>{});
    for_np(std::integral_constant<int, 2>{});
    if (!launched) throw CudaError("join_emit: unsupported column configuration");
    RJ_LAUNCH_CHECK();
}

} // namespace rj
