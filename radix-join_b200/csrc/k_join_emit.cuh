// Root join fused with page output: build + probe in shared memory, result PAGES written straight from
// the join kernel.  (The kernel template; k_join_emit.cu / k_join_emit_b1.cu / k_join_emit_b2.cu instantiate
// it for 0 / 1 / 2 carried build columns so that the instantiations compile in parallel.)
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
// Memory pipeline.  A unit's probe tuples are consumed in batches of 512 (one ITEM of 32 tuples per warp); the
// keys, carried values and validity bytes of a batch arrive by 1-D TMA bulk copies into one of eight
// shared-memory buffers.  A buffer is handed back by its 16 consumer warps through an mbarrier (one arrival per
// warp), so warps never wait for one another inside a unit; the thread that requests batches does so as far
// ahead as buffers are free and never waits for a reader unless the batch it needs itself is missing.  (With
// two buffers of 2048 tuples that thread waited for the slowest reader at every batch, became the slowest
// itself, and the whole CTA marched in its step: ~10 polls of the barrier per item.)
//
// Result pages.  Row alignment across columns (include/plan.h:102-105: columns are row-aligned by
// cumulative row index, page boundaries are free) is kept by emitting CHUNKS of 1984 rows: one page of
// every 4-byte column (1984 rows, the engine's fixed fill) and two pages of 992 rows of every 8-byte
// column.  Chunk c owns pages c / 2c, 2c+1 of every column, so the columns' page lists enumerate the same
// rows in the same order.  (8-byte pages hold 992 instead of 1007 rows: 1.5 % more pages, the price of
// page-aligned chunks.)
//
// EVERY WARP OWNS ITS OPEN CHUNK (round-2 rewrite; the earlier version shared two open chunks per CTA and
// spent two thirds of its 1320 warp instructions per 128 tuples on the packed shared-memory atomic, the
// REDUX counts, the page-boundary hand-over between warps and three CTA barriers per batch).  A page stores
// only its non-NULL values, packed (src/build_table.cpp:484-501), so a row's value slot depends on every
// earlier row of its page; with a chunk per warp all of that is warp-local register state: rows so far,
// non-NULL values so far per nullable column (and at the 992-row page boundary), and the chunk's validity
// bitmaps -- 62 words per nullable column, two per lane.  A batch item (32 tuples) is placed with one
// ballot per nullable column; values go straight from registers to their final place in global memory,
// consecutive rows of the warp to consecutive addresses.  A chunk is reserved with one global atomic,
// issued a few items before the open one fills up so that its latency is never waited for.
// Cost: every emitting warp ends the launch with a partly filled chunk; the launcher lets only as many warps per
// CTA emit as get kEmitMinChunksPerWarp (results shipped over PCIe) / kEmitMinChunksResident chunks' worth of probe
// tuples each (<= ~3 % / 6 % more pages), the others idle.
//
// Build keys must be unique inside every table (every key / foreign-key join): the 64-bit CAS insert sees
// an equal key for free, raises a global flag and the whole launch is abandoned -- the engine then runs
// the general path, which handles duplicates with chains.
#pragma once
#include "rj_common.cuh"
#include "rj_internal.h"

#include <type_traits>

namespace rj {
namespace emit {

constexpr int      kThreads    = 512;
constexpr int      kWarps      = kThreads / 32;
constexpr uint32_t kSlots      = kEmitSlots;       // 4096 x 64-bit (key | local build index << 32)
constexpr uint32_t kSlotMask   = kSlots - 1;
constexpr uint32_t kCap        = kEmitBuildCap;    // 3072 build tuples per table (75 % fill)
constexpr int      kBuildItems = kCap / kThreads;  // 6
constexpr int      kConsumers  = kWarps - 1;       // warps that probe; the last warp requests the data and looks the next unit up
constexpr uint32_t kBatch      = kConsumers * 32;  // probe tuples per batch: one ITEM of 32 per consumer warp
constexpr uint32_t kStages     = 8;                // probe batches in flight per CTA (power of two)
constexpr uint32_t kBarBytes   = 256;              // mbarriers in front of the table
constexpr uint32_t kUpFront    = 4;                // batches requested before the table is built (the others: while it is probed)
constexpr uint32_t kChunkRows  = kEmitChunkRows;   // 1984 = 62 bitmap words
constexpr uint32_t kHalfRows   = kChunkRows / 2;   // 992 rows per 8-byte page = 31 bitmap words
constexpr uint32_t kHalfWords  = kHalfRows / 32;
constexpr uint32_t kNone       = 0xffffffffu;
constexpr uint32_t kReserveAt  = kChunkRows - 6 * 32; // the next chunk is reserved ~6 items before it is needed
constexpr int      kSources    = 1 + 2 * kEmitMaxPay; // the key, the build columns, the probe columns

__device__ __forceinline__ uint32_t probe_step(uint32_t k) { return ((k * 0x9E3779B1u) >> 20) | 1u; }

struct EmitArgs {
    const uint32_t* bkeys;
    const uint32_t* pkeys;
    const uint32_t* off_b;
    const uint32_t* off_p;
    const uint32_t* unit_start;
    uint32_t*       unit_cursor;
    const uint32_t* unit_part;     // partition of unit u for u < unit_part_cap
    uint32_t        unit_part_cap;
    uint32_t        nparts;
    int             part_bits;
    uint32_t        probe_chunk; // probe tuples per work unit
    // carried columns, in final partition order beside the keys
    const void*    bpay[kEmitMaxPay];
    const uint8_t* bvalid[kEmitMaxPay]; // one byte per tuple, NULL = the column holds no NULL
    int            bwidth[kEmitMaxPay];
    const void*    ppay[kEmitMaxPay];
    const uint8_t* pvalid[kEmitMaxPay];
    int            pwidth[kEmitMaxPay];
    // output columns by SOURCE (0 = the join key, 1 + c = build column c, 1 + kEmitMaxPay + c = probe column c):
    // src_pages = the pages of the first output column that shows the source (NULL: none does), src_rest = bit j
    // for every further output column j that shows it (SELECT a, a ...: rare)
    uint8_t* src_pages[kSources];
    uint32_t src_rest[kSources];
    uint8_t* out_pages[kEmitMaxOut];
    // shared-memory layout (byte offsets into the dynamic segment, computed by the launcher); the second probe
    // buffer lies sm_pstride bytes behind the first
    uint32_t sm_bpay[kEmitMaxPay], sm_bvalid[kEmitMaxPay];
    uint32_t sm_pkeys, sm_ppay[kEmitMaxPay], sm_pvalid[kEmitMaxPay], sm_pstride;
    // results
    uint32_t*           chunk_counter;
    unsigned long long* row_counter;
    uint32_t*           abort_flag; // set when a table meets a duplicate build key
    int direct;   // the table is the rank structure over the hash bits left by the partitioning (see the kernel)
    uint32_t n_active; // consumer warps that probe (1 .. kConsumers): the others idle, see the launcher
    int all_once; // the key and every carried column are shown by exactly one output column (no SELECT a, a; key shown)
};

__device__ __forceinline__ uint32_t round16(uint32_t b) { return (b + 15u) & ~15u; }

constexpr int popc_c(int x) { return x == 0 ? 0 : (x & 1) + popc_c(x >> 1); }

// NB / NP: carried build / probe columns; NM: which of them hold NULLs (bit c = build column c, bit
// kEmitMaxPay + c = probe column c; the nullable columns are numbered in bit order); WM: which of them are 8
// bytes wide (same bits), or -1 = the widths are read from the arguments at run time.
template <int NB, int NP, int NM, int WM>
__global__ void __launch_bounds__(kThreads, 2) join_emit_kernel(const __grid_constant__ EmitArgs a) {
    constexpr int NN  = popc_c(NM);
    constexpr int NNX = NN > 0 ? NN : 1;
    // compile-time loops over the carried columns: f(column index constant, nullable index constant (-1: no NULLs))
    auto for_build = [](auto&& f) {
        if constexpr (NB > 0) f(std::integral_constant<int, 0>{}, std::integral_constant<int, (NM & 1) ? 0 : -1>{});
        if constexpr (NB > 1) f(std::integral_constant<int, 1>{}, std::integral_constant<int, (NM & 2) ? popc_c(NM & 1) : -1>{});
    };
    auto for_probe = [](auto&& f) {
        if constexpr (NP > 0) f(std::integral_constant<int, 0>{}, std::integral_constant<int, (NM & 4) ? popc_c(NM & 3) : -1>{});
        if constexpr (NP > 1) f(std::integral_constant<int, 1>{}, std::integral_constant<int, (NM & 8) ? popc_c(NM & 7) : -1>{});
    };
    auto bwide = [&](int c) -> bool { return WM >= 0 ? ((WM >> c) & 1) != 0 : a.bwidth[c] == 8; };
    auto pwide = [&](int c) -> bool { return WM >= 0 ? ((WM >> (kEmitMaxPay + c)) & 1) != 0 : a.pwidth[c] == 8; };
    extern __shared__ __align__(128) uint8_t smem_all[];
    // the mbarriers live at the start of the dynamic segment so that every shared-memory address the probe loop
    // needs is ONE register (sbase) plus a constant or an argument
    uint64_t* const s_bars  = reinterpret_cast<uint64_t*>(smem_all);
    uint64_t&       s_bbar  = s_bars[0];
    uint64_t* const s_full  = s_bars + 1;
    uint64_t* const s_empty = s_bars + 1 + kStages;
    uint8_t* const  smem    = smem_all + kBarBytes; // the table, then the columns (offsets: EmitArgs::sm_*)
    const uint32_t  sbase   = smem_u32(smem);
    __shared__ uint32_t s_desc[2][6]; // [round & 1]: the round's unit -- its number, then unit_start / off_b / off_b+1 / off_p / off_p+1 of its partition

    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t lt = lanemask_lt();
    const uint32_t n_units = a.unit_start[a.nparts];
    const int      part_bits = a.part_bits;
    if (tid == 0) {
        mbar_init(&s_bbar, 1);
        for (uint32_t st = 0; st < kStages; ++st) {
            mbar_init(&s_full[st], 1);
            mbar_init(&s_empty[st], a.n_active);
        }
        fence_mbar_init();
    }
    uint32_t unit_no = 0, batch_no = 0; // mbarrier phases

    // ---- the warp's open chunk: warp-uniform register state (c_next: lane 0) ---------------------------------
    uint32_t c_open = kNone;     // chunk being filled
    uint32_t c_next = kNone;     // lane 0: chunk reserved ahead (valid when have_next)
    bool     have_next = false;
    uint32_t rows = 0;           // rows placed in the open chunk
    uint32_t closed = 0;         // full chunks this warp has written
    uint32_t nv[NNX], nvh[NNX];  // non-NULL values of nullable column nn among all rows / among rows [0, 992)
    uint32_t pend[NNX];          // validity bits of the chunk's last, incomplete bitmap word
    uint32_t held[NNX];          // PER LANE: the latest complete bitmap word w with (w & 31) == lane, until it is flushed
#pragma unroll
    for (int nn = 0; nn < NNX; ++nn) nv[nn] = nvh[nn] = pend[nn] = held[nn] = 0u;
    // first page of the open chunk in the (first) output column of the key / of every carried column
    uint8_t* cb_key = nullptr;
    uint8_t* cb_b[NB > 0 ? NB : 1] = {};
    uint8_t* cb_p[NP > 0 ? NP : 1] = {};
    const bool direct = a.direct != 0;
    const bool all_once = a.all_once != 0;
    const bool all_active = a.n_active == static_cast<uint32_t>(kConsumers); // every consumer warp probes: one item per warp and batch

    // every output column that shows source S: f(first page of chunk c in that column)
    auto for_outputs = [&](int S, bool wide, uint32_t c, auto&& f) {
        const uint64_t chunk_bytes = wide ? 2ull * RJ_PAGE : 1ull * RJ_PAGE;
        uint8_t* const first = a.src_pages[S];
        if (first != nullptr) f(first + chunk_bytes * c);
        for (uint32_t m = a.src_rest[S]; m; m &= m - 1) f(a.out_pages[__ffs(m) - 1] + chunk_bytes * c);
    };
    // where word w of a chunk's validity bitmap lies in a column's FULL pages (4-byte column: one page of 62
    // words; 8-byte column: two pages of 31 words); cb = the chunk's first page in the column
    auto bitmap_word = [&](uint8_t* cb, bool wide, uint32_t w) -> uint32_t* {
        if (wide) {
            const uint32_t h = w >= kHalfWords ? 1u : 0u;
            return reinterpret_cast<uint32_t*>(cb + h * RJ_PAGE + (RJ_PAGE - kHalfRows / 8)) + (w - h * kHalfWords);
        }
        return reinterpret_cast<uint32_t*>(cb + (RJ_PAGE - kChunkRows / 8)) + w;
    };

    // one batch of probe tuples [base, base + cnt) into its buffer (one thread).  q = the batch's number: the
    // buffer was last used by batch q - kStages, whose readers hand it back through s_empty
    auto issue_probe = [&](uint32_t q, uint32_t base, uint32_t cnt) {
        const uint32_t s = q & (kStages - 1);
        uint8_t* const buf = smem + s * a.sm_pstride;
        if (q >= kStages) mbar_wait(&s_empty[s], ((q / kStages) - 1) & 1);
        uint32_t bytes = round16((cnt + (base & 3u)) * 4u);
        uint32_t pb[kEmitMaxPay] = {}, vb[kEmitMaxPay] = {};
#pragma unroll
        for (int c = 0; c < NP; ++c) {
            const uint32_t w = pwide(c) ? 8u : 4u;
            pb[c] = round16((cnt + (base & (16u / w - 1u))) * w);
            vb[c] = (NM >> (kEmitMaxPay + c)) & 1 ? round16(cnt + (base & 15u)) : 0u;
            bytes += pb[c] + vb[c];
        }
        mbar_arrive_expect_tx(&s_full[s], bytes);
        tma_load_1d(buf + a.sm_pkeys, a.pkeys + (base - (base & 3u)), round16((cnt + (base & 3u)) * 4u), &s_full[s]);
#pragma unroll
        for (int c = 0; c < NP; ++c) {
            const uint32_t w = pwide(c) ? 8u : 4u;
            const uint32_t sk = base & (16u / w - 1u);
            tma_load_1d(buf + a.sm_ppay[c], static_cast<const char*>(a.ppay[c]) + static_cast<uint64_t>(base - sk) * w, pb[c], &s_full[s]);
            if (vb[c]) tma_load_1d(buf + a.sm_pvalid[c], a.pvalid[c] + (base - (base & 15u)), vb[c], &s_full[s]);
        }
    };

    // Pages of chunk c, which holds n rows, in every output column of source S: headers, and the validity
    // bitmaps that are not in place yet.  (x0, x1) = the chunk's bitmap in lane layout (lane l: words l and
    // 32 + l); in_place: the bitmaps of FULL pages have been written word by word already (nullable sources);
    // v_all / v_half = non-NULL values among all rows / among rows [0, 992).
    auto write_meta_source = [&](int S, bool wide, uint32_t c, uint32_t n, uint32_t x0, uint32_t x1, bool in_place, uint32_t v_all, uint32_t v_half) {
        for_outputs(S, wide, c, [&](uint8_t* cb) {
            const int pages = wide ? 2 : 1;
            for (int h = 0; h < pages; ++h) {
                uint8_t* pg = cb + h * RJ_PAGE;
                const uint32_t rows_pg = !wide ? n : (h == 0 ? (n < kHalfRows ? n : kHalfRows) : (n > kHalfRows ? n - kHalfRows : 0u));
                const uint32_t fw = h * kHalfWords; // first chunk word of the page's bitmap
                const uint32_t bytes = (rows_pg + 7) >> 3;
                uint8_t* dst = pg + RJ_PAGE - bytes;
                if (in_place && n == kChunkRows) {
                    // a full chunk of a nullable source: every word has been flushed to its place
                } else if ((rows_pg & 31u) == 0) { // word-aligned
                    const uint32_t words = rows_pg >> 5;
                    if (lane >= fw && lane - fw < words) reinterpret_cast<uint32_t*>(dst)[lane - fw] = x0;
                    if (lane + 32 - fw < words) reinterpret_cast<uint32_t*>(dst)[lane + 32 - fw] = x1;
                } else {
                    for (uint32_t b0 = 0; b0 < bytes; b0 += 32) {
                        const uint32_t b = b0 + lane;
                        const uint32_t cw = fw + (b >> 2);
                        const uint32_t w0 = __shfl_sync(RJ_FULL_MASK, x0, cw & 31u), w1 = __shfl_sync(RJ_FULL_MASK, x1, cw & 31u);
                        if (b < bytes) dst[b] = static_cast<uint8_t>((cw < 32 ? w0 : w1) >> (8 * (b & 3u)));
                    }
                }
                if (lane == 0) *reinterpret_cast<uint32_t*>(pg) = rows_pg | ((!wide ? v_all : (h == 0 ? v_half : v_all - v_half)) << 16);
            }
        });
    };
    auto write_meta = [&](uint32_t c, uint32_t n) {
        // a column without NULLs: ones below row n
        const uint32_t fullw = n >> 5, tail = (1u << (n & 31u)) - 1u;
        const uint32_t o0 = lane < fullw ? 0xffffffffu : (lane == fullw ? tail : 0u);
        const uint32_t o1 = lane + 32 < fullw ? 0xffffffffu : (lane + 32 == fullw ? tail : 0u);
        const uint32_t nh = n < kHalfRows ? n : kHalfRows;
        // a nullable column of a partly filled chunk: the complete words that were flushed come back from where
        // they were stored (the positions of a full page), the others from `held`, the incomplete one from `pend`
        auto nullable_meta = [&](int S, bool wide, auto n_c) {
            constexpr int N = decltype(n_c)::value;
            uint32_t x0 = 0u, x1 = 0u;
            if (n != kChunkRows && (a.src_pages[S] != nullptr || a.src_rest[S] != 0u)) {
                uint8_t* const pages = a.src_pages[S] != nullptr ? a.src_pages[S] : a.out_pages[__ffs(a.src_rest[S]) - 1];
                uint8_t* const cb = pages + (wide ? 2ull * RJ_PAGE : 1ull * RJ_PAGE) * c;
                const uint32_t first_flush = wide ? kHalfWords - 1 : 31u;         // words [0, first_flush] are flushed together
                const bool     flushed = fullw > first_flush;
                __syncwarp();
                if (lane < fullw) x0 = (flushed && lane <= first_flush) ? __ldcg(bitmap_word(cb, wide, lane)) : held[N];
                else if (lane == fullw) x0 = pend[N];
                if (lane + 32 < fullw) x1 = held[N];
                else if (lane + 32 == fullw) x1 = pend[N];
                __syncwarp();
            }
            write_meta_source(S, wide, c, n, x0, x1, true, nv[N], n >= kHalfRows ? nvh[N] : nv[N]);
        };
        write_meta_source(0, false, c, n, o0, o1, false, n, nh); // the key: INT32, never NULL in a match
        for_build([&](auto c_c, auto n_c) {
            constexpr int C = decltype(c_c)::value, N = decltype(n_c)::value;
            if constexpr (N >= 0) nullable_meta(1 + C, bwide(C), n_c);
            else write_meta_source(1 + C, bwide(C), c, n, o0, o1, false, n, nh);
        });
        for_probe([&](auto c_c, auto n_c) {
            constexpr int C = decltype(c_c)::value, N = decltype(n_c)::value;
            if constexpr (N >= 0) nullable_meta(1 + kEmitMaxPay + C, pwide(C), n_c);
            else write_meta_source(1 + kEmitMaxPay + C, pwide(C), c, n, o0, o1, false, n, nh);
        });
    };

    // a chunk for the rows that are about to be placed
    auto ensure_open = [&]() {
        if (c_open == kNone) {
            if (!have_next && lane == 0) c_next = atomicAdd(a.chunk_counter, 1u);
            c_open = __shfl_sync(RJ_FULL_MASK, c_next, 0);
            have_next = false;
            cb_key = a.src_pages[0] + static_cast<uint64_t>(c_open) * RJ_PAGE;
            for_build([&](auto c_c, auto) { cb_b[decltype(c_c)::value] = a.src_pages[1 + decltype(c_c)::value] + static_cast<uint64_t>(c_open) * (bwide(decltype(c_c)::value) ? 2 * RJ_PAGE : RJ_PAGE); });
            for_probe([&](auto c_c, auto) { cb_p[decltype(c_c)::value] = a.src_pages[1 + kEmitMaxPay + decltype(c_c)::value] + static_cast<uint64_t>(c_open) * (pwide(decltype(c_c)::value) ? 2 * RJ_PAGE : RJ_PAGE); });
        }
    };
    // rows have been placed: close the chunk when it is full, reserve the next one shortly before
    auto after_place = [&]() {
        if (rows == kChunkRows) {
            write_meta(c_open, kChunkRows);
            ++closed;
            c_open = kNone;
            rows = 0;
#pragma unroll
            for (int nn = 0; nn < NNX; ++nn) nv[nn] = nvh[nn] = pend[nn] = held[nn] = 0u;
        } else if (!have_next && rows >= kReserveAt) {
            if (lane == 0) c_next = atomicAdd(a.chunk_counter, 1u); // consumed when the open chunk closes
            have_next = true;
        }
    };

    // Work units are handed out in global order (see k_join.cu) by an atomic cursor.  The producer warp's first
    // lane looks the NEXT unit up while the consumers probe the current one, so none of its dependent global loads
    // (cursor -> the unit's partition -> the partition's offsets) is ever waited for by the CTA; it leaves the
    // descriptor in s_desc[next round & 1] and pulls the unit's build tuples towards L2.
    auto lookup_unit = [&](uint32_t* d) { // one thread
        uint32_t u = atomicAdd(a.unit_cursor, 1u);
        if (*reinterpret_cast<volatile uint32_t*>(a.abort_flag)) u = 0xffffffffu; // somebody met a duplicate key
        d[0] = u;
        if (u >= n_units) return;
        uint32_t part;
        if (u < a.unit_part_cap) {
            part = a.unit_part[u];
        } else { // beyond the table (heavily skewed partitions): unit_start[lo] <= u < unit_start[hi]
            uint32_t lo = 0, hi = a.nparts;
            while (hi - lo > 1) {
                const uint32_t mid = (lo + hi) >> 1;
                if (a.unit_start[mid] <= u) lo = mid; else hi = mid;
            }
            part = lo;
        }
        const uint32_t us = a.unit_start[part], blo = a.off_b[part], bhi = a.off_b[part + 1], plo = a.off_p[part], phi = a.off_p[part + 1];
        d[1] = us; d[2] = blo; d[3] = bhi; d[4] = plo; d[5] = phi;
        // the unit's build tuples towards L2 (its probe tuples are requested by TMA while the table is built)
        const uint32_t n_pc = (phi - plo + a.probe_chunk - 1) / a.probe_chunk;
        const uint32_t bs_n = blo + ((u - us) / n_pc) * kCap;
        const uint32_t nb_n = bhi - bs_n > kCap ? kCap : bhi - bs_n;
        // (whole 16-byte pieces strictly inside the unit's tuples: a prefetch never leaves the arrays)
        const uint32_t f4 = (bs_n + 3u) & ~3u, l4 = (bs_n + nb_n) & ~3u;     // 4-element (>= 16-byte) boundaries
        const uint32_t f16 = (bs_n + 15u) & ~15u, l16 = (bs_n + nb_n) & ~15u;
        if (l4 > f4) prefetch_l2_bulk(a.bkeys + f4, (l4 - f4) * 4u);
#pragma unroll
        for (int c = 0; c < NB; ++c) {
            const uint32_t w = bwide(c) ? 8u : 4u;
            if (l4 > f4) prefetch_l2_bulk(static_cast<const char*>(a.bpay[c]) + static_cast<uint64_t>(f4) * w, (l4 - f4) * w);
            if (((NM >> c) & 1) && l16 > f16) prefetch_l2_bulk(a.bvalid[c] + f16, l16 - f16);
        }
    };
    const bool producer = (tid >> 5) == static_cast<uint32_t>(kConsumers);
    if (producer && lane == 0) lookup_unit(s_desc[0]);
    for (uint32_t round = 0;; ++round) {
        __syncthreads(); // every warp is done with the previous unit's table and columns; this round's descriptor is in place
        const uint32_t* const dsc = s_desc[round & 1];
        const uint32_t u = dsc[0];
        if (u >= n_units) break;
        const uint32_t local = u - dsc[1];
        const uint32_t b_lo = dsc[2], b_hi = dsc[3];
        const uint32_t p_lo = dsc[4], p_hi = dsc[5];
        const uint32_t n_pchunks = (p_hi - p_lo + a.probe_chunk - 1) / a.probe_chunk;
        const uint32_t bc = local / n_pchunks, pc = local - bc * n_pchunks;
        const uint32_t bs = b_lo + bc * kCap;
        const uint32_t nb = (b_hi - bs > kCap) ? kCap : b_hi - bs;
        const uint32_t ps = p_lo + pc * a.probe_chunk;
        const uint32_t pe = (p_hi - ps > a.probe_chunk) ? ps + a.probe_chunk : p_hi;
        const uint32_t n_batches = (pe - ps + kBatch - 1) / kBatch;

        // ---- build: the first probe batch is requested, then the table is filled ------------------------------
        // Two kinds of table.
        // HASH (any partition count): 4096 slots of (key, build index), double hashing; the carried build columns
        //   arrive by TMA in partition order and are indexed with the build index.
        // DIRECT (the partitioning consumed so many hash bits that at most 17 are left, i.e. >= 2^15 partitions):
        //   fmix32 is a bijection on 32-bit keys, so inside a partition the remaining hash bits x IDENTIFY the
        //   key.  The table is a bitmap over x with a running count per 32-bit word, (bits, rank of the word's first
        //   bit) in one 64-bit entry: a probe is ONE shared-memory load, a bit test and a popcount -- no key
        //   comparison, no probe sequence, no divergence between the lanes of a warp (the hash table's probe loop
        //   ran as long as the unluckiest of 32 lanes: ~4 rounds at 50 % fill).  The carried build columns are
        //   stored by rank.  A duplicate build key shows up as a bit that is already set.
        const int      rem_words_log = 32 - part_bits - 5;             // direct: 2^(32 - part_bits) bits
        const uint32_t bskew1 = direct ? 0u : (bs & 15u); // TMA windows start on a 16-byte boundary: elements to skip
        if (producer && lane == 0) {
            // (the barrier at the top of the loop ended every read of the previous unit's table and columns)
            if (NB > 0 && !direct) {
                uint32_t bytes = 0, pb[kEmitMaxPay] = {}, vb[kEmitMaxPay] = {};
#pragma unroll
                for (int c = 0; c < NB; ++c) {
                    const uint32_t w = bwide(c) ? 8u : 4u;
                    pb[c] = round16((nb + (bs & (16u / w - 1u))) * w);
                    vb[c] = (NM >> c) & 1 ? round16(nb + (bs & 15u)) : 0u;
                    bytes += pb[c] + vb[c];
                }
                mbar_arrive_expect_tx(&s_bbar, bytes);
#pragma unroll
                for (int c = 0; c < NB; ++c) {
                    const uint32_t w = bwide(c) ? 8u : 4u;
                    const uint32_t sk = bs & (16u / w - 1u);
                    tma_load_1d(smem + a.sm_bpay[c], static_cast<const char*>(a.bpay[c]) + static_cast<uint64_t>(bs - sk) * w, pb[c], &s_bbar);
                    if (vb[c]) tma_load_1d(smem + a.sm_bvalid[c], a.bvalid[c] + (bs - (bs & 15u)), vb[c], &s_bbar);
                }
            }
            for (uint32_t j = 0; j < kUpFront && j < n_batches; ++j) {
                const uint32_t base = ps + j * kBatch;
                issue_probe(batch_no + j, base, pe - base < kBatch ? pe - base : kBatch);
            }
        }
        uint32_t bkey[kBuildItems];
#pragma unroll
        for (int k = 0; k < kBuildItems; ++k) {
            const uint32_t i = k * kThreads + tid;
            bkey[k] = i < nb ? a.bkeys[bs + i] : 0u;
        }
        unsigned long long* const slots = reinterpret_cast<unsigned long long*>(smem); // table first: 32 KB
        bool dup = false;
        if (direct) {
            uint2* const   ent = reinterpret_cast<uint2*>(smem);
            const uint32_t n_words = 1u << rem_words_log;
            for (uint32_t s = tid; s < n_words; s += kThreads) ent[s] = make_uint2(0u, 0u);
            __syncthreads();
#pragma unroll
            for (int k = 0; k < kBuildItems; ++k) {
                if (k * kThreads + tid < nb) {
                    const uint32_t x = hash_key(bkey[k]) >> part_bits;
                    const uint32_t bit = 1u << (x & 31u);
                    if (atomicOr(&ent[x >> 5].x, bit) & bit) dup = true;
                }
            }
            __syncthreads();
            // rank of every word's first bit: thread t scans words [t * per, (t + 1) * per)
            {
                __shared__ uint32_t s_wsum[kWarps];
                const uint32_t per = n_words > kThreads ? n_words / kThreads : 1u; // 8 at 17 bits
                uint32_t mine = 0;
                if (tid * per < n_words)
                    for (uint32_t j = 0; j < per; ++j) mine += __popc(ent[tid * per + j].x);
                uint32_t inc = mine;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t o = __shfl_up_sync(RJ_FULL_MASK, inc, d);
                    if (lane >= static_cast<uint32_t>(d)) inc += o;
                }
                if (lane == 31) s_wsum[tid >> 5] = inc;
                __syncthreads();
                uint32_t run = inc - mine;
#pragma unroll
                for (int w = 0; w < kWarps; ++w) run += w < static_cast<int>(tid >> 5) ? s_wsum[w] : 0u;
                if (tid * per < n_words)
                    for (uint32_t j = 0; j < per; ++j) {
                        ent[tid * per + j].y = run;
                        run += __popc(ent[tid * per + j].x);
                    }
            }
            __syncthreads();
            // the carried build columns, by rank
            if (NB > 0) {
#pragma unroll
                for (int k = 0; k < kBuildItems; ++k) {
                    const uint32_t i = k * kThreads + tid;
                    if (i < nb) {
                        const uint32_t x = hash_key(bkey[k]) >> part_bits;
                        const uint2    e = ent[x >> 5];
                        const uint32_t pos = e.y + __popc(e.x & ((1u << (x & 31u)) - 1u));
                        for_build([&](auto c_c, auto n_c) {
                            constexpr int C = decltype(c_c)::value, N = decltype(n_c)::value;
                            if (bwide(C)) reinterpret_cast<uint64_t*>(smem + a.sm_bpay[C])[pos] = static_cast<const uint64_t*>(a.bpay[C])[bs + i];
                            else reinterpret_cast<uint32_t*>(smem + a.sm_bpay[C])[pos] = static_cast<const uint32_t*>(a.bpay[C])[bs + i];
                            if constexpr (N >= 0) (smem + a.sm_bvalid[C])[pos] = a.bvalid[C][bs + i];
                        });
                    }
                }
            }
        } else {
            for (uint32_t s = tid; s < kSlots; s += kThreads) slots[s] = ~0ull;
            __syncthreads();
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
        }
        if (__syncthreads_or(dup ? 1 : 0)) {
            // not a key / foreign-key join: leave it to the general path (outstanding bulk copies land in this
            // CTA's shared memory before it retires)
            if (tid == 0) atomicExch(a.abort_flag, 1u);
            if (NB > 0 && !direct) mbar_wait(&s_bbar, unit_no & 1);
            for (uint32_t j = 0; j < kUpFront && j < n_batches; ++j) mbar_wait(&s_full[(batch_no + j) & (kStages - 1)], ((batch_no + j) / kStages) & 1);
            return;
        }
        if (NB > 0 && !direct) {
            mbar_wait(&s_bbar, unit_no & 1);
            ++unit_no;
        }

        // ---- probe ------------------------------------------------------------------------------------------
        if (producer) {
            // the data: every further batch as soon as its buffer is free; then the next unit
            if (lane == 0) {
                for (uint32_t j = kUpFront; j < n_batches; ++j) {
                    const uint32_t base = ps + j * kBatch;
                    issue_probe(batch_no + j, base, pe - base < kBatch ? pe - base : kBatch);
                }
                lookup_unit(s_desc[(round + 1) & 1]);
            }
            __syncwarp();
            batch_no += n_batches;
            continue;
        }
        if ((tid >> 5) >= a.n_active) { // a probe side too small to give every warp a few chunks of rows: this warp idles
            batch_no += n_batches;
            continue;
        }
        for (uint32_t j = 0; j < n_batches; ++j, ++batch_no) {
            const uint32_t base = ps + j * kBatch;
            const uint32_t cnt = pe - base < kBatch ? pe - base : kBatch;
            const uint32_t stage = batch_no & (kStages - 1);
            mbar_wait_addr(sbase - kBarBytes + 8u + stage * 8u, (batch_no / kStages) & 1);
            const uint32_t        buf = sbase + stage * a.sm_pstride; // shared-memory address of the batch's buffer
            // an ITEM = 32 consecutive tuples of the batch; the active warps share the batch's kConsumers items
            auto process_item = [&](const uint32_t i) {
                const uint32_t        key = lds_u32(buf + a.sm_pkeys + ((base & 3u) + i) * 4u); // past cnt: stale bytes, never used
                uint32_t lidx = kNone;
                if (direct) {
                    // the remaining hash bits identify the key: one load, a bit test, a popcount
                    const uint32_t x = hash_key(key) >> part_bits;
                    const uint2    e = lds_v2(sbase + (x >> 5) * 8u);
                    const uint32_t below = e.x & ((1u << (x & 31u)) - 1u);
                    if (i < cnt && ((e.x >> (x & 31u)) & 1u)) lidx = e.y + __popc(below);
                } else if (i < cnt) {
                    // at most one match: the table holds distinct keys.  A slot is (key, build index); an empty
                    // one has index 0xffffffff
                    uint32_t       off   = ((hash_key(key) >> part_bits) & kSlotMask) * 8u;
                    const uint32_t step8 = probe_step(key) * 8u;
                    for (;;) {
                        uint32_t ex, ey;
                        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(ex), "=r"(ey) : "r"(smem_u32(smem) + off));
                        if (ey == kNone) break;
                        if (ex == key) {
                            lidx = ey;
                            break;
                        }
                        off = (off + step8) & (kSlotMask * 8u);
                    }
                }
                __syncwarp();
                bool           act = lidx != kNone;
                const uint32_t bal = __ballot_sync(RJ_FULL_MASK, act);
                if (bal != 0) {
                    // the row's validity in the nullable columns, and its values
                    bool ok[NNX];
                    uint64_t bval[NB > 0 ? NB : 1], pval[NP > 0 ? NP : 1];
                    for_build([&](auto c_c, auto n_c) {
                        constexpr int C = decltype(c_c)::value, N = decltype(n_c)::value;
                        const uint32_t at = act ? lidx : 0u;
                        if constexpr (N >= 0) ok[N] = act && lds_u8(sbase + a.sm_bvalid[C] + bskew1 + at) != 0;
                        if (bwide(C)) bval[C] = lds_u64(sbase + a.sm_bpay[C] + ((bskew1 & 1u) + at) * 8u);
                        else bval[C] = lds_u32(sbase + a.sm_bpay[C] + ((bskew1 & 3u) + at) * 4u);
                    });
                    for_probe([&](auto c_c, auto n_c) {
                        constexpr int C = decltype(c_c)::value, N = decltype(n_c)::value;
                        if constexpr (N >= 0) ok[N] = act && lds_u8(buf + a.sm_pvalid[C] + (base & 15u) + i) != 0;
                        if (pwide(C)) pval[C] = lds_u64(buf + a.sm_ppay[C] + ((base & 1u) + i) * 8u);
                        else pval[C] = lds_u32(buf + a.sm_ppay[C] + ((base & 3u) + i) * 4u);
                    });

                    // place the matched rows: row index in the open chunk; rows past its end go to the next one (a second
                    // round)
                    uint32_t r = rows + __popc(bal & lt);
                    uint32_t left = __popc(bal);
                    for (;;) {
                        ensure_open();
                        if (all_once && left == 32u && rows + 32u <= (rows < kHalfRows ? kHalfRows : kChunkRows)) {
                            // DENSE item (every lane matched -- the rule in a key / foreign-key join) that lies inside one page
                            // of every column: rows are the lanes in order and complete exactly one bitmap word.  Everything
                            // but the value slots is warp-uniform.  (The two items of a chunk that straddle row 992 or 1984,
                            // and items with unmatched lanes, take the general round below.)
                            const bool     h = rows >= kHalfRows;
                            const uint32_t w0 = rows >> 5, sh = rows & 31u;
                            auto dense = [&](auto n_c, bool wide, uint64_t val, uint8_t* cb) {
                                constexpr int N = decltype(n_c)::value;
                                uint32_t at;
                                bool     st = true;
                                if constexpr (N >= 0) {
                                    const uint32_t okb = __ballot_sync(RJ_FULL_MASK, ok[N]);
                                    st = ok[N];
                                    at = nv[N] + __popc(okb & lt);
                                    if (wide && h) at += RJ_PAGE / 8 - nvh[N];
                                    nv[N] += __popc(okb);
                                    if (lane == (w0 & 31u)) held[N] = pend[N] | (okb << sh);
                                    pend[N] = sh ? okb >> (32u - sh) : 0u;
                                    const uint32_t first_flush = wide ? kHalfWords - 1 : 31u;
                                    if (w0 == first_flush || w0 == kChunkRows / 32 - 1) {
                                        const uint32_t wl = w0 - ((w0 - lane) & 31u);
                                        if (wl <= w0 && (w0 == first_flush || wl > first_flush)) *bitmap_word(cb, wide, wl) = held[N];
                                    }
                                } else {
                                    at = rows + lane;
                                    if (wide && h) at += RJ_PAGE / 8 - kHalfRows;
                                }
                                if (wide) {
                                    if (st) (reinterpret_cast<uint64_t*>(cb + 8))[at] = val;
                                } else {
                                    if (st) (reinterpret_cast<uint32_t*>(cb + 4))[at] = static_cast<uint32_t>(val);
                                }
                            };
                            dense(std::integral_constant<int, -1>{}, false, key, cb_key);
                            for_build([&](auto c_c, auto n_c) { dense(n_c, bwide(decltype(c_c)::value), bval[decltype(c_c)::value], cb_b[decltype(c_c)::value]); });
                            for_probe([&](auto c_c, auto n_c) { dense(n_c, pwide(decltype(c_c)::value), pval[decltype(c_c)::value], cb_p[decltype(c_c)::value]); });
                            rows += 32u;
                            left = 0;
                            if (NN > 0 && rows == kHalfRows) { // the 8-byte columns' first page is complete
#pragma unroll
                                for (int nn = 0; nn < NN; ++nn) nvh[nn] = nv[nn];
                            }
                        } else {
                            const bool     now = act && r < kChunkRows;
                            const uint32_t nowb = __ballot_sync(RJ_FULL_MASK, now);
                            const uint32_t n_now = __popc(nowb);
                            uint32_t okb[NNX], v[NNX];
#pragma unroll
                            for (int nn = 0; nn < NN; ++nn) {
                                okb[nn] = __ballot_sync(RJ_FULL_MASK, now && ok[nn]);
                                v[nn]   = nv[nn] + __popc(okb[nn] & lt);
                            }
                            if (NN > 0 && rows < kHalfRows && rows + n_now >= kHalfRows) {
                                // the rows of this round reach the second page of the 8-byte columns
                                const uint32_t lowb = __ballot_sync(RJ_FULL_MASK, now && r < kHalfRows);
#pragma unroll
                                for (int nn = 0; nn < NN; ++nn) nvh[nn] = nv[nn] + __popc(okb[nn] & lowb);
                            }
                            const uint32_t w0 = rows >> 5, sh = rows & 31u; // the bitmap word the round's first row falls into
                            // One source's value into every output column that shows it, and -- for a nullable source -- the
                            // validity bits of the rows placed now (rows [rows, rows + n_now) of the chunk: two bitmap words
                            // at most).  A complete word stays with lane (w & 31) until 31 / 32 of them (one page's worth, or
                            // the chunk's second half) go out in one store.  N = index among the nullable columns, -1: the
                            // source holds no NULL.  kOnce: every source is shown by exactly one output column, whose chunk
                            // base is cb.
                            auto emit = [&](auto once_c, auto n_c, int S, bool wide, uint64_t val, uint8_t* cb) {
                                constexpr int  N = decltype(n_c)::value;
                                constexpr bool kOnce = decltype(once_c)::value;
                                bool st = now;
                                if constexpr (N >= 0) st = now && ok[N];
                                uint32_t at; // value slot, in units of the value width, from the chunk's first data byte
                                if (wide) {
                                    const bool h = r >= kHalfRows;
                                    if constexpr (N >= 0) at = v[N] - (h ? nvh[N] : 0u); else at = r - (h ? kHalfRows : 0u);
                                    at += h ? RJ_PAGE / 8 : 0u;
                                } else {
                                    if constexpr (N >= 0) at = v[N]; else at = r;
                                }
                                bool flush = false;
                                if constexpr (N >= 0) {
                                    uint32_t lo, hi;
                                    if (nowb == RJ_FULL_MASK) { // rows are the lanes in order
                                        lo = okb[N] << sh;
                                        hi = sh ? okb[N] >> (32u - sh) : 0u;
                                    } else {
                                        const uint32_t bit = st ? (1u << (r & 31u)) : 0u;
                                        const bool     first = (r >> 5) == w0;
                                        lo = __reduce_or_sync(RJ_FULL_MASK, first ? bit : 0u);
                                        hi = __reduce_or_sync(RJ_FULL_MASK, first ? 0u : bit);
                                    }
                                    const uint32_t word = pend[N] | lo;
                                    const bool     complete = sh + n_now >= 32u;
                                    pend[N] = complete ? hi : word;
                                    if (complete && lane == (w0 & 31u)) held[N] = word;
                                    flush = complete && (w0 == (wide ? kHalfWords - 1 : 31u) || w0 == kChunkRows / 32 - 1);
                                }
                                auto store = [&](uint8_t* base_page) {
                                    if (wide) {
                                        if (st) (reinterpret_cast<uint64_t*>(base_page + 8))[at] = val;
                                    } else {
                                        if (st) (reinterpret_cast<uint32_t*>(base_page + 4))[at] = static_cast<uint32_t>(val);
                                    }
                                    if constexpr (N >= 0) {
                                        if (flush) {
                                            // lane l holds word wl = the largest w <= w0 with (w & 31) == l; the words up to
                                            // the first flush point went out then
                                            const uint32_t wl = w0 - ((w0 - lane) & 31u);
                                            const uint32_t first_flush = wide ? kHalfWords - 1 : 31u;
                                            const bool     mine = wl <= w0 && (w0 == first_flush || wl > first_flush); // (wl wraps above w0 when no such word exists)
                                            if (mine) *bitmap_word(base_page, wide, wl) = held[N];
                                        }
                                    }
                                };
                                if constexpr (kOnce) store(cb);
                                else for_outputs(S, wide, c_open, store);
                            };
                            auto emit_all = [&](auto once_c) {
                                emit(once_c, std::integral_constant<int, -1>{}, 0, false, key, cb_key);
                                for_build([&](auto c_c, auto n_c) { emit(once_c, n_c, 1 + decltype(c_c)::value, bwide(decltype(c_c)::value), bval[decltype(c_c)::value], cb_b[decltype(c_c)::value]); });
                                for_probe([&](auto c_c, auto n_c) { emit(once_c, n_c, 1 + kEmitMaxPay + decltype(c_c)::value, pwide(decltype(c_c)::value), pval[decltype(c_c)::value], cb_p[decltype(c_c)::value]); });
                            };
                            if (all_once) emit_all(std::true_type{}); else emit_all(std::false_type{});
                            rows += n_now;
                            left -= n_now;
#pragma unroll
                            for (int nn = 0; nn < NN; ++nn) nv[nn] += __popc(okb[nn]);
                            act = act && !now;
                            r -= kChunkRows; // (meaningful for the rows of a second round only)
                        }
                        after_place();
                        if (left == 0) break;
                    }
                }
            };
            if (all_active) process_item(tid);
            else for (uint32_t i = tid; i < kBatch; i += a.n_active * 32u) process_item(i);
            // this warp is done with the batch's buffer
            __syncwarp();
            if (lane == 0) mbar_arrive_addr(sbase - kBarBytes + 8u + (kStages + stage) * 8u);
        }
    }
    // the partly filled chunk of this warp, and the chunk it may hold in reserve (an empty one)
    if (c_open != kNone) write_meta(c_open, rows);
    if (have_next) {
        const uint32_t c = __shfl_sync(RJ_FULL_MASK, c_next, 0);
#pragma unroll
        for (int nn = 0; nn < NNX; ++nn) nv[nn] = nvh[nn] = pend[nn] = held[nn] = 0u;
        write_meta(c, 0u);
    }
    if (lane == 0) {
        const unsigned long long total = static_cast<unsigned long long>(closed) * kChunkRows + (c_open != kNone ? rows : 0u);
        if (total) atomicAdd(a.row_counter, total);
    }
}

// one instantiation's launch: its attributes (dynamic shared-memory limit, carve-out) are set once per device.
// (The statics must be per INSTANTIATION: every join_emit_kernel<...> has the same function-pointer type, so a
// generic lambda taking the pointer would share one set of statics among all of them.)
template <int NB, int NP, int NM, int WM>
static void launch_emit_instance(const EmitArgs& a, size_t smem, unsigned grid, cudaStream_t s) {
    auto kern = join_emit_kernel<NB, NP, NM, WM>;
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

// Every (probe columns, nullable columns) combination of one build-column count; false: no instantiation
// matches.  With at most one carried column per side (configs 1 and 2, every key / foreign-key join with one
// payload per table) the column widths are compile-time constants as well; wider shapes read them at run time.
template <int NB>
static bool launch_emit_nb(const EmitArgs& a, int n_ppay, int null_mask, int width_mask, size_t smem, unsigned grid, cudaStream_t s) {
    bool launched = false;
    auto try_launch = [&](auto np_c, auto nm_c) {
        constexpr int P = decltype(np_c)::value, M = decltype(nm_c)::value;
        constexpr int allowed = ((1 << NB) - 1) | (((1 << P) - 1) << kEmitMaxPay);
        if constexpr ((M & ~allowed) == 0) {
            if (!launched && n_ppay == P && null_mask == M) {
                if constexpr (NB <= 1 && P <= 1) {
                    switch (width_mask) {
                    case 0: launch_emit_instance<NB, P, M, 0>(a, smem, grid, s); break;
                    case 1: if constexpr (NB == 1) launch_emit_instance<NB, P, M, 1>(a, smem, grid, s); break;
                    case 4: if constexpr (P == 1) launch_emit_instance<NB, P, M, 4>(a, smem, grid, s); break;
                    case 5: if constexpr (NB == 1 && P == 1) launch_emit_instance<NB, P, M, 5>(a, smem, grid, s); break;
                    default: return;
                    }
                } else {
                    launch_emit_instance<NB, P, M, -1>(a, smem, grid, s);
                }
                launched = true;
            }
        }
    };
    auto for_nm = [&](auto np_c) {
#ifdef RJ_EMIT_DEV_ONLY  // development builds: only config 2's shape (see Makefile)
        try_launch(np_c, std::integral_constant<int, 5>{});
#else
        try_launch(np_c, std::integral_constant<int, 0>{});  try_launch(np_c, std::integral_constant<int, 1>{});
        try_launch(np_c, std::integral_constant<int, 2>{});  try_launch(np_c, std::integral_constant<int, 3>{});
        try_launch(np_c, std::integral_constant<int, 4>{});  try_launch(np_c, std::integral_constant<int, 5>{});
        try_launch(np_c, std::integral_constant<int, 6>{});  try_launch(np_c, std::integral_constant<int, 7>{});
        try_launch(np_c, std::integral_constant<int, 8>{});  try_launch(np_c, std::integral_constant<int, 9>{});
        try_launch(np_c, std::integral_constant<int, 10>{}); try_launch(np_c, std::integral_constant<int, 11>{});
        try_launch(np_c, std::integral_constant<int, 12>{}); try_launch(np_c, std::integral_constant<int, 13>{});
        try_launch(np_c, std::integral_constant<int, 14>{}); try_launch(np_c, std::integral_constant<int, 15>{});
#endif
    };
    for_nm(std::integral_constant<int, 0>{});
    for_nm(std::integral_constant<int, 1>{});
    for_nm(std::integral_constant<int, 2>{});
    return launched;
}

} // namespace emit

// defined in k_join_emit.cu / k_join_emit_b1.cu / k_join_emit_b2.cu
bool launch_join_emit_b0(const emit::EmitArgs& a, int n_ppay, int null_mask, int width_mask, size_t smem, unsigned grid, cudaStream_t s);
bool launch_join_emit_b1(const emit::EmitArgs& a, int n_ppay, int null_mask, int width_mask, size_t smem, unsigned grid, cudaStream_t s);
bool launch_join_emit_b2(const emit::EmitArgs& a, int n_ppay, int null_mask, int width_mask, size_t smem, unsigned grid, cudaStream_t s);

} // namespace rj
