// Host-side internal interface of the engine: the context, the launch wrappers of every kernel
// file, and small RAII helpers.  Nothing here is exported; the C-ABI lives in engine.cu.
#pragma once
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/rj_b200.h"

// bits of the context's device-visible error word (malformed input detected by a kernel)
#define RJ_ERR_ORPHAN_LONG_PAGE 1u

namespace rj {

struct CudaError: std::runtime_error {
    using std::runtime_error::runtime_error;
};

#define RJ_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t rj_e_ = (expr);                                                                \
        if (rj_e_ != cudaSuccess) {                                                                \
            throw ::rj::CudaError(std::string(#expr) + ": " + cudaGetErrorString(rj_e_) + " (" +   \
                                  __FILE__ + ":" + std::to_string(__LINE__) + ")");                \
        }                                                                                          \
    } while (0)

// every kernel launch of the engine goes through this macro; the counter backs rj_kernel_launch_count()
extern std::atomic<uint64_t> g_kernel_launches;
#define RJ_LAUNCH_CHECK()                                  \
    do {                                                   \
        ::rj::g_kernel_launches.fetch_add(1, std::memory_order_relaxed); \
        RJ_CUDA(cudaGetLastError());                       \
    } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per DEVICE: remember the largest size configured on
// each one (a process may hold contexts on several GPUs)
struct SmemConfigured {
    size_t bytes[64] = {};
    // raise the kernel's dynamic shared-memory limit to `want` if it is below; the new size is recorded
    // only once the attribute call has succeeded, so a refused request does not poison later ones
    template <class Kern>
    void ensure(Kern kern, size_t want) {
        int dev = 0;
        cudaGetDevice(&dev);
        size_t& have = bytes[dev & 63];
        if (want <= have) return;
        RJ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(want)));
        have = want;
    }
};

// ---- join geometry (shared by the partition planner and the join kernel) ---------------------------
constexpr uint32_t kJoinSlots      = 8192;  // shared-memory hash table slots per CTA
constexpr uint32_t kJoinBuildCap   = 6144;  // build tuples per table (75 % fill); larger partitions are chunked
constexpr uint32_t kJoinTargetFill = 2048;  // partition fan-out aims at <= this many build tuples on average (25 % fill)
constexpr uint32_t kJoinProbeChunk = 16384; // probe tuples per work unit
// tuples per scatter tile: 256 threads x 16 (4-byte keys) or x 8 (8-byte keys)
constexpr uint32_t scatter_tile(int key_bytes) { return key_bytes == 4 ? 4096u : 2048u; }
constexpr int      kMaxPassBits    = 8;     // radix bits per scatter pass
constexpr int      kMaxTotalBits   = 15;    // 2^15 histogram bins = 128 KB of shared memory (2^16 would not fit one CTA)

// Device-resident partition plan, filled by launch_partition_plan (single small kernel, no host sync)
struct PartitionPlanDev {
    uint32_t* off_b;        // [nparts+1] final build offsets (exclusive prefix of the histogram)
    uint32_t* off_p;        // [nparts+1]
    uint32_t* cur_b;        // [nparts]   pass-2 (or single-pass) cursors, advanced by the scatter
    uint32_t* cur_p;        // [nparts]
    uint32_t* reg_b;        // [nreg+1]   pass-1 region starts (two-pass only)
    uint32_t* reg_p;        // [nreg+1]
    uint32_t* cur1_b;       // [nreg]     pass-1 cursors
    uint32_t* cur1_p;       // [nreg]
    uint32_t* tile_b;       // [nreg+1]   exclusive prefix of pass-2 tiles per region
    uint32_t* tile_p;       // [nreg+1]
    uint32_t* unit_start;   // [nparts+1] exclusive prefix of join work units per partition
    uint32_t* unit_cursor;  // [1]        dynamic work distribution of the join kernel
    uint32_t* unit_part;    // [2 * nparts] partition of work unit u, for the first 2 * nparts units (the others: search unit_start)
};

size_t partition_plan_words(int total_bits, int pass1_bits);
void   partition_plan_carve(uint32_t* base, int total_bits, int pass1_bits, PartitionPlanDev* plan);

// ---- host_copy.cpp ---------------------------------------------------------------------------------
void copy_page(void* dst, const void* src); // 8192 bytes, non-temporal stores when dst is 16-byte aligned
void copy_fence();                          // before another agent reads what copy_page wrote

// ---- k_scan.cu ------------------------------------------------------------------------------------
size_t scan_tmp_bytes(uint64_t n);
// out[i] = sum_{k<i} in[k], out[n] = total
void launch_exclusive_scan_u32_u64(const uint32_t* in, uint64_t* out, uint64_t n, void* tmp, cudaStream_t s);
// inclusive scans over uint64 (sum or max)
void launch_inclusive_sum_u64(const uint64_t* in, uint64_t* out, uint64_t n, void* tmp, cudaStream_t s);
void launch_inclusive_max_u64(const uint64_t* in, uint64_t* out, uint64_t n, void* tmp, cudaStream_t s);

// ---- k_decode.cu ----------------------------------------------------------------------------------
void launch_page_rows(const void* pages, uint64_t n_pages, int type, uint32_t* rows, uint64_t* totals,
                      cudaStream_t s);
void launch_decode_fixed(const void* pages, uint64_t n_pages, int type, const uint64_t* row_start,
                         void* values, uint32_t* valid, int sm_count, cudaStream_t s);
// err_flags (may be NULL): device-visible word that collects RJ_ERR_* bits of malformed input
void launch_decode_varchar(const void* pages, uint64_t n_pages, const uint64_t* row_start, uint64_t* desc,
                           uint32_t* valid, int sm_count, cudaStream_t s, uint32_t* err_flags = nullptr);

// ---- k_partition.cu -------------------------------------------------------------------------------
void launch_radix_histogram(const void* keys, const uint32_t* valid, uint64_t n, int key_bytes, int shift,
                            int bits, uint32_t* hist, int sm_count, cudaStream_t s);
// build_cap x probe_chunk: tuples per join work unit (kJoinBuildCap x kJoinProbeChunk for join_kernel,
// kEmitBuildCap x kEmitProbeChunk for join_emit_kernel)
void launch_partition_plan(const uint32_t* hist_b, const uint32_t* hist_p, uint32_t flat_b, uint32_t flat_p,
                           int total_bits, int pass1_bits, int key_bytes, const PartitionPlanDev& plan,
                           cudaStream_t s, uint32_t build_cap = kJoinBuildCap, uint32_t probe_chunk = kJoinProbeChunk);
// Payload columns that travel with the tuples of a flat scatter: dst[c][pos] = src[c][row of the tuple].
// The source is read inside the tile's own row window (sequential DRAM traffic), so a later gather
// through positions of the scattered order stays inside one partition / region instead of the table.
struct ScatterPayload {
    static constexpr int kMax = 6;
    int         n = 0;
    const void* src[kMax] = {};
    void*       dst[kMax] = {};
    int         width[kMax] = {}; // 4 or 8 bytes; 1 = src is a validity BITMAP, dst one byte per tuple
};
// flat scatter over [0, n): cursor index = radix digit
void launch_radix_scatter(const void* keys, const uint32_t* valid, const uint32_t* idx_in, uint64_t n,
                          int key_bytes, int shift, int bits, uint32_t* cursor, void* keys_out,
                          uint32_t* idx_out, const ScatterPayload& payload, int sm_count, cudaStream_t s);
// flat scatter with one output base per partition (<= 8) -- see rj_radix_scatter_multi
void launch_radix_scatter_multi(const void* keys, const uint32_t* valid, uint64_t n, int key_bytes, int shift, int bits,
                                uint32_t* cursor, const rj_scatter_multi_t& out, int sm_count, cudaStream_t s);
// segmented scatter (pass 2): region r covers [region_start[r], region_start[r+1]) of the input,
// tiles are enumerated through tile_start, cursor index = (r << bits) | digit
// Flags folded into the row index written by the segmented scatter: bit (30 + c) of idx_out is set
// iff flag_src[c][i] != 0 for input tuple i (carried validity bytes; indices must stay below 2^30).
struct RegionFlags {
    int            n = 0;
    const uint8_t* src[2] = {nullptr, nullptr};
    uint8_t*       dst[2] = {nullptr, nullptr}; // optional: the bytes again, in the order this pass produces
};
constexpr uint32_t kPosMask = 0x3fffffffu;
void launch_radix_scatter_regions(const void* keys, const uint32_t* idx_in, const uint32_t* region_start,
                                  const uint32_t* tile_start, uint32_t n_regions, uint64_t n_upper,
                                  int key_bytes, int shift, int bits, uint32_t* cursor, void* keys_out,
                                  uint32_t* idx_out, const RegionFlags& flags, const ScatterPayload& payload, int sm_count,
                                  cudaStream_t s);

// ---- k_scatter_carry.cu: scatter of whole tuples (4-byte key + the columns the root outputs) ------------
// Flat pass: region_start == NULL; `valid` = key validity bitmap (tuples with a NULL key are dropped),
// flag_src = validity BITMAPS of the carried columns by row.  Region pass (second pass): the input is the
// flat pass's output, flag_src = validity BYTES by position, tiles never straddle a region.  flag_dst is
// one byte per tuple in both.  Keys and value sources must be 16-byte aligned (1-D TMA).
struct CarryScatter {
    const uint32_t* keys = nullptr;
    const uint32_t* valid = nullptr;
    uint64_t        n = 0; // tuples (region pass: an upper bound, the exact count is region_start[n_regions])
    const uint32_t* region_start = nullptr;
    const uint32_t* tile_start = nullptr;
    uint32_t        n_regions = 0;
    int             shift = 0, bits = 0;
    uint32_t*       cursor = nullptr;
    uint32_t*       keys_out = nullptr;
    int             n_val = 0;
    const void*     val_src[2] = {nullptr, nullptr};
    void*           val_dst[2] = {nullptr, nullptr};
    int             val_width[2] = {0, 0}; // 4 or 8
    int             n_flag = 0;
    const void*     flag_src[2] = {nullptr, nullptr};
    uint8_t*        flag_dst[2] = {nullptr, nullptr};
    // multi-GPU exchange (flat pass only): n_owners > 0 -> partition p goes to owner p >> owner_shift, whose
    // arrays (peer-mapped) are listed here; cursor[p] indexes the owner's arrays; keys_out / val_dst / flag_dst unused
    int             n_owners = 0;
    int             owner_shift = 0;
    void*           keys_dst_multi[8] = {};
    void*           val_dst_multi[2][8] = {};
    void*           flag_dst_multi[2][8] = {};
    // region pass over scattered inputs (multi-GPU pull): per-region source addresses (see k_scatter_carry.cu)
    // and the cursor group of every region; keys / val_src / flag_src are then unused
    const uint64_t* src_tab = nullptr;      // [n_regions][5] device
    const uint32_t* region_group = nullptr; // [n_regions] device
};
void launch_scatter_carry(const CarryScatter& c, int sm_count, cudaStream_t s);

void launch_dist_layout(const uint32_t* H, int me, int g, int bits, int p1, const uint64_t* ptrs, const int* widths, uint32_t* cursor,
                        uint64_t* table, uint32_t* start, uint32_t* tile, uint32_t* group, uint32_t* local_hist,
                        unsigned long long* scalars, cudaStream_t s); // k_filter.cu
void launch_varchar_desc_from_offsets(const uint64_t* off, uint64_t n, uint64_t* desc, int sm_count, cudaStream_t s); // k_varchar.cu

// ---- k_filter.cu: predicates on decoded columns, bitmap algebra, bitmap -> row ids -------------------------
void launch_filter_compare(const void* values, const uint32_t* valid, uint64_t n, int type, int op, int64_t rhs_i, double rhs_d,
                           uint32_t* out, int sm_count, cudaStream_t s);
void launch_filter_null(const uint32_t* valid, uint64_t n, bool want_null, uint32_t* out, int sm_count, cudaStream_t s);
void launch_bitmap_logic(const uint32_t* a, const uint32_t* b, uint64_t n, int op, uint32_t* out, int sm_count, cudaStream_t s);
void launch_filter_varchar(const void* pages, const uint64_t* desc, const uint32_t* valid, uint64_t n, int op, const uint8_t* d_rhs,
                           uint32_t rhs_len, uint32_t* out, int sm_count, cudaStream_t s);
void launch_bitmap_popc(const uint32_t* bits, uint64_t n_words, uint32_t* counts, int sm_count, cudaStream_t s);
void launch_bitmap_expand(const uint32_t* bits, const uint64_t* start, uint64_t n_words, uint32_t* ids, int sm_count, cudaStream_t s);

// ---- k_sort.cu: result validation (row hashes, LSD radix sort, pairwise comparison) ----------------------
void launch_hash_fixed_cells(const void* values, const uint32_t* valid, uint64_t n, int width, uint64_t* h, int sm_count, cudaStream_t s);
void launch_hash_combine(const uint64_t* cell, const uint32_t* valid, uint64_t n, uint64_t* h, int sm_count, cudaStream_t s);
uint64_t sort_tmp_words(uint64_t n);
void launch_radix_sort_u64(uint64_t* keys, uint32_t* vals, uint64_t* alt_keys, uint32_t* alt_vals, uint64_t n, uint32_t* counts,
                           uint64_t* base, void* scan_tmp, cudaStream_t s);
void launch_iota_u32(uint32_t* out, uint64_t n, int sm_count, cudaStream_t s);
void launch_pairs_equal_fixed(const void* va, const uint32_t* valid_a, const uint32_t* idx_a, const void* vb, const uint32_t* valid_b,
                              const uint32_t* idx_b, uint64_t n, int width, unsigned long long* mismatches, int sm_count, cudaStream_t s);
void launch_pairs_equal_varchar_finish(const uint32_t* valid_a, const uint32_t* idx_a, const uint32_t* valid_b, const uint32_t* idx_b,
                                       const uint32_t* keep, uint64_t n, unsigned long long* mismatches, int sm_count, cudaStream_t s);
void launch_keys_differ(const uint64_t* a, const uint64_t* b, uint64_t n, unsigned long long* mismatches, int sm_count, cudaStream_t s);

// ---- k_join.cu ------------------------------------------------------------------------------------
struct JoinLaunch {
    const void*     bkeys;
    const uint32_t* bidx;   // NULL = identity
    const uint32_t* bvalid; // NULL = all valid
    const void*     pkeys;
    const uint32_t* pidx;
    const uint32_t* pvalid;
    const uint32_t* off_b;  // [nparts+1]
    const uint32_t* off_p;
    const uint32_t* unit_start; // [nparts+1]
    uint32_t        probe_chunk; // probe tuples per work unit (0: kJoinProbeChunk); must match the partition plan's
    uint32_t*       unit_cursor; // device word, zeroed before every launch: next unit to hand out
    uint32_t        nparts;
    int             part_bits;  // hash bits consumed by the partitioning
    int             key_bytes;
    uint32_t*       out_b;
    uint32_t*       out_p;
    uint64_t        capacity;
    unsigned long long* out_count; // device counter (zeroed by the caller)
    // scratch for the duplicate chains of 4-byte keys (k_join.cu): no initialisation needed
    uint32_t*       dup_next;      // [join_grid(sm_count) * kJoinBuildCap]
    uint32_t*       dup_head;      // [join_grid(sm_count) * kJoinSlots]
};
inline unsigned join_grid(int sm_count) { return static_cast<unsigned>(sm_count) * 2; }
void launch_join(const JoinLaunch& a, int sm_count, cudaStream_t s);

// ---- k_join_emit.cu: root join fused with page output ------------------------------------------------
constexpr uint32_t kEmitSlots     = 4096;  // table slots per CTA (two CTAs of ~100 KB per SM)
constexpr uint32_t kEmitBuildCap  = 3072;  // build tuples per table; larger partitions are chunked
constexpr uint32_t kEmitProbeChunk = 65536; // probe tuples per work unit (a table is built once per unit)
constexpr uint32_t kEmitChunkRows = 1984;  // rows per output chunk: 1 page per 4-byte column, 2 x 992 rows per 8-byte column
constexpr int      kEmitMaxPay    = 2;     // carried columns per side
constexpr int      kEmitMaxOut    = 4;     // output columns
constexpr uint32_t kEmitWarps = 16;            // warps per CTA of join_emit_kernel: every warp owns an open chunk
constexpr uint32_t kEmitMinChunksPerWarp = 16; // an emitting warp should fill this many chunks before it leaves one partly filled (<= 3 % more pages): results that travel over PCIe
constexpr uint32_t kEmitMinChunksResident = 8; // ... when the result pages stay in HBM (<= 6 %): more warps emit (N = 8: 14 instead of 7 per CTA)
inline uint64_t join_emit_env(const char* name, uint64_t dflt) { // tuning knobs (profiling)
    const char* e = getenv(name);
    return e && atoll(e) > 0 ? static_cast<uint64_t>(atoll(e)) : dflt;
}
struct JoinEmitLaunch {
    const uint32_t* bkeys = nullptr; // both sides fully partitioned, 4-byte keys, NULL keys already dropped
    const uint32_t* pkeys = nullptr;
    const uint32_t* off_b = nullptr; // [nparts+1]
    const uint32_t* off_p = nullptr;
    const uint32_t* unit_start = nullptr; // [nparts+1], units of (kEmitBuildCap build) x (kEmitProbeChunk probe) tuples
    uint32_t*       unit_cursor = nullptr;
    const uint32_t* unit_part = nullptr;  // [2 * nparts] (PartitionPlanDev::unit_part), may be NULL
    uint32_t        nparts = 1;
    int             part_bits = 0;
    int             n_bpay = 0, n_ppay = 0;
    const void*     bpay[kEmitMaxPay] = {};
    const uint8_t*  bvalid[kEmitMaxPay] = {};
    int             bwidth[kEmitMaxPay] = {};
    const void*     ppay[kEmitMaxPay] = {};
    const uint8_t*  pvalid[kEmitMaxPay] = {};
    int             pwidth[kEmitMaxPay] = {};
    int             n_out = 0;
    int             out_src[kEmitMaxOut] = {};      // 0 = join key, 1 = build payload, 2 = probe payload
    int             out_idx[kEmitMaxOut] = {};
    int             out_width[kEmitMaxOut] = {};
    int             out_nullable[kEmitMaxOut] = {};
    uint8_t*        out_pages[kEmitMaxOut] = {};    // room for join_emit_max_chunks() chunks each
    uint32_t*           chunk_counter = nullptr;    // zeroed by the caller
    unsigned long long* row_counter = nullptr;
    uint32_t*           abort_flag = nullptr;
    uint32_t            min_chunks_per_warp = kEmitMinChunksPerWarp; // see join_emit_active_warps
};
// Two CTAs per SM, or one per partition when there are fewer partitions than that (a partition is at least one
// work unit; CTAs without a unit retire at once).  How many chunks are left partly filled is governed by the
// number of warps that emit, not by the grid: see join_emit_active_warps.
inline unsigned join_emit_grid(uint64_t n_probe, uint64_t n_parts, int sm_count) {
    (void)n_probe;
    const uint64_t cap = static_cast<uint64_t>(sm_count) * 2;
    return static_cast<unsigned>(n_parts < 1 ? 1 : (n_parts > cap ? cap : n_parts));
}
// Warps per CTA that probe and emit (the others idle): as many as get kEmitMinChunksPerWarp chunks' worth of probe
// tuples each, so that the partly filled chunk every emitting warp ends with stays a small fraction (<= ~3 %) of
// the result's pages.  A small probe side is better served by a few warps on EVERY SM than by all warps of a few
// CTAs (config 1, 10 M probe tuples: 0.39 ms per execute with the former, 1.2 ms with the latter).
inline uint32_t join_emit_active_warps(uint64_t n_probe, unsigned grid, uint32_t min_chunks_per_warp = kEmitMinChunksPerWarp) {
    static const uint64_t forced = join_emit_env("RJ_EMIT_MIN_CHUNKS", 0);
    const uint64_t min_chunks = forced ? forced : (min_chunks_per_warp ? min_chunks_per_warp : kEmitMinChunksPerWarp);
    const uint64_t per_warp = uint64_t(kEmitChunkRows) * min_chunks * grid;
    const uint64_t w = (n_probe + per_warp - 1) / per_warp;
    return static_cast<uint32_t>(w < 1 ? 1 : (w > kEmitWarps - 1 ? kEmitWarps - 1 : w));
}
// chunks a launch can produce: every probe tuple matches at most once; every warp ends with at most one partly
// filled chunk and one (empty) chunk held in reserve
inline uint64_t join_emit_max_chunks(uint64_t n_probe, uint64_t n_parts, int sm_count) {
    return n_probe / kEmitChunkRows + 2ull * kEmitWarps * join_emit_grid(n_probe, n_parts, sm_count) + 1;
}
bool join_emit_fits(const JoinEmitLaunch& L);
void launch_join_emit(const JoinEmitLaunch& L, uint64_t n_probe, int sm_count, cudaStream_t s);

// ---- k_gather_encode.cu ---------------------------------------------------------------------------
void launch_gather(const void* src, const uint32_t* src_valid, const uint32_t* idx, uint64_t n,
                   int elem_bytes, void* out, uint32_t* out_valid, int sm_count, cudaStream_t s,
                   uint32_t idx_mask = 0xffffffffu);
// values are read through idx, validity bits through vidx (both NULL = identity; vidx == idx is the
// plain row-id case, vidx != idx when the values were carried into a partitioned order);
// valid_bytes (one byte per value, read through idx) replaces the bitmap when it was carried too
// idx_mask strips flag bits from idx; valid_bit >= 0: the row is non-NULL iff that bit of idx is set
void launch_encode_fixed(const void* values, const uint32_t* valid, const uint8_t* valid_bytes, const uint32_t* idx,
                         const uint32_t* vidx, uint64_t n, int type, void* pages_out, int sm_count, cudaStream_t s,
                         uint32_t idx_mask = 0xffffffffu, int valid_bit = -1);
void launch_fill_u32(uint32_t* p, uint32_t v, uint64_t n, cudaStream_t s);
// *out += number of NULL cells among rows idx[0..n) of a column (idx NULL = identity; valid NULL = no NULLs)
void launch_count_nulls(const uint32_t* valid, const uint32_t* idx, uint64_t n, uint32_t idx_mask, unsigned long long* out,
                        int sm_count, cudaStream_t s);
void launch_bitmap_to_bytes(const uint32_t* bits, uint64_t n, uint8_t* out, int sm_count, cudaStream_t s);
void launch_bytes_to_bitmap(const uint8_t* bytes, uint64_t n, uint32_t* out, int sm_count, cudaStream_t s);

// ---- k_varchar.cu ---------------------------------------------------------------------------------
struct VarcharLayoutDev {
    uint64_t  n = 0;          // output rows
    const uint8_t*  src_pages = nullptr;
    const uint64_t* desc = nullptr;
    const uint32_t* valid = nullptr;
    const uint32_t* idx = nullptr;
    uint64_t* weight_scan = nullptr; // [n] inclusive prefix of row weights (bits)
    uint64_t* base_scan = nullptr;   // [n] weight prefix at the last segment breaker
    uint32_t* head_pages = nullptr;  // [n] pages started by row j (0 = not a page head)
    uint64_t* page_of = nullptr;     // [n+1] exclusive prefix of head_pages
    uint32_t* page_row = nullptr;    // [n_pages] first row of each page
    uint64_t* scalars = nullptr;     // [4] device: max normal weight, ...
    uint64_t  n_pages = 0;
};
// per-row 64-bit hash of a VARCHAR column (join keys); NULL rows hash to 0
void launch_varchar_hash(const uint8_t* pages, const uint64_t* desc, const uint32_t* valid, uint64_t n,
                         uint64_t* out_hash, int sm_count, cudaStream_t s);
// keep[i] = 1 iff the strings of pair i are byte-equal
void launch_varchar_pairs_equal(const uint8_t* pages_a, const uint64_t* desc_a, const uint32_t* idx_a,
                                const uint8_t* pages_b, const uint64_t* desc_b, const uint32_t* idx_b,
                                uint64_t n, uint32_t* keep, int sm_count, cudaStream_t s);
void launch_compact_pairs(const uint32_t* a, const uint32_t* b, const uint32_t* keep, const uint64_t* pos,
                          uint64_t n, uint32_t* out_a, uint32_t* out_b, cudaStream_t s);
void launch_varchar_weights(const VarcharLayoutDev& L, uint64_t* weights, int sm_count, cudaStream_t s);
void launch_varchar_marks(const VarcharLayoutDev& L, uint64_t* marks, int sm_count, cudaStream_t s);
void launch_varchar_heads(const VarcharLayoutDev& L, const uint64_t* weights, int sm_count, cudaStream_t s);
void launch_varchar_page_rows(const VarcharLayoutDev& L, int sm_count, cudaStream_t s);
void launch_varchar_write(const VarcharLayoutDev& L, uint8_t* pages_out, int sm_count, cudaStream_t s);

} // namespace rj
