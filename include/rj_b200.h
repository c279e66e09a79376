/*
 * rj_b200.h -- C-ABI of the B200-native radix hash-join engine.
 *
 * This is the drop-in boundary for the hot path of cliarie/radix-join:
 *   Contest::build_context / destroy_context / execute   (reference include/plan.h:337-344,
 *   implemented by the reference in src/execute.cpp:316-330).
 * Those three symbols are C++-mangled and traffic in C++ containers, so a foreign-function binding
 * cannot name them directly.  The entry points below carry the same information as plain pointers
 * and sizes; `radix-join_b200/csrc/contest_execute.cpp` is the 100-line adapter that turns a
 * `const Plan&` into an `rj_plan_t` and an `rj_result` back into a `ColumnarTable`.
 *
 * Conventions
 *   - every function that can fail returns int: 0 = ok, non-zero = error; the message is kept in the
 *     context (`rj_last_error`).  The C++ adapter rethrows it as std::runtime_error, which is the
 *     reference's error contract (src/execute.cpp:280, tests/read_sql.cpp:1329-1332).
 *   - no torch / C++ types in any signature; device pointers are `void*` / `uint64_t` addresses in the
 *     CUDA primary context of the device the context was created on; streams are `void*`
 *     (a `cudaStream_t`; NULL = the context's own non-blocking stream, which a stage call first orders behind the
 *     work already queued on the legacy default stream -- where a caller without streams prepared its buffers).
 *   - there is NO CPU fallback: if no sm_100 device is present every entry point fails.
 */
#ifndef RJ_B200_H
#define RJ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RJ_PAGE_SIZE 8192u /* reference include/plan.h:54 */

/* reference include/attribute.h:8-13 (same numeric values as `enum class DataType`) */
enum rj_data_type { RJ_INT32 = 0, RJ_INT64 = 1, RJ_FP64 = 2, RJ_VARCHAR = 3 };

/* ------------------------------------------------------------------------------------------------
 * Flattened plan -- mirrors Plan / PlanNode / ScanNode / JoinNode / ColumnarTable / Column
 * (reference include/plan.h:32-52, 60-62, 102-116).
 * ---------------------------------------------------------------------------------------------- */

/* Column (plan.h:60-62): a typed list of 8 KB pages.  Give either `pages` (n_pages pointers to
 * individually allocated pages, as the reference's `std::vector<Page*>`) or `contiguous`
 * (n_pages * 8192 bytes back to back); `pages` wins if both are set. */
typedef struct rj_column_t {
    int32_t            type;       /* rj_data_type */
    uint32_t           reserved;
    uint64_t           n_pages;
    const void* const* pages;
    const void*        contiguous;
} rj_column_t;

/* ColumnarTable (plan.h:102-105) */
typedef struct rj_table_t {
    uint64_t           num_rows;
    uint32_t           n_columns;
    uint32_t           reserved;
    const rj_column_t* columns;
} rj_table_t;

/* one element of PlanNode::output_attrs (plan.h:46): (source column index, declared type) */
typedef struct rj_attr_t {
    uint64_t index;
    int32_t  type;
    uint32_t reserved;
} rj_attr_t;

/* PlanNode (plan.h:44-52) with its ScanNode (:32-34) or JoinNode (:36-42) payload inlined */
typedef struct rj_node_t {
    int32_t          is_join;       /* 0 = ScanNode, 1 = JoinNode */
    int32_t          build_left;    /* JoinNode::build_left */
    uint64_t         base_table_id; /* ScanNode::base_table_id (index into inputs) */
    uint64_t         left, right;   /* JoinNode child node indices */
    uint64_t         left_attr, right_attr; /* index into the child's output_attrs */
    uint32_t         n_output_attrs;
    uint32_t         reserved;
    const rj_attr_t* output_attrs;
} rj_node_t;

/* Plan (plan.h:112-116) */
typedef struct rj_plan_t {
    uint32_t          n_nodes;
    uint32_t          n_inputs;
    const rj_node_t*  nodes;
    const rj_table_t* inputs;
    uint64_t          root;
} rj_plan_t;

typedef struct rj_ctx    rj_ctx;    /* the opaque `void* context` of Contest::build_context      */
typedef struct rj_inputs rj_inputs; /* base tables resident in HBM (uploaded pages)               */
typedef struct rj_result rj_result; /* result ColumnarTable: pages resident in HBM until fetched  */

/* ------------------------------------------------------------------------------------------------
 * Context  -- replaces Contest::build_context / destroy_context (src/execute.cpp:326-330)
 * ---------------------------------------------------------------------------------------------- */
int         rj_ctx_create(int device, rj_ctx** out);
/* A device GROUP driven by one process (n_devices a power of two, all with peer access to each other):
 * the context is a normal context on devices[0]; rj_execute_pages -- what Contest::execute calls -- runs an
 * eligible plan (one key / foreign-key join of two scans on an INT32 key whose outputs are the key and at
 * most two fixed-width columns per side, build side > 512 Ki rows) on ALL devices: each takes 1/n of both
 * tables' rows, scatter pass 1 stays local, the owner of a hash range runs pass 2 out of the other devices'
 * memory over NVLink, and the devices' result pages are appended in device order.  Everything else runs on
 * devices[0].  Replaces nothing in the reference (its execute() is a one-socket CPU program); SURVEY 8e. */
int         rj_ctx_create_multi(const int* devices, uint32_t n_devices, rj_ctx** out);
int         rj_ctx_group_size(const rj_ctx* ctx);
void        rj_ctx_destroy(rj_ctx* ctx);
const char* rj_last_error(const rj_ctx* ctx); /* ctx may be NULL: error of the last failed rj_ctx_create */
int         rj_ctx_device(const rj_ctx* ctx);
int         rj_ctx_sm_count(const rj_ctx* ctx);
/* the cudaStream_t every whole-path call of this context runs on (for timing with CUDA events) */
void*       rj_ctx_stream(const rj_ctx* ctx);
/* number of kernels this library has launched in the process so far (all contexts) */
uint64_t    rj_kernel_launch_count(void);
/* host threads used to gather/scatter individually allocated pages through pinned staging */
int         rj_ctx_set_host_threads(rj_ctx* ctx, int n);

/* ------------------------------------------------------------------------------------------------
 * Whole path -- replaces Contest::execute (src/execute.cpp:316-324)
 * ---------------------------------------------------------------------------------------------- */

/* Host pages in -> result.  H2D upload, decode, joins, gather+encode all happen inside. */
int rj_execute(rj_ctx* ctx, const rj_plan_t* plan, rj_result** out);

/* Host pages in -> host pages out, streamed.  Replaces the same call (src/execute.cpp:316-324) for
 * inputs whose transfer dominates: the plan runs once per row window of its largest table (an inner
 * join distributes over a union of its inputs; the table must be read by a single ScanNode and its
 * referenced columns be fixed-width, otherwise the call degrades to upload + execute + download) while
 * the next window is uploaded and the previous window's result pages are downloaded.  `chunk_bytes` =
 * page bytes per window (0 = 256 MiB).  Every time result pages of a column are ready, `sink` is asked
 * for a host buffer of n_pages * 8192 contiguous bytes; the pages are complete when the call returns.
 * Pages of different windows are independent (a window's last page may be partly filled), the row
 * order is the engine's usual free order.  Copies overlap only if the host buffers are pinned and the
 * input columns contiguous. */
typedef void* (*rj_page_sink_t)(void* user, uint32_t column, int32_t type, uint64_t n_pages);
int rj_execute_streamed(rj_ctx* ctx, const rj_plan_t* plan, uint64_t chunk_bytes, rj_page_sink_t sink,
                        void* user, uint64_t* num_rows);

/* Host pages in -> host pages out where EVERY page is an individually allocated object -- the shape of
 * the reference's ColumnarTable (std::vector<Page*> per column, include/plan.h:60-68), whose result
 * pages must come from `new Page` because Column::~Column deletes them (plan.h:95-99).  This is what
 * Contest::execute (src/execute.cpp:316-324) calls.  Same windowed pipeline as rj_execute_streamed; in
 * addition worker threads gather the input pages into pinned staging buffers and scatter the result
 * pages out of them, so host copies, both DMA directions and the kernels overlap.
 *   new_pages   allocate n pages of 8192 bytes (8-byte aligned) into out[0..n); non-zero = failure.
 *               Called from the engine's WORKER threads, several at a time (the worker that fills a
 *               batch of pages allocates it), so it must be thread-safe -- `new Page` is
 *   append      hand over the next n filled pages of result column `column` (ownership passes to the
 *               caller); called on the calling thread, for every column in the same window order, so
 *               the columns stay row-aligned; never called for a column without pages
 *   free_pages  give back pages that were allocated but will not be appended (error paths); may be NULL */
typedef struct rj_page_alloc_t {
    void* user;
    int  (*new_pages)(void* user, uint64_t n, void** out);
    int  (*append)(void* user, uint32_t column, int32_t type, void* const* pages, uint64_t n);
    void (*free_pages)(void* user, uint64_t n, void* const* pages);
} rj_page_alloc_t;
int rj_execute_pages(rj_ctx* ctx, const rj_plan_t* plan, uint64_t chunk_bytes, const rj_page_alloc_t* alloc,
                     uint64_t* num_rows);

/* Same, split so a benchmark can keep the inputs resident in HBM:
 * rj_inputs_upload copies every column of every table once (plan->inputs of the later call is
 * ignored); rj_execute_resident runs decode -> joins -> encode on device only. */
int  rj_inputs_upload(rj_ctx* ctx, const rj_table_t* tables, uint32_t n_tables, rj_inputs** out);
/* adopt pages that already are in device memory (e.g. produced by rj_gen_* or an exchange):
 * `contiguous` of every column is then a DEVICE address; nothing is copied or owned. */
int  rj_inputs_adopt_device(rj_ctx* ctx, const rj_table_t* tables, uint32_t n_tables, rj_inputs** out);
/* adopt columns that are already DECODED in device memory (dense values + optional validity bitmap),
 * e.g. what a rank owns after the multi-GPU exchange: the page decode is skipped for them.
 * Fixed-width types only; nothing is copied or owned. */
typedef struct rj_dense_column_t {
    int32_t         type;     /* RJ_INT32 / RJ_INT64 / RJ_FP64 */
    uint32_t        reserved;
    const void*     d_values; /* num_rows elements of 4 or 8 bytes */
    const uint32_t* d_valid;  /* bitmap words, bit i of word i/32; NULL = no NULLs */
} rj_dense_column_t;
typedef struct rj_dense_table_t {
    uint64_t                 num_rows;
    uint32_t                 n_columns;
    uint32_t                 reserved;
    const rj_dense_column_t* columns;
} rj_dense_table_t;
int  rj_inputs_adopt_dense(rj_ctx* ctx, const rj_dense_table_t* tables, uint32_t n_tables, rj_inputs** out);
void rj_inputs_free(rj_ctx* ctx, rj_inputs* in);
int  rj_execute_resident(rj_ctx* ctx, const rj_plan_t* plan, const rj_inputs* in, rj_result** out);

/* Result inspection -- the fields of the returned ColumnarTable (plan.h:102-105) */
uint64_t rj_result_num_rows(const rj_result* r);
uint32_t rj_result_num_columns(const rj_result* r);
int32_t  rj_result_column_type(const rj_result* r, uint32_t col);
uint64_t rj_result_column_pages(const rj_result* r, uint32_t col);
/* device address of the column's pages (n_pages * 8192 bytes, contiguous) */
uint64_t rj_result_column_device_ptr(const rj_result* r, uint32_t col);
/* D2H: copy the pages of one column into caller-owned host pages.  `dst_pages` = n_pages pointers
 * (the adapter passes freshly `new Page`d pages, plan.h:64-68,95-99) or `dst_contiguous`. */
int  rj_result_fetch(rj_ctx* ctx, const rj_result* r, uint32_t col, void* const* dst_pages,
                     void* dst_contiguous);
void rj_result_free(rj_ctx* ctx, rj_result* r);

/* ------------------------------------------------------------------------------------------------
 * Per-stage entry points (device pointers in, device pointers out).  The whole-path functions above
 * are composed of exactly these; tests, the bench and the multi-GPU driver call them directly.
 * All of them are asynchronous on `stream` unless they return a count.
 * ---------------------------------------------------------------------------------------------- */

/* -- page ingest: replaces Table::from_columnar (src/build_table.cpp:312-436) ------------------- */

/* Per-page row counts and their exclusive prefix: page_row_start[n_pages+1] (uint64, device).
 * A VARCHAR 0xffff page counts one row, a 0xfffe page none (build_table.cpp:384-405).
 * totals[0] = rows, totals[1] = non-null values (device uint64[2], may be NULL). */
int rj_page_row_offsets(rj_ctx* ctx, const void* d_pages, uint64_t n_pages, int32_t type,
                        uint64_t* d_page_row_start, uint64_t* d_totals, void* stream);

/* INT32 / INT64 / FP64 pages -> dense values[rows] (4 or 8 bytes each, NULL rows = 0) and a
 * validity bitmap (uint32 words, bit i of word i/32, LSB first; must be zeroed by the caller;
 * pass NULL when the column is known to hold no NULLs). */
int rj_decode_fixed(rj_ctx* ctx, const void* d_pages, uint64_t n_pages, int32_t type,
                    const uint64_t* d_page_row_start, void* d_values, uint32_t* d_valid,
                    void* stream);

/* VARCHAR pages -> string descriptors desc[rows] (uint64: bits 0-39 byte address of the first char
 * relative to d_pages, bits 40-62 length, bit 63 = long string spread over a 0xffff/0xfffe page
 * chain) and the validity bitmap.  Characters stay in the page buffer (late materialisation). */
int rj_decode_varchar(rj_ctx* ctx, const void* d_pages, uint64_t n_pages,
                      const uint64_t* d_page_row_start, uint64_t* d_desc, uint32_t* d_valid,
                      void* stream);

/* -- radix partitioning: replaces the histogram/prefix/scatter of src/execute.cpp:124-184 -------- */

/* Histogram of hash(key) over `bits` radix bits starting at bit `shift` of the 32-bit hash.
 * keys: uint32 (key_bytes=4) or uint64 (key_bytes=8); valid may be NULL; NULL keys are not counted
 * (execute.cpp:61-83: NULL keys never match).  d_hist: uint32[1<<bits], zeroed by the caller. */
int rj_radix_histogram(rj_ctx* ctx, const void* d_keys, const uint32_t* d_valid, uint64_t n,
                       int32_t key_bytes, int32_t shift, int32_t bits, uint32_t* d_hist,
                       void* stream);

/* Scatter (key, row index) into partition order.  d_idx_in may be NULL (= identity row ids).
 * d_cursor: uint32[1<<bits] holding each partition's start offset (exclusive prefix of the
 * histogram); it is advanced by the kernel.  Output order inside a partition is unspecified.
 * d_idx_out may be NULL when the row ids are not needed (only the keys are scattered). */
int rj_radix_scatter(rj_ctx* ctx, const void* d_keys, const uint32_t* d_valid,
                     const uint32_t* d_idx_in, uint64_t n, int32_t key_bytes, int32_t shift,
                     int32_t bits, uint32_t* d_cursor, void* d_keys_out, uint32_t* d_idx_out,
                     void* stream);

/* Scatter with one output base PER PARTITION (<= 8 partitions) and payload columns that travel with
 * the tuples.  The bases may be peer-mapped memory of other GPUs: with the partitions being the owner
 * ranks of a multi-GPU join this one kernel is partition + exchange (its write-combined runs become
 * coalesced stores over NVLink; no separate collective).  d_cursor[p] = first free slot in partition p's
 * buffer (advanced by the kernel).  rows_out may be all NULL.  pay_width: 4 / 8 = values, 1 = pay_src is a
 * validity BITMAP and one byte per tuple is written. */
typedef struct rj_scatter_multi_t {
    void*       keys_out[8];
    uint32_t*   rows_out[8];
    uint32_t    n_payload; /* <= 6 */
    uint32_t    reserved;
    const void* pay_src[6];
    int32_t     pay_width[6];
    void*       pay_dst[6][8];
} rj_scatter_multi_t;
int rj_radix_scatter_multi(rj_ctx* ctx, const void* d_keys, const uint32_t* d_valid, uint64_t n, int32_t key_bytes,
                           int32_t shift, int32_t bits, uint32_t* d_cursor, const rj_scatter_multi_t* out,
                           void* stream);

/* -- join: replaces hash_join_omp steps 2-6 (src/execute.cpp:61-261) ---------------------------- */

/* Inner equi-join of two key columns.  Emits (build row, probe row) pairs in unspecified order.
 * Returns the number of matches in *n_matches (synchronises).  If it exceeds `capacity` the output
 * arrays are incomplete and the call must be repeated with larger arrays (the count is exact). */
int rj_join_keys(rj_ctx* ctx, const void* d_build_keys, const uint32_t* d_build_valid,
                 uint64_t n_build, const void* d_probe_keys, const uint32_t* d_probe_valid,
                 uint64_t n_probe, int32_t key_bytes, uint64_t capacity, uint32_t* d_out_build,
                 uint32_t* d_out_probe, uint64_t* n_matches, void* stream);

/* -- late materialisation ---------------------------------------------------------------------- */

/* out[i] = src[idx[i]] for 4- or 8-byte elements; validity bits gathered alongside when
 * d_src_valid != NULL (d_out_valid: (n+31)/32 words, fully written). */
int rj_gather(rj_ctx* ctx, const void* d_src, const uint32_t* d_src_valid, const uint32_t* d_idx,
              uint64_t n, int32_t elem_bytes, void* d_out, uint32_t* d_out_valid, void* stream);

/* validity bitmap (bit i of word i/32) <-> one byte per row: the byte form is what travels through
 * an all-to-all-v, whose segments are not word aligned.  d_bits of rj_bytes_to_bitmap: (n+31)/32 words. */
int rj_bitmap_to_bytes(rj_ctx* ctx, const uint32_t* d_bits, uint64_t n, uint8_t* d_bytes, void* stream);
int rj_bytes_to_bitmap(rj_ctx* ctx, const uint8_t* d_bytes, uint64_t n, uint32_t* d_bits, void* stream);

/* -- page output: replaces Table::to_columnar (src/build_table.cpp:456-681) --------------------- */

/* rows-per-page used by rj_encode_fixed for a type (1984 for INT32, 1007 for INT64/FP64) */
uint32_t rj_fixed_rows_per_page(int32_t type);

/* Gather `n` rows of a fixed-width column through d_idx (NULL = identity) and write them as pages:
 * d_pages_out must hold ceil(n / rows_per_page) * 8192 bytes. */
int rj_encode_fixed(rj_ctx* ctx, const void* d_values, const uint32_t* d_valid,
                    const uint32_t* d_idx, uint64_t n, int32_t type, void* d_pages_out,
                    void* stream);

/* VARCHAR output is a two-step call because the page count depends on the data:
 * plan computes the page layout and returns the page count; write fills the pages. */
typedef struct rj_varchar_layout rj_varchar_layout;
int  rj_encode_varchar_plan(rj_ctx* ctx, const void* d_src_pages, const uint64_t* d_desc,
                            const uint32_t* d_valid, const uint32_t* d_idx, uint64_t n,
                            rj_varchar_layout** layout, uint64_t* n_pages_out, void* stream);
int  rj_encode_varchar_write(rj_ctx* ctx, rj_varchar_layout* layout, void* d_pages_out,
                             void* stream);
void rj_encode_varchar_free(rj_ctx* ctx, rj_varchar_layout* layout);

/* -- synthetic inputs on device (harness side: the role of ColumnInserter, plan.h:151-228) ------- */

/* Fill INT32/INT64/FP64 pages from dense device arrays, 1984/1007 rows per page (no greedy
 * packing).  d_valid may be NULL.  Returns the page count via *n_pages_out when d_pages_out is
 * NULL (size query). */
int rj_gen_fixed_pages(rj_ctx* ctx, const void* d_values, const uint32_t* d_valid, uint64_t n,
                       int32_t type, void* d_pages_out, uint64_t* n_pages_out, void* stream);

/* -- whole-tuple scatter and the join on partitioned inputs: the two halves of the fused root join
 *    (k_scatter_carry.cu, k_join_emit.cu), exposed so that the multi-GPU driver can run the FIRST scatter
 *    pass as the exchange (every pass-1 region lives on the GPU that owns its hash range) and the rest
 *    locally.  Replaces src/execute.cpp:169-261 + src/build_table.cpp:456-594 like the fused root join. -- */

/* Radix scatter of (4-byte key, up to two 4/8-byte value columns, up to two validity flags per tuple).
 *   flat pass    d_region_start == NULL: rows in input order; d_valid = key validity bitmap (tuples with a
 *                NULL key are dropped); flag_src = validity BITMAPS by row; flag_dst = one BYTE per tuple
 *   region pass  d_region_start / d_tile_start [n_regions + 1] (device): the input is a flat pass's output,
 *                flag_src = validity BYTES by position; tiles never straddle a region; cursor[region << bits | digit]
 *   exchange     n_owners > 0 (flat pass only): digit d goes to owner d >> owner_shift, whose arrays are
 *                keys_dst_multi / val_dst_multi / flag_dst_multi [owner] (peer-mapped device memory);
 *                cursor[d] indexes the owner's arrays.
 * digit = (hash(key) >> shift) & (2^bits - 1), bits <= 8; d_cursor [2^bits] (per region in a region pass) holds
 * the first output index of every digit and is advanced.  Keys and value sources must be 16-byte aligned. */
typedef struct rj_carry_scatter_t {
    const void*     d_keys;
    const uint32_t* d_valid;
    uint64_t        n;
    const uint32_t* d_region_start;
    const uint32_t* d_tile_start;
    uint32_t        n_regions;
    int32_t         shift, bits;
    uint32_t*       d_cursor;
    void*           d_keys_out;
    uint32_t        n_val;
    const void*     val_src[2];
    void*           val_dst[2];
    int32_t         val_width[2];
    uint32_t        n_flag;
    const void*     flag_src[2];
    void*           flag_dst[2];
    uint32_t        n_owners;
    int32_t         owner_shift;
    void*           keys_dst_multi[8];
    void*           val_dst_multi[2][8];
    void*           flag_dst_multi[2][8];
    /* region pass over SCATTERED inputs (the multi-GPU pull): d_src_table[x][0..4] = byte addresses (possibly
     * peer-mapped) of the keys, value 0, value 1, flag 0, flag 1 of region x, biased so that the element
     * index is the region's virtual position d_region_start[x] + i; d_region_group[x] = the cursor group
     * (cursor[group << bits | digit]) region x feeds.  d_keys / val_src / flag_src are then unused. */
    const uint64_t* d_src_table;
    const uint32_t* d_region_group;
} rj_carry_scatter_t;
int rj_scatter_carry(rj_ctx* ctx, const rj_carry_scatter_t* desc, void* stream);

/* One side of a join whose tuples are already grouped by the first `pass1_bits` of their partition number
 * (or by all of it): INT32 keys, value columns and one validity byte per tuple beside them. */
typedef struct rj_part_side_t {
    const void*    d_keys;
    uint64_t       n;
    uint32_t       n_cols;             /* <= 2 */
    const void*    d_vals[2];
    int32_t        types[2];           /* RJ_INT32 / RJ_INT64 / RJ_FP64 */
    const uint8_t* d_valid_bytes[2];   /* NULL: the column holds no NULL (pull: any non-NULL value = has NULLs) */
    /* pull (local_pass1_bits > 0 only): the pass-1 regions this rank owns are scattered over the ranks' pass-1
     * arrays; n_sub sub-regions (region-major, sender-minor) described like rj_carry_scatter_t's scattered
     * inputs.  n_sub == 0: d_keys / d_vals / d_valid_bytes are this rank's own arrays, regions in order. */
    uint32_t        n_sub;
    const uint64_t* d_src_table;       /* [n_sub][5] */
    const uint32_t* d_sub_start;       /* [n_sub + 1] virtual positions */
    const uint32_t* d_sub_tile;        /* [n_sub + 1] exclusive prefix of ceil(count / 4096) */
    const uint32_t* d_sub_group;       /* [n_sub] pass-1 region (local numbering) of every sub-region */
} rj_part_side_t;
typedef struct rj_part_out_t {
    int32_t side;                      /* 0 = build side, 1 = probe side */
    int32_t col;                       /* column of that side, -1 = the join key */
} rj_part_out_t;
/* Finish a partitioned key / foreign-key join: second scatter pass inside every pass-1 region (skipped when
 * local_pass1_bits == 0: the inputs are fully partitioned), then build + probe + page output in one kernel.
 * d_hist_build / d_hist_probe [2^local_bits]: tuples per local partition, in partition order = the order of the
 * inputs' groups; hash_bits = radix bits of the WHOLE job (the table slot is taken from the hash bits above).
 * *out = NULL (and 0 returned) when a table met a duplicate build key: run the general path instead. */
int rj_join_partitioned(rj_ctx* ctx, const rj_part_side_t* build, const rj_part_side_t* probe,
                        const uint32_t* d_hist_build, const uint32_t* d_hist_probe, int32_t local_bits,
                        int32_t local_pass1_bits, int32_t hash_bits, const rj_part_out_t* outs, uint32_t n_out,
                        rj_result** out);

/* Layout of the multi-GPU pull exchange, computed on the device from the all-gathered histograms:
 * d_hist [G][2][2^bits] (rank, side, final partition).  With ndig = 2^pass1_bits, for each side (0 build, 1 probe):
 *   d_cursor [2][ndig]        where this rank's run of every pass-1 digit starts in its OWN pass-1 arrays
 *   d_table  [2][ndig][5]     biased source addresses of the sub-regions this rank owns (rj_part_side_t.d_src_table)
 *   d_start / d_tile [2][ndig+1], d_group [2][ndig]   the sub-regions' tuple / tile prefixes and pass-1 regions
 *   d_local_hist [2][2^bits / G]   tuples per final partition of the owned range
 *   d_scalars [2][2]          tuples this rank owns / stores into other ranks' ranges
 * d_ptrs [2][5][8]: every rank's base address of (keys, value 0, value 1, flag 0, flag 1) per side (0 = absent),
 * d_widths [2][5] their element widths.  G = 2^g <= 8. */
int rj_dist_layout(rj_ctx* ctx, const uint32_t* d_hist, int32_t me, int32_t g, int32_t bits, int32_t pass1_bits,
                   const uint64_t* d_ptrs, const int32_t* d_widths, uint32_t* d_cursor, uint64_t* d_table, uint32_t* d_start,
                   uint32_t* d_tile, uint32_t* d_group, uint32_t* d_local_hist, uint64_t* d_scalars, void* stream);

/* -- pre-filter evaluation (harness side: Statement::eval on InnerColumns, src/statement.cpp:46-133,
 *    186-200; include/inner_column.h:170-325,386-562) and filter + emit (src/build_table.cpp:94-119,
 *    247-303) ------------------------------------------------------------------------------------- */

/* operators of a comparison: the values of the reference's Comparison::Op (include/statement.h:54-65) */
enum rj_cmp_op {
    RJ_OP_EQ = 0, RJ_OP_NEQ, RJ_OP_LT, RJ_OP_GT, RJ_OP_LEQ, RJ_OP_GEQ, RJ_OP_LIKE, RJ_OP_NOT_LIKE,
    RJ_OP_IS_NULL, RJ_OP_IS_NOT_NULL
};
/* logical operators: the values of LogicalOperation::Type (include/statement.h:186-190) */
enum rj_logic_op { RJ_LOGIC_AND = 0, RJ_LOGIC_OR, RJ_LOGIC_NOT };

/* Result of every predicate: one bit per row in uint32 words, LSB first (= the reference's
 * std::vector<uint8_t> on a little-endian host); bit = "not NULL and the comparison holds".  Bits past
 * row n are zero.  d_out holds (n + 31) / 32 words.
 * Fixed-width columns (decoded values + validity bitmap, rj_decode_fixed): ops EQ..GEQ; the literal is
 * rhs_i for INT32 (narrowed with static_cast<int32_t>, statement.cpp:55) and INT64, rhs_d for FP64. */
int rj_filter_compare(rj_ctx* ctx, const void* d_values, const uint32_t* d_valid, uint64_t n, int32_t type,
                      int32_t op, int64_t rhs_i, double rhs_d, uint32_t* d_out, void* stream);
/* VARCHAR columns (descriptors over the page buffer, rj_decode_varchar): ops EQ..GEQ compare bytes like
 * std::string_view, LIKE / NOT_LIKE follow statement.h:118-161 ('%' = ".*", '_' = ".", full match, '.'
 * never matches a newline).  `rhs` is a HOST pointer to rhs_len bytes. */
int rj_filter_varchar(rj_ctx* ctx, const void* d_pages, const uint64_t* d_desc, const uint32_t* d_valid,
                      uint64_t n, int32_t op, const char* rhs, uint64_t rhs_len, uint32_t* d_out, void* stream);
/* IS NULL (is_null != 0) / IS NOT NULL: the complement of / the validity bitmap (NULL = no NULLs) */
int rj_filter_null(rj_ctx* ctx, const uint32_t* d_valid, uint64_t n, int32_t is_null, uint32_t* d_out, void* stream);
/* AND / OR / NOT of result bitmaps (bitmap_and / bitmap_or / bitmap_not, statement.cpp:8-44); NOT ignores
 * d_b and, like the reference, turns a NULL row's false into true. */
int rj_bitmap_logic(rj_ctx* ctx, const uint32_t* d_a, const uint32_t* d_b, uint64_t n, int32_t op,
                    uint32_t* d_out, void* stream);
/* Row ids of the set bits, ascending (what from_inner_to_column's loop visits, build_table.cpp:94-119):
 * d_row_ids must hold n entries; *count (host) receives how many were written (synchronises). */
int rj_bitmap_select(rj_ctx* ctx, const uint32_t* d_bits, uint64_t n, uint32_t* d_row_ids, uint64_t* count,
                     void* stream);

/* A filter as a postfix program over the columns of ONE table: comparisons push a bitmap, AND / OR pop
 * two and push one, NOT pops one and pushes one; exactly one bitmap must remain. */
typedef struct rj_pred_t {
    int32_t     kind;      /* 0 = comparison, 1 = logical operator                                  */
    int32_t     op;        /* rj_cmp_op or rj_logic_op                                               */
    uint32_t    column;    /* comparison: column of the table                                        */
    int32_t     lit_type;  /* literal: RJ_INT64 (rhs_i), RJ_FP64 (rhs_d), RJ_VARCHAR (rhs_s), or -1   */
    int64_t     rhs_i;
    double      rhs_d;
    const char* rhs_s;     /* host pointer */
    uint64_t    rhs_s_len;
} rj_pred_t;
/* Filter + emit (Table::from_csv's tail, build_table.cpp:247-303): the rows of `table` (host pages) that
 * pass the program, in row order, as result pages of ALL its columns; n_prog = 0 keeps every row. */
int rj_filter_table(rj_ctx* ctx, const rj_table_t* table, const rj_pred_t* prog, uint32_t n_prog, rj_result** out);

/* VARCHAR pages from dense device strings (the role of ColumnInserter<std::string>, plan.h:230-335):
 * row i = bytes [d_offsets[i], d_offsets[i+1]) of one character buffer (n + 1 offsets).  This call turns the
 * offsets into string descriptors over that buffer; rj_encode_varchar_plan(d_chars as the source "pages",
 * d_desc, d_valid, NULL, n) + rj_encode_varchar_write then produce the pages (strings longer than 8185 bytes
 * become 0xffff / 0xfffe chains).  Strings are limited to 2^23 - 1 bytes, the buffer to 2^40 bytes. */
int rj_varchar_descriptors(rj_ctx* ctx, const uint64_t* d_offsets, uint64_t n, uint64_t* d_desc, void* stream);

/* -- result validation on the device (harness side: the sorted-multiset comparison of
 *    tests/read_sql.cpp:1159-1222) ---------------------------------------------------------------- */

/* Are two ColumnarTables (host pages) the same MULTISET of rows?  Both are decoded on the device, every row
 * is hashed over all its cells, each table's (hash, row id) pairs are radix-sorted, and the rows at equal
 * rank are compared in full: validity, value bit patterns (so NaN == NaN and -0.0 != 0.0, like the
 * reference's variant compare of the bytes it wrote), string bytes.  *equal = 1 iff the column types, row
 * counts and all pairs agree; *mismatches (may be NULL) = number of cell / hash pairs that differed. */
int rj_tables_equal(rj_ctx* ctx, const rj_table_t* a, const rj_table_t* b, int32_t* equal, uint64_t* mismatches);

/* ------------------------------------------------------------------------------------------------
 * Profiling: per-kernel-class CUDA-event timings accumulated by the whole-path functions.
 * ---------------------------------------------------------------------------------------------- */
enum rj_stage {
    RJ_ST_H2D = 0,       /* host page gather + H2D copies                  */
    RJ_ST_ROW_OFFSETS,   /* page header scan                               */
    RJ_ST_DECODE,        /* page decode kernels                            */
    RJ_ST_HISTOGRAM,     /* radix histogram                                */
    RJ_ST_SCATTER,       /* radix scatter (all passes)                     */
    RJ_ST_JOIN,          /* shared-memory build + probe                    */
    RJ_ST_GATHER,        /* row-id / key gathers                           */
    RJ_ST_ENCODE,        /* gather + page encode                           */
    RJ_ST_D2H,           /* D2H copies + host page scatter                 */
    RJ_ST_JOIN_EMIT,     /* root join fused with page output               */
    RJ_ST_COUNT
};
typedef struct rj_stage_stat_t {
    double   ms;        /* summed device time                      */
    uint64_t launches;  /* kernel launches (or copies) in the sum  */
    uint64_t bytes;     /* algorithmic bytes (SURVEY 8d accounting)*/
} rj_stage_stat_t;
int         rj_profile_enable(rj_ctx* ctx, int on);
int         rj_profile_reset(rj_ctx* ctx);
int         rj_profile_read(rj_ctx* ctx, rj_stage_stat_t* stats /* [RJ_ST_COUNT] */);
const char* rj_stage_name(int stage);

const char* rj_version(void);

#ifdef __cplusplus
}
#endif
#endif /* RJ_B200_H */
