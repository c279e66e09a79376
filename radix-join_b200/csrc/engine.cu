// Host side of the engine: context, stream-ordered memory, page upload/download through pinned
// staging, the plan-tree driver with late materialisation, and the C-ABI of include/rj_b200.h.
//
// Plan driver (replaces execute_impl / execute_scan / execute_hash_join / hash_join_omp of the
// reference, src/execute.cpp:43-314) -- the reference materialises every intermediate as
// vector<vector<variant>> rows; here an intermediate is a set of ROW-ID LISTS, one per base-table scan
// below the node, kept in HBM.  Join keys of an intermediate are gathered through its row ids, only the
// root's output_attrs are ever materialised, straight into result pages.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <deque>
#include <map>
#include <mutex>
#include <set>
#include <thread>

#include "host_pipe.h"
#include "rj_internal.h"

using namespace rj;

namespace rj {
std::atomic<uint64_t> g_kernel_launches{0};
}

// ================================================================================================
// context
// ================================================================================================
struct PendingTiming {
    int         stage;
    cudaEvent_t a, b;
};

namespace {
struct BlockCache;
}

struct rj_ctx {
    int          device    = 0;
    int          sm_count  = 0;
    cudaStream_t stream    = nullptr;
    std::string  err;
    int          host_threads = 8;
    bool         profiling = false;
    rj_stage_stat_t            stats[RJ_ST_COUNT] = {};
    std::vector<PendingTiming> pending;
    std::vector<cudaEvent_t>   free_events;
    // worker pool + pinned staging rings for H2D / D2H of individually allocated pages (host_pipe.h);
    // created on first use, kept for the life of the context
    std::unique_ptr<HostPipe> pipe;
    std::shared_ptr<BlockCache> cache;
    // RJ_ERR_* bits set by kernels that meet malformed input: mapped pinned memory, read after a sync
    uint32_t* err_host = nullptr;
    uint32_t* err_dev = nullptr;
    // rj_execute_streamed: uploads and downloads run beside the kernels on their own streams
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    cudaEvent_t  up_ev[2] = {nullptr, nullptr};
    // rj_ctx_create_multi: the contexts of the other devices of the group (this one is device 0 of it)
    std::vector<rj_ctx*> subs;
};

static thread_local std::string g_create_error;

namespace {

struct EngineError: std::runtime_error {
    using std::runtime_error::runtime_error;
};

// ---- device memory: a caching block allocator ------------------------------------------------------
// execute() allocates ~35 buffers (tens of GB at config 2).  cudaMallocAsync cost 4-290 ms per call
// sequence on B200 (pool growth / remapping), more than all kernels together, so freed blocks are kept
// in a per-device cache keyed by size and handed out again: after the first execute() of a given plan
// shape no driver allocation happens.  All engine work is ordered on the context stream, so a block
// may be reused as soon as it is released; releases from a foreign stream synchronise that stream.
// host-side cost of the allocator, printed per execute() when RJ_TRACE is set
uint64_t g_alloc_ns = 0, g_free_ns = 0, g_alloc_bytes = 0, g_alloc_calls = 0, g_alloc_misses = 0;

struct BlockCache {
    std::mutex                   mu;
    std::multimap<size_t, void*> free_blocks;
    size_t                       cached_bytes = 0;

    // Size classes: 512-byte multiples below 64 KB, above that four classes per octave (1, 1.25, 1.5, 1.75 x 2^k),
    // so that the intermediates of different joins and plans -- whose sizes are all over the place -- fall into
    // the same few classes and a freed block is found again.  (With exact 2 MB rounding and a 12.5 % fit window
    // the JOB suite missed the cache ~60 times per plan, and a cudaMalloc of a fresh block costs ~3 ms: 190 of
    // the 240 ms of plan 30a.)
    static size_t round(size_t n) {
        if (n < 512) return 512;
        if (n < (size_t(64) << 10)) return (n + 511) & ~size_t(511);
        int k = 63;
        while (!((n >> k) & 1)) --k;              // 2^k <= n
        const size_t q = size_t(1) << (k - 2);    // quarter of the octave
        return (n + q - 1) & ~(q - 1);
    }
    // a free block of the request's class, or of one of the next classes up to 1.5 x the request
    void* take(size_t n, size_t* got) {
        std::lock_guard<std::mutex> lk(mu);
        auto it = free_blocks.lower_bound(n);
        if (it != free_blocks.end() && it->first <= n + n / 2) {
            void* p = it->second;
            *got = it->first;
            cached_bytes -= it->first;
            free_blocks.erase(it);
            return p;
        }
        return nullptr;
    }
    void give(void* p, size_t n) {
        std::lock_guard<std::mutex> lk(mu);
        free_blocks.emplace(n, p);
        cached_bytes += n;
    }
    void trim() {
        std::lock_guard<std::mutex> lk(mu);
        for (auto& kv: free_blocks) cudaFree(kv.second);
        free_blocks.clear();
        cached_bytes = 0;
    }
    ~BlockCache() { trim(); }
};
// One cache per CONTEXT (blocks released on a context's stream are reused without synchronisation, which
// is only sound while every user of the cache orders its work on that stream).  Buffers keep their cache
// alive, so a result may outlive its context.

struct DevMem {
    void*        p = nullptr;
    size_t       bytes = 0;    // requested
    size_t       block = 0;    // size of the underlying block
    cudaStream_t stream = nullptr;
    cudaStream_t home = nullptr; // the context stream: releases from it need no synchronisation
    std::shared_ptr<BlockCache> cache;
    DevMem(size_t n, cudaStream_t s, cudaStream_t ctx_stream, std::shared_ptr<BlockCache> c): bytes(n), stream(s), home(ctx_stream), cache(std::move(c)) {
        auto t0 = std::chrono::steady_clock::now();
        const size_t want = BlockCache::round(n);
        p = cache->take(want, &block);
        if (!p) {
            ++g_alloc_misses;
            block = want;
            cudaError_t e = cudaMalloc(&p, block);
            if (e == cudaErrorMemoryAllocation) {
                cudaGetLastError();
                cudaDeviceSynchronize();
                cache->trim(); // give the cached blocks back and retry once
                e = cudaMalloc(&p, block);
            }
            if (e != cudaSuccess) {
                p = nullptr;
                throw CudaError(std::string("cudaMalloc(") + std::to_string(block) + "): " + cudaGetErrorString(e));
            }
        }
        g_alloc_ns += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
        g_alloc_bytes += n;
        ++g_alloc_calls;
    }
    ~DevMem() {
        if (!p || view) return;
        if (stream != home) cudaStreamSynchronize(stream);
        cache->give(p, block);
    }
    // non-owning view of caller memory (adopted dense columns)
    DevMem(void* ptr, size_t n): p(ptr), bytes(n), view(true) {}
    bool view = false;
    DevMem(const DevMem&) = delete;
    DevMem& operator=(const DevMem&) = delete;
    template <class T>
    T* as() const { return static_cast<T*>(p); }
};
using Buf = std::shared_ptr<DevMem>;

thread_local cudaStream_t t_home_stream = nullptr; // set by guarded() for the duration of an API call
thread_local std::shared_ptr<BlockCache> t_cache;

Buf dev_alloc(size_t bytes, cudaStream_t s) {
    if (!t_cache) throw CudaError("internal: device allocation outside an engine call");
    return std::make_shared<DevMem>(bytes, s, t_home_stream, t_cache);
}

Buf dev_alloc_zero(size_t bytes, cudaStream_t s) {
    Buf b = dev_alloc(bytes, s);
    RJ_CUDA(cudaMemsetAsync(b->p, 0, bytes ? bytes : 16, s));
    return b;
}

// ---- profiling ---------------------------------------------------------------------------------
struct StageScope {
    rj_ctx*      ctx;
    int          stage;
    cudaStream_t stream;
    cudaEvent_t  a = nullptr, b = nullptr;
    StageScope(rj_ctx* c, int st, cudaStream_t s, uint64_t launches, uint64_t bytes): ctx(c), stage(st), stream(s) {
        if (!ctx->profiling) return;
        ctx->stats[stage].launches += launches;
        ctx->stats[stage].bytes += bytes;
        a = take();
        b = take();
        cudaEventRecord(a, stream);
    }
    ~StageScope() {
        if (!a) return;
        cudaEventRecord(b, stream);
        ctx->pending.push_back({stage, a, b});
    }
    cudaEvent_t take() {
        if (!ctx->free_events.empty()) {
            cudaEvent_t e = ctx->free_events.back();
            ctx->free_events.pop_back();
            return e;
        }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
};

void profile_collect(rj_ctx* ctx) {
    for (auto& t: ctx->pending) {
        cudaEventSynchronize(t.b);
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, t.a, t.b) == cudaSuccess) ctx->stats[t.stage].ms += ms;
        ctx->free_events.push_back(t.a);
        ctx->free_events.push_back(t.b);
    }
    ctx->pending.clear();
}

// ---- host helpers ------------------------------------------------------------------------------
size_t type_width(int t) { return t == RJ_INT32 ? 4 : 8; }

const uint8_t* host_page(const rj_column_t& c, uint64_t i) {
    if (c.pages) return static_cast<const uint8_t*>(c.pages[i]);
    return static_cast<const uint8_t*>(c.contiguous) + i * RJ_PAGE_SIZE;
}

// rows / non-NULL values a page contributes (reference src/build_table.cpp:326,383-407)
inline void page_counts(const uint8_t* pg, int type, uint64_t* rows, uint64_t* vals) {
    uint16_t n_r, n_v;
    std::memcpy(&n_r, pg, 2);
    std::memcpy(&n_v, pg + 2, 2);
    if (type == RJ_VARCHAR && n_r >= 0xfffe) {
        *rows += n_r == 0xffff ? 1 : 0;
        *vals += n_r == 0xffff ? 1 : 0;
    } else {
        *rows += n_r;
        *vals += n_v;
    }
}

} // namespace

// ================================================================================================
// resident inputs / results
// ================================================================================================
struct ColumnDev {
    int            type = 0;
    uint64_t       n_pages = 0;
    uint64_t       page_rows = 0; // rows held by the pages
    uint64_t       non_null = 0;
    const uint8_t* pages = nullptr;
    Buf            owned;
    // already decoded by the caller (rj_inputs_adopt_dense)
    bool            dense = false;
    const void*     dense_values = nullptr;
    const uint32_t* dense_valid = nullptr;
    // row window of a streamed table (rj_execute_streamed): the pages hold `skip_rows` rows before the
    // table's first row and possibly rows behind its last one
    bool            windowed = false;
    uint64_t        skip_rows = 0;
    bool            window_nulls = false; // the column holds NULLs somewhere (keep a validity bitmap)
};

struct TableDev {
    uint64_t               num_rows = 0;
    std::vector<ColumnDev> cols;
};

struct rj_inputs {
    std::vector<TableDev> tables;
};

struct ResultColumn {
    int      type = 0;
    uint64_t n_pages = 0;
    Buf      pages;
};

struct rj_result {
    uint64_t                  num_rows = 0;
    std::vector<ResultColumn> cols;
};

namespace {

HostPipe* ensure_pipe(rj_ctx* ctx) {
    if (!ctx->pipe) {
        // RJ_PIPE_BUFS caps each staging ring (tests: a ring of 2 buffers wraps after 1024 pages)
        const char* env = getenv("RJ_PIPE_BUFS");
        const int   cap = env && atoi(env) > 0 ? atoi(env) : 0;
        ctx->pipe = std::make_unique<HostPipe>(ctx->host_threads, ctx->device, cap);
    }
    return ctx->pipe.get();
}

void ensure_copy_streams(rj_ctx* ctx) {
    if (ctx->h2d_stream) return;
    RJ_CUDA(cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking));
    RJ_CUDA(cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) RJ_CUDA(cudaEventCreateWithFlags(&ctx->up_ev[i], cudaEventDisableTiming));
}

// fn(begin, end) over [0, n) on the context's worker pool; returns when every piece has run
template <class F>
void pool_for(rj_ctx* ctx, uint64_t n, uint64_t grain, F fn) {
    if (n == 0) return;
    HostPipe* hp = ensure_pipe(ctx);
    const uint64_t pieces = std::min<uint64_t>((n + grain - 1) / grain, uint64_t(hp->pool.size()) * 4);
    if (pieces <= 1) {
        fn(uint64_t(0), n);
        return;
    }
    TaskGroup g;
    g.open();
    for (uint64_t i = 0; i < pieces; ++i) {
        const uint64_t b = n * i / pieces, e = n * (i + 1) / pieces;
        g.add();
        hp->pool.submit([&g, &fn, b, e] {
            try {
                fn(b, e);
                g.done();
            } catch (...) {
                g.fail(std::current_exception());
            }
        });
    }
    g.seal();
    g.wait();
}

struct ColCounters {
    std::atomic<uint64_t> rows{0}, vals{0};
};

// H2D of pages [p0, p0 + cnt) of one host column, asynchronously: the copies are ISSUED on `stream` by
// the time `group` completes.  Individually allocated pages (plan.h:60-68) are gathered into pinned
// ring buffers by pool workers, 512 pages per task; each task issues its own DMA, so the gather of
// one buffer overlaps the transfer of the previous ones.  A contiguous host buffer is one DMA straight
// from the caller's memory.  `counters` (optional) receives the row / non-NULL totals of the page headers.
// host_copy.cpp: one 8 KB page, host to host, with non-temporal stores
void upload_pages_async(rj_ctx* ctx, const rj_column_t& c, uint64_t p0, uint64_t cnt, uint8_t* dst, cudaStream_t stream,
                        TaskGroup* group, ColCounters* counters) {
    if (cnt == 0) return;
    if (!c.pages && !c.contiguous) throw EngineError("column has pages but no page pointers");
    HostPipe* hp = ensure_pipe(ctx);
    const int type = c.type;
    // a contiguous buffer in PAGEABLE host memory (a numpy array, a std::vector) would be staged by the driver on
    // one thread at 6-10 GB/s; above a few MB it takes the same road as individually allocated pages -- worker
    // threads copy it into the pinned ring, the DMA engine reads from there (~25 GB/s and overlapped)
    bool pageable_contiguous = false;
    if (!c.pages && cnt * size_t(RJ_PAGE_SIZE) >= (size_t(4) << 20)) {
        cudaPointerAttributes attr{};
        if (cudaPointerGetAttributes(&attr, c.contiguous) != cudaSuccess) cudaGetLastError();
        pageable_contiguous = attr.type == cudaMemoryTypeUnregistered;
    }
    if (!c.pages && !pageable_contiguous) {
        const uint8_t* src = static_cast<const uint8_t*>(c.contiguous) + p0 * RJ_PAGE_SIZE;
        RJ_CUDA(cudaMemcpyAsync(dst, src, cnt * size_t(RJ_PAGE_SIZE), cudaMemcpyHostToDevice, stream));
        if (counters) {
            const uint64_t grain = 8192;
            for (uint64_t q = 0; q < cnt; q += grain) {
                const uint64_t m = std::min<uint64_t>(grain, cnt - q);
                group->add();
                hp->pool.submit([=] {
                    uint64_t r = 0, v = 0;
                    for (uint64_t i = 0; i < m; ++i) page_counts(src + (q + i) * RJ_PAGE_SIZE, type, &r, &v);
                    counters->rows.fetch_add(r, std::memory_order_relaxed);
                    counters->vals.fetch_add(v, std::memory_order_relaxed);
                    group->done();
                });
            }
        }
        return;
    }
    const void* const* pages = c.pages ? c.pages + p0 : nullptr;
    const uint8_t*     flat = c.pages ? nullptr : static_cast<const uint8_t*>(c.contiguous) + p0 * RJ_PAGE_SIZE;
    const uint64_t per = HostPipe::kBufBytes / RJ_PAGE_SIZE;
    for (uint64_t q = 0; q < cnt; q += per) {
        const uint64_t m = std::min<uint64_t>(per, cnt - q);
        group->add();
        hp->pool.submit([=] {
            try {
                PinnedBuf* buf = hp->up.acquire();
                uint64_t r = 0, v = 0;
                for (uint64_t i = 0; i < m; ++i) {
                    const uint8_t* pg = pages ? static_cast<const uint8_t*>(pages[q + i]) : flat + (q + i) * RJ_PAGE_SIZE;
                    copy_page(buf->p + i * RJ_PAGE_SIZE, pg);
                    if (counters) page_counts(pg, type, &r, &v);
                }
                if (counters) {
                    counters->rows.fetch_add(r, std::memory_order_relaxed);
                    counters->vals.fetch_add(v, std::memory_order_relaxed);
                }
                copy_fence();
                cudaError_t e = cudaMemcpyAsync(dst + q * RJ_PAGE_SIZE, buf->p, m * RJ_PAGE_SIZE, cudaMemcpyHostToDevice, stream);
                if (e == cudaSuccess) e = cudaEventRecord(buf->ev, stream);
                hp->up.release(buf);
                if (e != cudaSuccess) throw CudaError(std::string("page upload: ") + cudaGetErrorString(e));
                group->done();
            } catch (...) {
                group->fail(std::current_exception());
            }
        });
    }
}

void check_column_rows(const TableDev& t, const ColumnDev& c) {
    // the reference throws "row_idx" when a page holds a value beyond num_rows (build_table.cpp:334-336)
    if (c.page_rows > t.num_rows) throw EngineError("row_idx");
}

// Upload of whole columns: every wanted column's copies are in flight together; returns once they are
// issued on `stream` (the kernels that follow on the context stream are ordered behind them).
std::unique_ptr<rj_inputs> upload_tables(rj_ctx* ctx, const rj_table_t* tables, uint32_t n, const std::set<std::pair<uint32_t, uint32_t>>* wanted) {
    auto in = std::make_unique<rj_inputs>();
    in->tables.resize(n);
    uint64_t bytes = 0, copies = 0;
    for (uint32_t t = 0; t < n; ++t)
        for (uint32_t c = 0; c < tables[t].n_columns; ++c)
            if (!wanted || wanted->count({t, c})) {
                bytes += tables[t].columns[c].n_pages * RJ_PAGE_SIZE;
                ++copies;
            }
    StageScope scope(ctx, RJ_ST_H2D, ctx->stream, copies, bytes);
    TaskGroup group;
    std::deque<ColCounters> counters; // stable addresses
    std::vector<std::pair<TableDev*, ColumnDev*>> cols;
    group.open();
    try {
        for (uint32_t t = 0; t < n; ++t) {
            TableDev& td = in->tables[t];
            td.num_rows = tables[t].num_rows;
            td.cols.resize(tables[t].n_columns);
            for (uint32_t c = 0; c < tables[t].n_columns; ++c) {
                const rj_column_t& hc = tables[t].columns[c];
                ColumnDev& cd = td.cols[c];
                cd.type = hc.type;
                if (wanted && !wanted->count({t, c})) continue; // never referenced by the plan
                cd.n_pages = hc.n_pages;
                cd.owned = dev_alloc(hc.n_pages * size_t(RJ_PAGE_SIZE), ctx->stream);
                cd.pages = cd.owned->as<uint8_t>();
                counters.emplace_back();
                cols.emplace_back(&td, &cd);
                upload_pages_async(ctx, hc, 0, hc.n_pages, cd.owned->as<uint8_t>(), ctx->stream, &group, &counters.back());
            }
        }
    } catch (...) {
        group.seal();
        try { group.wait(); } catch (...) {}
        throw;
    }
    group.seal();
    group.wait();
    for (size_t i = 0; i < cols.size(); ++i) {
        cols[i].second->page_rows = counters[i].rows.load();
        cols[i].second->non_null = counters[i].vals.load();
        check_column_rows(*cols[i].first, *cols[i].second);
    }
    return in;
}

// ================================================================================================
// execution state
// ================================================================================================
struct DecodedCol {
    int             type = 0;
    uint64_t        rows = 0;
    Buf             values; // INT32: u32[rows]; INT64/FP64: u64[rows]; VARCHAR: desc u64[rows]
    Buf             valid;  // null when the column holds no NULL
    const uint8_t*  pages = nullptr;
    Buf             str_hash; // VARCHAR join keys: 64-bit hash per row (lazily)
    Buf             hold_values, hold_valid; // owners when values / valid are views into a larger decode
    const uint32_t* valid_ptr() const { return valid ? valid->as<uint32_t>() : nullptr; }
};

struct Attr {
    int      leaf;  // scan node index
    uint32_t table;
    uint32_t col;
};

// Row ids of a scan reached through a partitioned join side: rid[i] = rows_of_pos[pos[i]], where
// `pos` indexes the side's position order (the pass-1 / single-pass scatter output).  Kept lazy: the
// root encodes carried columns straight from position order and may never need the row ids.
struct PosSpace {
    Buf      pos;         // [rows]; the top bits may carry validity flags
    Buf      rows_of_pos; // [side rows]
    uint32_t mask = 0xffffffffu;
};

// A column that travelled through the scatter: values_of_pos[pos[i]] is row i's value.
struct CarriedCol {
    Buf         hold;              // owner of values_of_pos when it is a scatter output
    const void* values_of_pos = nullptr;
    Buf         valid_hold;        // one validity byte per position (carried with the values), or null
    Buf         pos;
    uint32_t    pos_mask = 0xffffffffu;
    int         valid_bit = -1;     // >= 0: validity of this column rides in that bit of pos
    bool        never_null = false; // a matched join key is never NULL
};

struct Rel {
    uint64_t                rows = 0;
    bool                    streamed = false; // holds the table that rj_execute_streamed feeds in windows
    mutable std::map<int, Buf> rid;  // leaf -> row ids into that scan's base table; null Buf = identity
    std::map<int, PosSpace>    lazy; // leaf -> row ids not materialised yet
    std::map<std::pair<int, uint32_t>, CarriedCol> carried; // (leaf, column) -> values in position order
    bool has_leaf(int leaf) const { return rid.count(leaf) || lazy.count(leaf); }
};

// one side of a join as the kernels see it
struct CarryCol {
    uint32_t    col = 0;
    const void* src = nullptr; // decoded values by row
    int         width = 0;
    Buf         out;           // values by position (filled by the scatter)
    const uint32_t* valid_src = nullptr; // validity bitmap by row (null: no NULLs)
    Buf         valid_out;     // validity bytes by position
    int         valid_bit = -1; // >= 0: also folded into that bit of the emitted positions
};

struct JoinSide {
    const void*     keys = nullptr;
    const uint32_t* valid = nullptr;
    uint64_t        n = 0;
    std::vector<CarryCol> carry; // columns to move with the tuples (only honoured when partitioned)
    bool need_rows = true;       // false: nothing downstream asks for row ids, the scatter skips them
    // results
    bool partitioned = false;
    uint32_t pos_mask = 0xffffffffu; // strips the validity flags pass 2 may fold into the positions
    Buf  pos;          // [M] position (= row when not partitioned)
    Buf  rows_of_pos;  // [n]
    Buf  keys_of_pos;  // [n]
};

struct Exec {
    rj_ctx*          ctx;
    const rj_plan_t* plan;
    const rj_inputs* in;
    cudaStream_t     s;
    using DecodedMap = std::map<std::pair<uint32_t, uint32_t>, DecodedCol>;
    DecodedMap  own_decoded;
    // rj_execute_streamed runs the plan once per chunk of one table: the decoded columns of all the
    // other tables are shared between those runs
    DecodedMap* shared_decoded = nullptr;
    uint32_t    streamed_table = 0xffffffffu;
    double      streamed_scale = 1.0; // whole table / this window: joins choose their sides as the whole job would
    DecodedMap& decoded_of(uint32_t t) { return shared_decoded && t != streamed_table ? *shared_decoded : own_decoded; }

    Exec(rj_ctx* c, const rj_plan_t* p, const rj_inputs* i): ctx(c), plan(p), in(i), s(c->stream) {}

    const rj_node_t& node(uint64_t i) const {
        if (i >= plan->n_nodes) throw EngineError("plan node index out of range");
        return plan->nodes[i];
    }

    // where does output attribute `a` of node `n` come from?
    Attr resolve(uint64_t n, uint64_t a) const {
        const rj_node_t& nd = node(n);
        if (a >= nd.n_output_attrs) throw EngineError("attribute index out of range");
        const uint64_t src = nd.output_attrs[a].index;
        if (!nd.is_join) {
            if (nd.base_table_id >= in->tables.size()) throw EngineError("base table out of range");
            if (src >= in->tables[nd.base_table_id].cols.size()) throw EngineError("scan attribute out of range");
            return {static_cast<int>(n), static_cast<uint32_t>(nd.base_table_id), static_cast<uint32_t>(src)};
        }
        const uint64_t left_w = node(nd.left).n_output_attrs; // src/execute.cpp:57,238-241
        return src < left_w ? resolve(nd.left, src) : resolve(nd.right, src - left_w);
    }

    const DecodedCol& column(uint32_t t, uint32_t c);
    const DecodedCol& string_hash(uint32_t t, uint32_t c);
    Rel  run(uint64_t n);
    Rel  join(uint64_t n, const Rel& L, const Rel& R);
    void join_keys(JoinSide& b, JoinSide& p, int key_bytes, uint64_t* n_out);
    Buf  gather_u32(const Buf& src, const Buf& idx, uint64_t n, uint32_t idx_mask = 0xffffffffu);
    Buf  side_rows(const JoinSide& sd, uint64_t m);
    Buf  rid_of(const Rel& r, int leaf);
    std::unique_ptr<rj_result> root(uint64_t n, const Rel& r);
    std::unique_ptr<rj_result> root_fused(uint64_t n);
    ResultColumn encode_varchar(const DecodedCol& col, const uint32_t* idx, uint64_t n);
};

// ---- page ingest -----------------------------------------------------------------------------------
// `bias` (< 256) rows of padding precede the first decoded row: a row window that starts in the middle
// of a page is shifted so that its first row lands on a validity-word boundary.
void decode_pages(rj_ctx* ctx, cudaStream_t s, const uint8_t* pages, uint64_t n_pages, int type, uint64_t rows,
                  bool need_valid, DecodedCol* out, uint32_t bias = 0, uint64_t covered_rows = 0) {
    out->type = type;
    out->rows = rows;
    out->pages = pages;
    const size_t w = type == RJ_INT32 ? 4 : 8;
    out->values = dev_alloc(rows * w, s);
    if (need_valid) out->valid = dev_alloc_zero(((rows + 31) / 32 + 1) * 4, s);
    if (n_pages == 0) return;
    // entry 0 of the count array is the padding; the scan then yields biased row starts
    Buf row_cnt = dev_alloc((n_pages + 1) * 4, s);
    Buf row_start_buf = dev_alloc((n_pages + 2) * 8, s);
    Buf tmp = dev_alloc(scan_tmp_bytes(n_pages + 1), s);
    const uint64_t* row_start = row_start_buf->as<uint64_t>() + 1;
    {
        StageScope sc(ctx, RJ_ST_ROW_OFFSETS, s, 4, n_pages * 4);
        RJ_CUDA(cudaMemsetAsync(row_cnt->p, 0, 4, s));
        if (bias) RJ_CUDA(cudaMemsetAsync(row_cnt->p, static_cast<int>(bias), 1, s)); // little-endian low byte
        launch_page_rows(pages, n_pages, type, row_cnt->as<uint32_t>() + 1, nullptr, s);
        launch_exclusive_scan_u32_u64(row_cnt->as<uint32_t>(), row_start_buf->as<uint64_t>(), n_pages + 1, tmp->p, s);
    }
    {
        // 8192 * pages read + rows * w + rows / 8 written
        StageScope sc(ctx, RJ_ST_DECODE, s, 1, n_pages * uint64_t(RJ_PAGE_SIZE) + rows * w + rows / 8);
        if (need_valid && rows * w) {
            // rows the pages do not cover stay NULL and read as 0 (the decode kernel itself stores 0 for the
            // NULL rows it covers): only the uncovered head / tail is cleared, not the whole array
            uint8_t* v = out->values->as<uint8_t>();
            const uint64_t lo = std::min<uint64_t>(bias, rows), hi = std::min<uint64_t>(rows, bias + covered_rows);
            if (lo) RJ_CUDA(cudaMemsetAsync(v, 0, lo * w, s));
            if (hi < rows) RJ_CUDA(cudaMemsetAsync(v + hi * w, 0, (rows - hi) * w, s));
        }
        if (type == RJ_VARCHAR) {
            launch_decode_varchar(pages, n_pages, row_start, out->values->as<uint64_t>(),
                                  need_valid ? out->valid->as<uint32_t>() : nullptr, ctx->sm_count, s, ctx->err_dev);
        } else {
            launch_decode_fixed(pages, n_pages, type, row_start, out->values->p,
                                need_valid ? out->valid->as<uint32_t>() : nullptr, ctx->sm_count, s);
        }
    }
}

const DecodedCol& Exec::column(uint32_t t, uint32_t c) {
    auto key = std::make_pair(t, c);
    DecodedMap& decoded = decoded_of(t);
    auto it = decoded.find(key);
    if (it != decoded.end()) return it->second;
    const TableDev&  td = in->tables[t];
    const ColumnDev& cd = td.cols[c];
    if (cd.windowed) {
        // decode every page of the chunk, shifted so that the window's first row starts a validity
        // word, then expose the window as a view
        if (cd.type == RJ_VARCHAR) throw EngineError("internal: VARCHAR columns are not streamed");
        const size_t   w     = type_width(cd.type);
        const uint32_t bias  = static_cast<uint32_t>((32 - cd.skip_rows % 32) % 32);
        const uint64_t first = bias + cd.skip_rows; // multiple of 32
        const uint64_t total = std::max<uint64_t>(bias + cd.page_rows, first + td.num_rows);
        DecodedCol full;
        decode_pages(ctx, s, cd.pages, cd.n_pages, cd.type, total, cd.window_nulls, &full, bias, cd.page_rows);
        DecodedCol d;
        d.type = cd.type;
        d.rows = td.num_rows;
        d.pages = cd.pages;
        d.hold_values = full.values;
        d.values = std::make_shared<DevMem>(full.values->as<uint8_t>() + first * w, td.num_rows * w);
        if (full.valid) {
            d.hold_valid = full.valid;
            d.valid = std::make_shared<DevMem>(full.valid->as<uint32_t>() + first / 32, ((td.num_rows + 31) / 32) * 4);
        }
        return decoded.emplace(key, std::move(d)).first->second;
    }
    if (cd.dense) {
        DecodedCol d;
        d.type = cd.type;
        d.rows = td.num_rows;
        const size_t w = cd.type == RJ_INT32 ? 4 : 8;
        d.values = std::make_shared<DevMem>(const_cast<void*>(cd.dense_values), td.num_rows * w);
        if (cd.dense_valid) d.valid = std::make_shared<DevMem>(const_cast<uint32_t*>(cd.dense_valid), ((td.num_rows + 31) / 32) * 4);
        return decoded.emplace(key, std::move(d)).first->second;
    }
    if (cd.n_pages && !cd.pages) throw EngineError("column was not uploaded");
    DecodedCol d;
    const bool need_valid = cd.non_null != td.num_rows;
    decode_pages(ctx, s, cd.pages, cd.n_pages, cd.type, td.num_rows, need_valid, &d, 0, cd.page_rows);
    return decoded.emplace(key, std::move(d)).first->second;
}

const DecodedCol& Exec::string_hash(uint32_t t, uint32_t c) {
    DecodedCol& d = const_cast<DecodedCol&>(column(t, c));
    if (!d.str_hash) {
        d.str_hash = dev_alloc(d.rows * 8, s);
        StageScope sc(ctx, RJ_ST_GATHER, s, 1, d.rows * 16);
        launch_varchar_hash(d.pages, d.values->as<uint64_t>(), d.valid_ptr(), d.rows, d.str_hash->as<uint64_t>(),
                            ctx->sm_count, s);
    }
    return d;
}

Buf Exec::gather_u32(const Buf& src, const Buf& idx, uint64_t n, uint32_t idx_mask) {
    Buf out = dev_alloc(n * 4, s);
    StageScope sc(ctx, RJ_ST_GATHER, s, 1, n * 12);
    launch_gather(src->p, nullptr, idx->as<uint32_t>(), n, 4, out->p, nullptr, ctx->sm_count, s, idx_mask);
    return out;
}

// ---- the join pipeline: histogram -> plan -> scatter (1 or 2 passes) -> shared-memory build+probe ------
// radix bits of the first of two scatter passes (RJ_PASS1_BITS overrides: profiling experiments)
int pass1_bits_of(int total_bits) {
    static const int forced = getenv("RJ_PASS1_BITS") ? atoi(getenv("RJ_PASS1_BITS")) : 0;
    if (forced > 0 && forced <= kMaxPassBits && total_bits - forced > 0 && total_bits - forced <= kMaxPassBits) return forced;
    return (total_bits + 1) / 2;
}

int choose_total_bits(uint64_t n_build) {
    if (n_build <= kJoinBuildCap) return 0; // small build side: one table, no partitioning
    int b = 0;
    while ((n_build >> b) > kJoinTargetFill && b < kMaxTotalBits) ++b;
    return b;
}

void Exec::join_keys(JoinSide& B, JoinSide& P, int key_bytes, uint64_t* n_out) {
    const void* bk = B.keys; const uint32_t* bv = B.valid; const uint64_t nb = B.n;
    const void* pk = P.keys; const uint32_t* pv = P.valid; const uint64_t np = P.n;
    if (nb >= 0xffffffffull || np >= 0xffffffffull) throw EngineError("relation exceeds 2^32-1 rows");
    const int      bits  = choose_total_bits(nb);
    const int      bits1 = bits > kMaxPassBits ? pass1_bits_of(bits) : 0; // two passes above 8 bits
    const int      bits2 = bits - bits1;
    const uint32_t nparts = 1u << bits;
    Buf plan_mem = dev_alloc(partition_plan_words(bits, bits1) * 4, s);
    PartitionPlanDev pl;
    partition_plan_carve(plan_mem->as<uint32_t>(), bits, bits1, &pl);

    JoinLaunch jl{};
    // probe tuples per work unit: 16384, less when that would leave most of the GPU without a unit (JOB: a 20 K-row
    // table probed by 1.4 M rows is 16 partitions; the table of a unit is rebuilt from <= 6144 tuples, cheap next to
    // 2048+ probes)
    uint32_t probe_chunk = kJoinProbeChunk;
    {
        const uint64_t want_units = 2ull * join_grid(ctx->sm_count);
        if (np / probe_chunk < want_units) {
            const uint64_t c = (np / want_units + 1023) & ~uint64_t(1023);
            probe_chunk = static_cast<uint32_t>(std::min<uint64_t>(kJoinProbeChunk, std::max<uint64_t>(2048, c)));
        }
    }
    jl.probe_chunk = probe_chunk;
    Buf keys_b, idx_b, keys_p, idx_p; // fully partitioned relations
    const uint64_t n_in = nb + np;
    if (bits == 0) {
        launch_partition_plan(nullptr, nullptr, static_cast<uint32_t>(nb), static_cast<uint32_t>(np), 0, 0, key_bytes, pl, s, kJoinBuildCap, probe_chunk);
        jl.bkeys = bk; jl.bidx = nullptr; jl.bvalid = bv;
        jl.pkeys = pk; jl.pidx = nullptr; jl.pvalid = pv;
    } else {
        Buf hist = dev_alloc_zero(size_t(2) * nparts * 4, s);
        uint32_t* hist_b = hist->as<uint32_t>();
        uint32_t* hist_p = hist_b + nparts;
        {
            StageScope sc(ctx, RJ_ST_HISTOGRAM, s, 2, n_in * key_bytes);
            launch_radix_histogram(bk, bv, nb, key_bytes, 0, bits, hist_b, ctx->sm_count, s);
            launch_radix_histogram(pk, pv, np, key_bytes, 0, bits, hist_p, ctx->sm_count, s);
        }
        launch_partition_plan(hist_b, hist_p, 0, 0, bits, bits1, key_bytes, pl, s, kJoinBuildCap, probe_chunk);
        // the row ids of a side travel only if something downstream will ask for them (two passes: the
        // final arrays then hold positions, which the join always needs)
        const bool two_pass = bits1 != 0;
        keys_b = dev_alloc(nb * key_bytes, s);
        if (two_pass || B.need_rows) idx_b = dev_alloc(nb * 4, s);
        keys_p = dev_alloc(np * key_bytes, s);
        if (two_pass || P.need_rows) idx_p = dev_alloc(np * 4, s);
        auto payload_of = [&](JoinSide& sd) {
            ScatterPayload pay;
            for (auto& c: sd.carry) {
                if (pay.n + (c.valid_src ? 2 : 1) > ScatterPayload::kMax) break;
                c.out = dev_alloc(sd.n * c.width, s);
                pay.src[pay.n] = c.src;
                pay.dst[pay.n] = c.out->p;
                pay.width[pay.n] = c.width;
                ++pay.n;
                if (c.valid_src) {
                    c.valid_out = dev_alloc(sd.n, s);
                    pay.src[pay.n] = c.valid_src;
                    pay.dst[pay.n] = c.valid_out->p;
                    pay.width[pay.n] = 1;
                    ++pay.n;
                }
            }
            return pay;
        };
        const ScatterPayload pay_b = payload_of(B), pay_p = payload_of(P);
        uint64_t carried_bytes = 0;
        for (auto& c: B.carry) if (c.out) carried_bytes += 2 * nb * (c.width + (c.valid_out ? 1 : 0));
        for (auto& c: P.carry) if (c.out) carried_bytes += 2 * np * (c.width + (c.valid_out ? 1 : 0));
        // SURVEY 8d numerator: one-pass scatter = N*w_k read + N*(w_k+4) written, whatever the pass count
        StageScope sc(ctx, RJ_ST_SCATTER, s, bits1 ? 4 : 2, n_in * (2 * key_bytes + 4) + carried_bytes);
        B.partitioned = P.partitioned = true;
        if (bits1 == 0) {
            // single pass: position order = final partition order
            launch_radix_scatter(bk, bv, nullptr, nb, key_bytes, 0, bits, pl.cur_b, keys_b->p, idx_b ? idx_b->as<uint32_t>() : nullptr, pay_b, ctx->sm_count, s);
            launch_radix_scatter(pk, pv, nullptr, np, key_bytes, 0, bits, pl.cur_p, keys_p->p, idx_p ? idx_p->as<uint32_t>() : nullptr, pay_p, ctx->sm_count, s);
            jl.bkeys = keys_b->p; jl.bidx = nullptr; jl.bvalid = nullptr; // emit positions
            jl.pkeys = keys_p->p; jl.pidx = nullptr; jl.pvalid = nullptr;
            B.rows_of_pos = idx_b; B.keys_of_pos = keys_b;
            P.rows_of_pos = idx_p; P.keys_of_pos = keys_p;
        } else {
            // pass 1 (high bits1 of the partition id) defines the position order: row ids, keys and the
            // carried payloads are written once, in regions of |side| / 2^bits1 tuples.  Pass 2 (low
            // bits2 inside each region) only moves (key, POSITION in the pass-1 arrays), so everything a
            // match refers to later lies inside one L2-sized region.
            Buf tk_b = dev_alloc(nb * key_bytes, s), ti_b = B.need_rows ? dev_alloc(nb * 4, s) : Buf();
            Buf tk_p = dev_alloc(np * key_bytes, s), ti_p = P.need_rows ? dev_alloc(np * 4, s) : Buf();
            launch_radix_scatter(bk, bv, nullptr, nb, key_bytes, bits2, bits1, pl.cur1_b, tk_b->p, ti_b ? ti_b->as<uint32_t>() : nullptr, pay_b, ctx->sm_count, s);
            launch_radix_scatter(pk, pv, nullptr, np, key_bytes, bits2, bits1, pl.cur1_p, tk_p->p, ti_p ? ti_p->as<uint32_t>() : nullptr, pay_p, ctx->sm_count, s);
            // validity of up to two carried columns per side rides in bits 30/31 of the positions
            auto flags_of = [&](JoinSide& sd) {
                RegionFlags f;
                if (sd.n >= kPosMask) return f;
                for (auto& c: sd.carry) {
                    if (c.valid_out && f.n < 2) {
                        c.valid_bit = 30 + f.n;
                        f.src[f.n++] = c.valid_out->as<uint8_t>();
                    }
                }
                if (f.n) sd.pos_mask = kPosMask;
                return f;
            };
            const RegionFlags fl_b = flags_of(B), fl_p = flags_of(P);
            launch_radix_scatter_regions(tk_b->p, nullptr, pl.reg_b, pl.tile_b, 1u << bits1, nb, key_bytes, 0, bits2,
                                         pl.cur_b, keys_b->p, idx_b->as<uint32_t>(), fl_b, ScatterPayload{}, ctx->sm_count, s);
            launch_radix_scatter_regions(tk_p->p, nullptr, pl.reg_p, pl.tile_p, 1u << bits1, np, key_bytes, 0, bits2,
                                         pl.cur_p, keys_p->p, idx_p->as<uint32_t>(), fl_p, ScatterPayload{}, ctx->sm_count, s);
            jl.bkeys = keys_b->p; jl.bidx = idx_b->as<uint32_t>(); jl.bvalid = nullptr; // idx = pass-1 position
            jl.pkeys = keys_p->p; jl.pidx = idx_p->as<uint32_t>(); jl.pvalid = nullptr;
            B.rows_of_pos = ti_b; B.keys_of_pos = tk_b;
            P.rows_of_pos = ti_p; P.keys_of_pos = tk_p;
        }
    }
    jl.off_b = pl.off_b; jl.off_p = pl.off_p; jl.unit_start = pl.unit_start; jl.unit_cursor = pl.unit_cursor;
    jl.nparts = nparts; jl.part_bits = bits; jl.key_bytes = key_bytes;

    Buf counter = dev_alloc(8, s);
    jl.out_count = counter->as<unsigned long long>();
    // scratch of the duplicate chains (4-byte keys): written before it is read, no initialisation
    Buf dup_next, dup_head;
    if (key_bytes == 4) {
        dup_next = dev_alloc(size_t(join_grid(ctx->sm_count)) * kJoinBuildCap * 4, s);
        dup_head = dev_alloc(size_t(join_grid(ctx->sm_count)) * kJoinSlots * 4, s);
        jl.dup_next = dup_next->as<uint32_t>();
        jl.dup_head = dup_head->as<uint32_t>();
    }
    // Output cardinality is unknown (non-unique keys on both sides are legal).  Run with room for
    // max(|build|, |probe|) pairs -- enough for every key/foreign-key join -- and let the kernel keep
    // counting when that overflows; the exact count then sizes a second run.
    uint64_t capacity = std::max(nb, np);
    uint64_t matches = 0;
    auto t_join0 = std::chrono::steady_clock::now();
    for (int attempt = 0; attempt < 2; ++attempt) {
        B.pos = dev_alloc(capacity * 4, s);
        P.pos = dev_alloc(capacity * 4, s);
        jl.out_b = B.pos->as<uint32_t>();
        jl.out_p = P.pos->as<uint32_t>();
        jl.capacity = capacity;
        RJ_CUDA(cudaMemsetAsync(counter->p, 0, 8, s));
        RJ_CUDA(cudaMemsetAsync(pl.unit_cursor, 0, 4, s));
        if (getenv("RJ_TRACE")) {
            RJ_CUDA(cudaStreamSynchronize(s));
            t_join0 = std::chrono::steady_clock::now();
        }
        {
            // N*(w_k+4) read + M*8 written (M is added once known)
            StageScope sc(ctx, RJ_ST_JOIN, s, 1, n_in * (key_bytes + 4));
            launch_join(jl, ctx->sm_count, s);
        }
        RJ_CUDA(cudaMemcpyAsync(&matches, counter->p, 8, cudaMemcpyDeviceToHost, s));
        RJ_CUDA(cudaStreamSynchronize(s));
        if (matches <= capacity) break;
        if (attempt == 1) throw EngineError("join output overflowed twice");
        capacity = matches;
    }
    if (ctx->profiling) ctx->stats[RJ_ST_JOIN].bytes += matches * 8;
    static const bool trace_joins = getenv("RJ_TRACE") != nullptr;
    if (trace_joins) fprintf(stderr, "[rj] join: table side %llu rows, probe side %llu rows, %d radix bits (%d + %d), %llu matches, kernel + sync %.3f ms\n",
                             (unsigned long long)nb, (unsigned long long)np, bits, bits1, bits2, (unsigned long long)matches,
                             std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_join0).count());
    if (matches >= 0xffffffffull) throw EngineError("join result exceeds 2^32-1 rows");
    *n_out = matches;
}

// row index (into the side's input relation) of every match
Buf Exec::side_rows(const JoinSide& sd, uint64_t m) {
    if (sd.partitioned && !sd.rows_of_pos) throw EngineError("internal: row ids of a join side were not kept");
    return sd.partitioned ? gather_u32(sd.rows_of_pos, sd.pos, m, sd.pos_mask) : sd.pos;
}

Buf Exec::rid_of(const Rel& r, int leaf) {
    auto it = r.rid.find(leaf);
    if (it != r.rid.end()) return it->second;
    auto lz = r.lazy.find(leaf);
    if (lz == r.lazy.end()) throw EngineError("internal: leaf not tracked");
    if (!lz->second.rows_of_pos) throw EngineError("internal: row ids of a join side were not kept");
    Buf rows = gather_u32(lz->second.rows_of_pos, lz->second.pos, r.rows, lz->second.mask);
    r.rid[leaf] = rows;
    return rows;
}

struct SideKeys {
    const void*     keys = nullptr;
    const uint32_t* valid = nullptr;
    Buf             hold_k, hold_v;
    int             key_bytes = 4;
};

Rel Exec::join(uint64_t n, const Rel& L, const Rel& R) {
    const rj_node_t& nd = node(n);
    Rel out;
    // key type = the BUILD side's declared attribute type (src/execute.cpp:271-273)
    const rj_node_t& bnode = node(nd.build_left ? nd.left : nd.right);
    const uint64_t   battr = nd.build_left ? nd.left_attr : nd.right_attr;
    if (battr >= bnode.n_output_attrs) throw EngineError("join attribute out of range");
    const int key_type = bnode.output_attrs[battr].type;
    if (key_type < RJ_INT32 || key_type > RJ_VARCHAR) throw EngineError("Unsupported join type"); // :280
    const Attr la = resolve(nd.left, nd.left_attr), ra = resolve(nd.right, nd.right_attr);
    if (L.rows == 0 || R.rows == 0) return out; // :50
    // a cell is a valid key only if it physically holds the key type (:61-83)
    if (in->tables[la.table].cols[la.col].type != key_type || in->tables[ra.table].cols[ra.col].type != key_type) return out;
    if (key_type == RJ_FP64) {
        // HashUtil<double>::hash recurses forever in the reference (:28-31); there is no behaviour to match
        throw EngineError("Unsupported join type: FP64 join key (the reference does not terminate on it)");
    }

    auto side_keys = [&](const Rel& rel, const Attr& a) {
        SideKeys k;
        const DecodedCol& col = key_type == RJ_VARCHAR ? string_hash(a.table, a.col) : column(a.table, a.col);
        const void* base = key_type == RJ_VARCHAR ? col.str_hash->p : col.values->p;
        k.key_bytes = key_type == RJ_INT32 ? 4 : 8;
        if (!rel.has_leaf(a.leaf)) throw EngineError("internal: join key leaf not tracked");
        Buf rid = rid_of(rel, a.leaf);
        if (!rid) { // scan: the decoded column is the key column
            k.keys = base;
            k.valid = col.valid_ptr();
            return k;
        }
        k.hold_k = dev_alloc(rel.rows * k.key_bytes, s);
        if (col.valid) k.hold_v = dev_alloc(((rel.rows + 31) / 32) * 4, s);
        StageScope sc(ctx, RJ_ST_GATHER, s, 1, rel.rows * (4 + 2 * k.key_bytes));
        launch_gather(base, col.valid_ptr(), rid->as<uint32_t>(), rel.rows, k.key_bytes, k.hold_k->p,
                      k.hold_v ? k.hold_v->as<uint32_t>() : nullptr, ctx->sm_count, s);
        k.keys = k.hold_k->p;
        k.valid = k.hold_v ? k.hold_v->as<uint32_t>() : nullptr;
        return k;
    };
    SideKeys lk = side_keys(L, la), rk = side_keys(R, ra);

    JoinSide ls, rs;
    ls.keys = lk.keys; ls.valid = lk.valid; ls.n = L.rows;
    rs.keys = rk.keys; rs.valid = rk.valid; rs.n = R.rows;
    // At the ROOT join, fixed-width output columns of a child that is a plain scan travel through the
    // scatter with the tuples ("carried"): the root then encodes them from position order, where a
    // match's neighbours are its partition's tuples, instead of gathering 8 bytes per 128-byte DRAM
    // line from the whole table through row ids.
    const bool is_root = n == plan->root;
    auto pure_scan = [&](const Rel& rel, const Attr& a) {
        auto it = rel.rid.find(a.leaf);
        return rel.rid.size() == 1 && rel.lazy.empty() && it != rel.rid.end() && !it->second;
    };
    auto plan_carry = [&](const Rel& rel, const Attr& key_attr, JoinSide& sd) {
        if (!is_root || key_type == RJ_VARCHAR || !pure_scan(rel, key_attr)) return;
        std::set<uint32_t> seen;
        for (uint32_t a = 0; a < nd.n_output_attrs; ++a) {
            const Attr at = resolve(n, a);
            if (at.leaf != key_attr.leaf || at.col == key_attr.col || seen.count(at.col)) continue;
            const int t = in->tables[at.table].cols[at.col].type;
            if (t == RJ_VARCHAR || t != nd.output_attrs[a].type) continue;
            seen.insert(at.col);
            CarryCol c;
            c.col = at.col;
            const DecodedCol& dc = column(at.table, at.col);
            c.src = dc.values->p;
            c.valid_src = dc.valid_ptr();
            c.width = t == RJ_INT32 ? 4 : 8;
            sd.carry.push_back(c);
        }
    };
    plan_carry(L, la, ls);
    plan_carry(R, ra, rs);
    // Row ids of a side are needed unless this is the root, the side is a plain scan and every root
    // output from it is the join key or a column the scatter will carry (same slot budget as join_keys).
    auto rows_needed = [&](const Rel& rel, const Attr& key_attr, const JoinSide& sd) {
        if (!is_root || key_type == RJ_VARCHAR || !pure_scan(rel, key_attr)) return true;
        std::set<uint32_t> moved;
        int slots = 0;
        for (auto& c: sd.carry) {
            const int want = c.valid_src ? 2 : 1;
            if (slots + want > ScatterPayload::kMax) break;
            slots += want;
            moved.insert(c.col);
        }
        for (uint32_t a = 0; a < nd.n_output_attrs; ++a) {
            const Attr at = resolve(n, a);
            if (at.leaf == key_attr.leaf && at.col != key_attr.col && !moved.count(at.col)) return true;
        }
        return false;
    };
    ls.need_rows = rows_needed(L, la, ls);
    rs.need_rows = rows_needed(R, ra, rs);

    // The hash table goes on the SMALLER side whatever build_left says: the result is the same
    // multiset of (left row, right row) pairs, and a small table side means fewer partitions.
    // (a window of a streamed table counts as the whole table: a window of a foreign-key side is smaller
    // than the key side it refers to, but full of duplicates)
    const double l_size = static_cast<double>(L.rows) * (L.streamed ? streamed_scale : 1.0);
    const double r_size = static_cast<double>(R.rows) * (R.streamed ? streamed_scale : 1.0);
    const bool table_left = l_size <= r_size;
    uint64_t m = 0;
    if (table_left) {
        join_keys(ls, rs, lk.key_bytes, &m);
    } else {
        join_keys(rs, ls, lk.key_bytes, &m);
    }

    Buf l_rows, r_rows; // row index into L / R of every match (materialised only when needed)
    auto rows_l = [&]() { if (!l_rows) l_rows = side_rows(ls, m); return l_rows; };
    auto rows_r = [&]() { if (!r_rows) r_rows = side_rows(rs, m); return r_rows; };

    if (key_type == RJ_VARCHAR && m > 0) {
        // the join ran on 64-bit string hashes: keep only pairs whose strings are byte-equal
        const DecodedCol& lc = column(la.table, la.col);
        const DecodedCol& rc = column(ra.table, ra.col);
        Buf l_idx = rows_l(), r_idx = rows_r();
        Buf lrow = rid_of(L, la.leaf) ? gather_u32(rid_of(L, la.leaf), l_idx, m) : l_idx;
        Buf rrow = rid_of(R, ra.leaf) ? gather_u32(rid_of(R, ra.leaf), r_idx, m) : r_idx;
        Buf keep = dev_alloc(m * 4, s), pos = dev_alloc((m + 1) * 8, s), tmp = dev_alloc(scan_tmp_bytes(m), s);
        StageScope sc(ctx, RJ_ST_GATHER, s, 5, m * 32);
        launch_varchar_pairs_equal(lc.pages, lc.values->as<uint64_t>(), lrow->as<uint32_t>(), rc.pages,
                                   rc.values->as<uint64_t>(), rrow->as<uint32_t>(), m, keep->as<uint32_t>(), ctx->sm_count, s);
        launch_exclusive_scan_u32_u64(keep->as<uint32_t>(), pos->as<uint64_t>(), m, tmp->p, s);
        uint64_t kept = 0;
        RJ_CUDA(cudaMemcpyAsync(&kept, pos->as<uint64_t>() + m, 8, cudaMemcpyDeviceToHost, s));
        RJ_CUDA(cudaStreamSynchronize(s));
        Buf nl = dev_alloc(kept * 4, s), nr = dev_alloc(kept * 4, s);
        launch_compact_pairs(l_idx->as<uint32_t>(), r_idx->as<uint32_t>(), keep->as<uint32_t>(), pos->as<uint64_t>(), m,
                             nl->as<uint32_t>(), nr->as<uint32_t>(), s);
        // from here on the sides are plain row lists again
        l_rows = nl;
        r_rows = nr;
        ls.partitioned = rs.partitioned = false;
        ls.pos = nl;
        rs.pos = nr;
        ls.carry.clear();
        rs.carry.clear();
        m = kept;
    }
    out.rows = m;
    out.streamed = L.streamed || R.streamed;
    if (m == 0) return out;
    // row-id lists of every scan the parents can still see through this node's output_attrs
    std::set<int> needed;
    for (uint32_t a = 0; a < nd.n_output_attrs; ++a) needed.insert(resolve(n, a).leaf);
    auto attach = [&](int leaf, const Rel& child, JoinSide& sd, const Attr& key_attr, const std::function<Buf()>& rows) {
        auto ci = child.rid.find(leaf);
        const bool identity = ci != child.rid.end() && !ci->second && child.rid.size() == 1 && child.lazy.empty();
        if (identity) {
            // the child is a scan of this leaf: its row index IS the row id
            if (sd.partitioned) {
                out.lazy[leaf] = PosSpace{sd.pos, sd.rows_of_pos, sd.pos_mask};
                // the join key and the carried columns are available in position order
                CarriedCol kc;
                kc.hold = sd.keys_of_pos;
                kc.values_of_pos = sd.keys_of_pos->p;
                kc.pos = sd.pos;
                kc.pos_mask = sd.pos_mask;
                kc.never_null = true;
                if (key_type != RJ_VARCHAR && leaf == key_attr.leaf) out.carried[{leaf, key_attr.col}] = kc;
                for (auto& c: sd.carry) {
                    if (!c.out) continue;
                    CarriedCol cc;
                    cc.hold = c.out;
                    cc.values_of_pos = c.out->p;
                    cc.valid_hold = c.valid_out;
                    cc.pos = sd.pos;
                    cc.pos_mask = sd.pos_mask;
                    cc.valid_bit = c.valid_bit;
                    out.carried[{leaf, c.col}] = cc;
                }
            } else {
                out.rid[leaf] = sd.pos;
            }
            return;
        }
        out.rid[leaf] = gather_u32(rid_of(child, leaf), rows(), m);
    };
    for (int leaf: needed) {
        if (L.has_leaf(leaf)) {
            attach(leaf, L, ls, la, rows_l);
        } else if (R.has_leaf(leaf)) {
            attach(leaf, R, rs, ra, rows_r);
        } else {
            throw EngineError("internal: output leaf not below this join");
        }
    }
    return out;
}

Rel Exec::run(uint64_t n) {
    const rj_node_t& nd = node(n);
    if (!nd.is_join) {
        if (nd.base_table_id >= in->tables.size()) throw EngineError("base table out of range");
        Rel r;
        r.rows = in->tables[nd.base_table_id].num_rows;
        r.streamed = nd.base_table_id == streamed_table;
        r.rid[static_cast<int>(n)] = nullptr; // identity
        return r;
    }
    if (nd.left >= plan->n_nodes || nd.right >= plan->n_nodes) throw EngineError("join child out of range");
    Rel L = run(nd.left);  // depth first, left then right (src/execute.cpp:48-49)
    Rel R = run(nd.right);
    return join(n, L, R);
}

// ---- page output -----------------------------------------------------------------------------------
ResultColumn Exec::encode_varchar(const DecodedCol& col, const uint32_t* idx, uint64_t n) {
    ResultColumn rc;
    rc.type = RJ_VARCHAR;
    VarcharLayoutDev L;
    L.n = n;
    L.src_pages = col.pages;
    L.desc = col.values->as<uint64_t>();
    L.valid = col.valid_ptr();
    L.idx = idx;
    Buf weights = dev_alloc(n * 8, s), wscan = dev_alloc(n * 8, s), marks = dev_alloc(n * 8, s), bscan = dev_alloc(n * 8, s);
    Buf heads = dev_alloc(n * 4, s), page_of = dev_alloc((n + 1) * 8, s), scalars = dev_alloc_zero(32, s);
    Buf tmp = dev_alloc(scan_tmp_bytes(n), s);
    L.weight_scan = wscan->as<uint64_t>();
    L.base_scan = bscan->as<uint64_t>();
    L.head_pages = heads->as<uint32_t>();
    L.page_of = page_of->as<uint64_t>();
    L.scalars = scalars->as<uint64_t>();
    uint64_t n_pages = 0;
    {
        StageScope sc(ctx, RJ_ST_ENCODE, s, 13, n * 48);
        launch_varchar_weights(L, weights->as<uint64_t>(), ctx->sm_count, s);
        launch_inclusive_sum_u64(weights->as<uint64_t>(), L.weight_scan, n, tmp->p, s);
        launch_varchar_marks(L, marks->as<uint64_t>(), ctx->sm_count, s);
        launch_inclusive_max_u64(marks->as<uint64_t>(), L.base_scan, n, tmp->p, s);
        launch_varchar_heads(L, weights->as<uint64_t>(), ctx->sm_count, s);
        launch_exclusive_scan_u32_u64(L.head_pages, L.page_of, n, tmp->p, s);
        RJ_CUDA(cudaMemcpyAsync(&n_pages, L.page_of + n, 8, cudaMemcpyDeviceToHost, s));
    }
    RJ_CUDA(cudaStreamSynchronize(s));
    L.n_pages = n_pages;
    Buf page_row = dev_alloc(n_pages * 4, s);
    L.page_row = page_row->as<uint32_t>();
    rc.n_pages = n_pages;
    rc.pages = dev_alloc(n_pages * size_t(RJ_PAGE_SIZE), s);
    {
        StageScope sc(ctx, RJ_ST_ENCODE, s, 2, n_pages * uint64_t(RJ_PAGE_SIZE) * 2);
        launch_varchar_page_rows(L, ctx->sm_count, s);
        launch_varchar_write(L, rc.pages->as<uint8_t>(), ctx->sm_count, s);
    }
    return rc;
}

std::unique_ptr<rj_result> Exec::root(uint64_t n, const Rel& r) {
    const rj_node_t& nd = node(n);
    auto res = std::make_unique<rj_result>();
    res->num_rows = r.rows;
    res->cols.resize(nd.n_output_attrs);
    for (uint32_t a = 0; a < nd.n_output_attrs; ++a) {
        ResultColumn& rc = res->cols[a];
        rc.type = nd.output_attrs[a].type;
        if (rc.type < RJ_INT32 || rc.type > RJ_VARCHAR) throw EngineError("unknown output attribute type");
        if (r.rows == 0) {
            resolve(n, a); // still validate the plan
            continue;      // typed, page-less column (tests/unit_tests.cpp:24-27)
        }
        const Attr at = resolve(n, a);
        if (in->tables[at.table].cols[at.col].type != rc.type) {
            // The column does not physically hold the declared type.  Table::to_columnar visits every
            // cell: a fixed-width column keeps the NULL cells and silently SKIPS the others
            // (build_table.cpp:484-501, :527-544, :570-587), so it comes out holding only the NULL rows;
            // a VARCHAR column throws "not string or null" at the first non-NULL cell (:667-669).
            const DecodedCol& col = column(at.table, at.col);
            if (!r.has_leaf(at.leaf)) throw EngineError("internal: root leaf not tracked");
            uint64_t n_null = 0;
            if (col.valid) {
                Buf rid = rid_of(r, at.leaf);
                Buf cnt = dev_alloc_zero(8, s);
                launch_count_nulls(col.valid_ptr(), rid ? rid->as<uint32_t>() : nullptr, r.rows, 0xffffffffu,
                                   cnt->as<unsigned long long>(), ctx->sm_count, s);
                RJ_CUDA(cudaMemcpyAsync(&n_null, cnt->p, 8, cudaMemcpyDeviceToHost, s));
                RJ_CUDA(cudaStreamSynchronize(s));
            }
            if (rc.type == RJ_VARCHAR && n_null != r.rows) throw EngineError("not string or null");
            // n_null all-NULL rows: header n_r, n_v = 0, zero bitmap -- the same bytes for every type
            const uint32_t rpp = rj_fixed_rows_per_page(RJ_INT32);
            rc.n_pages = (n_null + rpp - 1) / rpp;
            if (n_null == 0) continue;
            rc.pages = dev_alloc(rc.n_pages * size_t(RJ_PAGE_SIZE), s);
            Buf no_valid = dev_alloc_zero(((n_null + 31) / 32 + 1) * 4, s);
            StageScope sc(ctx, RJ_ST_ENCODE, s, 1, rc.n_pages * uint64_t(RJ_PAGE_SIZE));
            launch_encode_fixed(no_valid->p, no_valid->as<uint32_t>(), nullptr, nullptr, nullptr, n_null, RJ_INT32, rc.pages->p, ctx->sm_count, s);
            continue;
        }
        const DecodedCol& col = column(at.table, at.col);
        if (!r.has_leaf(at.leaf)) throw EngineError("internal: root leaf not tracked");
        if (rc.type == RJ_VARCHAR) {
            Buf rid = rid_of(r, at.leaf);
            rc = encode_varchar(col, rid ? rid->as<uint32_t>() : nullptr, r.rows);
            continue;
        }
        const uint32_t rpp = rj_fixed_rows_per_page(rc.type);
        rc.n_pages = (r.rows + rpp - 1) / rpp;
        rc.pages = dev_alloc(rc.n_pages * size_t(RJ_PAGE_SIZE), s);
        const size_t w = type_width(rc.type);
        const void*     values = col.values->p;
        const uint32_t* idx = nullptr;
        const uint32_t* vidx = nullptr;
        const uint8_t*  valid_bytes = nullptr;
        uint32_t        idx_mask = 0xffffffffu;
        int             valid_bit = -1;
        Buf             hold_rid;
        auto carried = r.carried.find({at.leaf, at.col});
        if (carried != r.carried.end()) {
            // values in position order; validity (if the column has NULLs) still lives at the row id
            values = carried->second.values_of_pos;
            idx = carried->second.pos->as<uint32_t>();
            vidx = idx;
            idx_mask = carried->second.pos_mask;
            if (carried->second.valid_bit >= 0) {
                valid_bit = carried->second.valid_bit; // validity rides in the position: no gather at all
            } else if (carried->second.valid_hold) {
                valid_bytes = carried->second.valid_hold->as<uint8_t>();
            } else if (col.valid && !carried->second.never_null) {
                hold_rid = rid_of(r, at.leaf);
                vidx = hold_rid ? hold_rid->as<uint32_t>() : nullptr;
            }
        } else {
            hold_rid = rid_of(r, at.leaf);
            idx = vidx = hold_rid ? hold_rid->as<uint32_t>() : nullptr;
        }
        // M*4 + M*w read + 8192 * pages written
        StageScope sc(ctx, RJ_ST_ENCODE, s, 1, r.rows * (4 + w) + rc.n_pages * uint64_t(RJ_PAGE_SIZE));
        const bool all_valid = carried != r.carried.end() && carried->second.never_null;
        launch_encode_fixed(values, all_valid ? nullptr : col.valid_ptr(), valid_bytes, idx, vidx, r.rows, rc.type, rc.pages->p, ctx->sm_count, s, idx_mask, valid_bit);
    }
    return res;
}

// ---- root join with page output fused into the join kernel (k_join_emit.cu) -----------------------------
// Taken when the root is a join of two scans on an INT32 key whose output columns are the key and at most
// kEmitMaxPay fixed-width columns per side: both sides are partitioned to their final order WITH those
// columns beside the keys (two scatter passes carry values and validity bytes), and the join kernel
// writes result pages itself.  Returns null when the plan does not qualify or a table met a duplicate
// build key; the caller then runs the general path.
std::unique_ptr<rj_result> Exec::root_fused(uint64_t n) {
    if (getenv("RJ_NO_FUSED_ROOT") != nullptr) return nullptr; // tests and profiling: force the general path
    const rj_node_t& nd = node(n);
    if (!nd.is_join || nd.left >= plan->n_nodes || nd.right >= plan->n_nodes) return nullptr;
    const rj_node_t &ln = node(nd.left), &rn = node(nd.right);
    if (ln.is_join || rn.is_join) return nullptr;
    if (nd.n_output_attrs < 1 || nd.n_output_attrs > static_cast<uint32_t>(kEmitMaxOut)) return nullptr;
    if (nd.left_attr >= ln.n_output_attrs || nd.right_attr >= rn.n_output_attrs) return nullptr;
    const rj_node_t& bnode = nd.build_left ? ln : rn;
    if (bnode.output_attrs[nd.build_left ? nd.left_attr : nd.right_attr].type != RJ_INT32) return nullptr; // key type, :271-273
    if (ln.base_table_id >= in->tables.size() || rn.base_table_id >= in->tables.size()) return nullptr;
    for (uint32_t a = 0; a < nd.n_output_attrs; ++a) {
        const uint64_t src = nd.output_attrs[a].index;
        if (src >= uint64_t(ln.n_output_attrs) + rn.n_output_attrs) return nullptr;
    }
    const Attr la = resolve(nd.left, nd.left_attr), ra = resolve(nd.right, nd.right_attr);
    const TableDev &tl = in->tables[la.table], &tr = in->tables[ra.table];
    if (tl.num_rows == 0 || tr.num_rows == 0) return nullptr;
    if (tl.cols[la.col].type != RJ_INT32 || tr.cols[ra.col].type != RJ_INT32) return nullptr;
    if (tl.num_rows >= 0xffffffffull || tr.num_rows >= 0xffffffffull) return nullptr;
    // the table goes on the smaller side, as in the general path
    const double l_size = static_cast<double>(tl.num_rows) * (la.table == streamed_table ? streamed_scale : 1.0);
    const double r_size = static_cast<double>(tr.num_rows) * (ra.table == streamed_table ? streamed_scale : 1.0);
    const bool   table_left = l_size <= r_size;
    const Attr&  ba = table_left ? la : ra;
    const Attr&  pa = table_left ? ra : la;
    const uint64_t nb = (table_left ? tl : tr).num_rows, np = (table_left ? tr : tl).num_rows;
    int bits = 0;
    while ((nb >> bits) > kJoinTargetFill && bits < kMaxTotalBits) ++bits;
    if (bits == 0) return nullptr; // a single small table: nothing to win

    // output columns -> the key or a carried column of one side
    struct Carry {
        uint32_t col;
        int      width;
    };
    std::vector<Carry> bcols, pcols;
    JoinEmitLaunch L;
    L.n_out = static_cast<int>(nd.n_output_attrs);
    for (uint32_t a = 0; a < nd.n_output_attrs; ++a) {
        const Attr at = resolve(n, a);
        const int  t  = in->tables[at.table].cols[at.col].type;
        if (t == RJ_VARCHAR || t != nd.output_attrs[a].type) return nullptr;
        const bool on_build = at.leaf == ba.leaf;
        if (!on_build && at.leaf != pa.leaf) return nullptr;
        L.out_width[a] = t == RJ_INT32 ? 4 : 8;
        if (at.col == (on_build ? ba.col : pa.col)) {
            L.out_src[a] = 0; // the join key: equal on both sides, never NULL in a match
            continue;
        }
        auto& list = on_build ? bcols : pcols;
        size_t i = 0;
        while (i < list.size() && list[i].col != at.col) ++i;
        if (i == list.size()) {
            if (list.size() == static_cast<size_t>(kEmitMaxPay)) return nullptr;
            list.push_back({at.col, L.out_width[a]});
        }
        L.out_src[a] = on_build ? 1 : 2;
        L.out_idx[a] = static_cast<int>(i);
    }
    const uint32_t bt = ba.table, pt = pa.table;
    const DecodedCol& bkey = column(bt, ba.col);
    const DecodedCol& pkey = column(pt, pa.col);
    std::vector<const DecodedCol*> bdec, pdec;
    for (auto& c: bcols) bdec.push_back(&column(bt, c.col));
    for (auto& c: pcols) pdec.push_back(&column(pt, c.col));
    L.n_bpay = static_cast<int>(bcols.size());
    L.n_ppay = static_cast<int>(pcols.size());
    for (int c = 0; c < L.n_bpay; ++c) {
        L.bwidth[c] = bcols[c].width;
        L.bvalid[c] = bdec[c]->valid ? reinterpret_cast<const uint8_t*>(1) : nullptr; // placeholder: "has NULLs" (for the layout)
    }
    for (int c = 0; c < L.n_ppay; ++c) L.pwidth[c] = pcols[c].width;
    for (int a = 0; a < L.n_out; ++a) {
        const DecodedCol* d = L.out_src[a] == 1 ? bdec[L.out_idx[a]] : (L.out_src[a] == 2 ? pdec[L.out_idx[a]] : nullptr);
        L.out_nullable[a] = d && d->valid ? 1 : 0;
    }
    if (!join_emit_fits(L)) return nullptr;
    {
        // 1-D TMA windows of the scatter need 16-byte aligned sources (adopted dense columns may not be)
        auto aligned = [](const DecodedCol& d) { return reinterpret_cast<uintptr_t>(d.values->p) % 16 == 0; };
        bool ok = aligned(bkey) && aligned(pkey);
        for (auto* d: bdec) ok = ok && aligned(*d);
        for (auto* d: pdec) ok = ok && aligned(*d);
        if (!ok) return nullptr;
    }

    // ---- histogram, plan, scatter both sides with their columns to the final partition order -------------
    const int      bits1 = bits > kMaxPassBits ? pass1_bits_of(bits) : 0;
    const int      bits2 = bits - bits1;
    const uint32_t nparts = 1u << bits;
    Buf plan_mem = dev_alloc(partition_plan_words(bits, bits1) * 4, s);
    PartitionPlanDev pl;
    partition_plan_carve(plan_mem->as<uint32_t>(), bits, bits1, &pl);
    Buf hist = dev_alloc_zero(size_t(2) * nparts * 4, s);
    {
        StageScope sc(ctx, RJ_ST_HISTOGRAM, s, 2, (nb + np) * 4);
        launch_radix_histogram(bkey.values->p, bkey.valid_ptr(), nb, 4, 0, bits, hist->as<uint32_t>(), ctx->sm_count, s);
        launch_radix_histogram(pkey.values->p, pkey.valid_ptr(), np, 4, 0, bits, hist->as<uint32_t>() + nparts, ctx->sm_count, s);
    }
    launch_partition_plan(hist->as<uint32_t>(), hist->as<uint32_t>() + nparts, 0, 0, bits, bits1, 4, pl, s, kEmitBuildCap, kEmitProbeChunk);

    struct SideOut {
        Buf keys;
        Buf val[kEmitMaxPay], ok[kEmitMaxPay];
    };
    uint64_t carried_bytes = 0;
    auto scatter_side = [&](const DecodedCol& key, uint64_t rows, const std::vector<Carry>& cols, const std::vector<const DecodedCol*>& dec,
                            uint32_t* cur1, uint32_t* cur2, const uint32_t* reg, const uint32_t* tile) {
        // k_scatter_carry.cu: keys, values and validity move together; nothing else does (the join needs
        // neither row ids nor positions)
        SideOut o;
        o.keys = dev_alloc(rows * 4 + 64, s);
        CarryScatter c1;
        c1.keys = key.values->as<uint32_t>();
        c1.valid = key.valid_ptr();
        c1.n = rows;
        c1.keys_out = o.keys->as<uint32_t>();
        for (size_t c = 0; c < cols.size(); ++c) {
            o.val[c] = dev_alloc(rows * cols[c].width + 64, s);
            c1.val_src[c1.n_val] = dec[c]->values->p;
            c1.val_dst[c1.n_val] = o.val[c]->p;
            c1.val_width[c1.n_val++] = cols[c].width;
            carried_bytes += 2 * rows * cols[c].width;
            if (dec[c]->valid) {
                o.ok[c] = dev_alloc(rows + 64, s);
                c1.flag_src[c1.n_flag] = dec[c]->valid_ptr(); // a bitmap by row becomes one byte per tuple
                c1.flag_dst[c1.n_flag++] = o.ok[c]->as<uint8_t>();
                carried_bytes += 2 * rows;
            }
        }
        if (bits1 == 0) {
            c1.shift = 0; c1.bits = bits; c1.cursor = cur2;
            launch_scatter_carry(c1, ctx->sm_count, s);
            return o;
        }
        c1.shift = bits2; c1.bits = bits1; c1.cursor = cur1;
        launch_scatter_carry(c1, ctx->sm_count, s);
        // pass 2, inside each pass-1 region, to the final order
        SideOut f;
        f.keys = dev_alloc(rows * 4 + 64, s);
        CarryScatter c2;
        c2.keys = o.keys->as<uint32_t>();
        c2.n = rows;
        c2.region_start = reg; c2.tile_start = tile; c2.n_regions = 1u << bits1;
        c2.shift = 0; c2.bits = bits2; c2.cursor = cur2;
        c2.keys_out = f.keys->as<uint32_t>();
        for (size_t c = 0; c < cols.size(); ++c) {
            f.val[c] = dev_alloc(rows * cols[c].width + 64, s);
            c2.val_src[c2.n_val] = o.val[c]->p;
            c2.val_dst[c2.n_val] = f.val[c]->p;
            c2.val_width[c2.n_val++] = cols[c].width;
            if (o.ok[c]) {
                f.ok[c] = dev_alloc(rows + 64, s);
                c2.flag_src[c2.n_flag] = o.ok[c]->p;
                c2.flag_dst[c2.n_flag++] = f.ok[c]->as<uint8_t>();
            }
        }
        launch_scatter_carry(c2, ctx->sm_count, s);
        // (the pass-1 arrays go back to the block cache here: blocks are reused in stream order)
        return f;
    };
    SideOut B, P;
    {
        StageScope sc(ctx, RJ_ST_SCATTER, s, bits1 ? 4 : 2, 0);
        B = scatter_side(bkey, nb, bcols, bdec, pl.cur1_b, pl.cur_b, pl.reg_b, pl.tile_b);
        P = scatter_side(pkey, np, pcols, pdec, pl.cur1_p, pl.cur_p, pl.reg_p, pl.tile_p);
        // SURVEY 8d numerator: one-pass scatter = N*w_k read + N*(w_k+4) written, whatever the pass count, plus
        // the carried columns once (read + written)
        if (ctx->profiling) ctx->stats[RJ_ST_SCATTER].bytes += (nb + np) * 12 + carried_bytes;
    }

    // ---- join + page output -------------------------------------------------------------------------------
    const uint64_t max_chunks = join_emit_max_chunks(np, nparts, ctx->sm_count);
    auto res = std::make_unique<rj_result>();
    res->cols.resize(nd.n_output_attrs);
    Buf counters = dev_alloc_zero(32, s); // chunk counter @0, abort flag @4, rows @8
    L.bkeys = B.keys->as<uint32_t>();
    L.pkeys = P.keys->as<uint32_t>();
    L.off_b = pl.off_b; L.off_p = pl.off_p; L.unit_start = pl.unit_start; L.unit_cursor = pl.unit_cursor; L.unit_part = pl.unit_part;
    L.nparts = nparts; L.part_bits = bits;
    for (int c = 0; c < L.n_bpay; ++c) {
        L.bpay[c] = B.val[c]->p;
        L.bvalid[c] = B.ok[c] ? B.ok[c]->as<uint8_t>() : nullptr;
    }
    for (int c = 0; c < L.n_ppay; ++c) {
        L.ppay[c] = P.val[c]->p;
        L.pvalid[c] = P.ok[c] ? P.ok[c]->as<uint8_t>() : nullptr;
    }
    uint64_t out_page_bytes = 0;
    for (int a = 0; a < L.n_out; ++a) {
        ResultColumn& rc = res->cols[a];
        rc.type = nd.output_attrs[a].type;
        const uint64_t pages = max_chunks * (L.out_width[a] == 4 ? 1 : 2);
        rc.pages = dev_alloc(pages * size_t(RJ_PAGE_SIZE), s);
        L.out_pages[a] = rc.pages->as<uint8_t>();
    }
    L.chunk_counter = counters->as<uint32_t>();
    L.abort_flag = counters->as<uint32_t>() + 1;
    L.row_counter = reinterpret_cast<unsigned long long*>(counters->as<uint8_t>() + 8);
    // a window of a streamed execute() ships its pages over PCIe right away: fewer partly filled ones
    L.min_chunks_per_warp = streamed_table != 0xffffffffu ? kEmitMinChunksPerWarp : kEmitMinChunksResident;
    RJ_CUDA(cudaMemsetAsync(pl.unit_cursor, 0, 4, s));
    uint32_t h[4] = {0, 0, 0, 0};
    {
        uint64_t in_bytes = (nb + np) * 4;
        for (int c = 0; c < L.n_bpay; ++c) in_bytes += nb * (L.bwidth[c] + (L.bvalid[c] ? 1 : 0));
        for (int c = 0; c < L.n_ppay; ++c) in_bytes += np * (L.pwidth[c] + (L.pvalid[c] ? 1 : 0));
        StageScope sc(ctx, RJ_ST_JOIN_EMIT, s, 1, in_bytes);
        launch_join_emit(L, np, ctx->sm_count, s);
    }
    RJ_CUDA(cudaMemcpyAsync(h, counters->p, 16, cudaMemcpyDeviceToHost, s));
    RJ_CUDA(cudaStreamSynchronize(s));
    if (h[1] != 0) return nullptr; // duplicate build keys: the general path handles them
    const uint64_t chunks = h[0];
    const uint64_t rows = static_cast<uint64_t>(h[2]) | (static_cast<uint64_t>(h[3]) << 32);
    if (chunks > max_chunks) throw EngineError("internal: the fused root join produced more chunks than planned");
    res->num_rows = rows;
    for (int a = 0; a < L.n_out; ++a) {
        ResultColumn& rc = res->cols[a];
        rc.n_pages = chunks * (L.out_width[a] == 4 ? 1 : 2);
        out_page_bytes += rc.n_pages * uint64_t(RJ_PAGE_SIZE);
        if (rc.n_pages == 0) rc.pages.reset();
    }
    if (ctx->profiling) ctx->stats[RJ_ST_JOIN_EMIT].bytes += out_page_bytes;
    return res;
}

std::unique_ptr<rj_result> execute_resident(rj_ctx* ctx, const rj_plan_t* plan, const rj_inputs* in,
                                            Exec::DecodedMap* shared_decoded = nullptr, uint32_t streamed_table = 0xffffffffu,
                                            double streamed_scale = 1.0) {
    if (plan->root >= plan->n_nodes) throw EngineError("plan root out of range");
    static const bool trace = getenv("RJ_TRACE") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    g_alloc_ns = g_free_ns = g_alloc_bytes = g_alloc_calls = g_alloc_misses = 0;
    std::unique_ptr<rj_result> res;
    double t_run = 0, t_root = 0;
    {
        Exec ex(ctx, plan, in);
        ex.shared_decoded = shared_decoded;
        ex.streamed_table = streamed_table;
        ex.streamed_scale = streamed_scale;
        if (plan->nodes[plan->root].is_join) res = ex.root_fused(plan->root); // null: not eligible / duplicate build keys
        if (!res) {
            Rel r = ex.run(plan->root);
            t_run = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            res = ex.root(plan->root, r);
        }
        RJ_CUDA(cudaStreamSynchronize(ctx->stream));
        if (const uint32_t bad = *static_cast<volatile uint32_t*>(ctx->err_host)) {
            *ctx->err_host = 0;
            if (bad & RJ_ERR_ORPHAN_LONG_PAGE) throw EngineError("long string page 0xfffe must follows a string"); // build_table.cpp:401-402
            throw EngineError("malformed input pages");
        }
        t_root = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    if (trace) {
        double t_all = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        fprintf(stderr, "[rj] execute: joins %.2f ms, +root/sync %.2f ms, +teardown %.2f ms | alloc %llu calls (%llu misses) %.1f MB %.2f ms\n",
                t_run, t_root, t_all, (unsigned long long)g_alloc_calls, (unsigned long long)g_alloc_misses, g_alloc_bytes / 1e6, g_alloc_ns / 1e6);
    }
    return res;
}

// columns the plan can reach (join keys + root outputs); only those are uploaded and decoded
void collect_wanted(const rj_plan_t* plan, std::set<std::pair<uint32_t, uint32_t>>* wanted) {
    struct Walker {
        const rj_plan_t* plan;
        std::pair<uint32_t, uint32_t> resolve(uint64_t n, uint64_t a) const {
            if (n >= plan->n_nodes) throw EngineError("plan node index out of range");
            const rj_node_t& nd = plan->nodes[n];
            if (a >= nd.n_output_attrs) throw EngineError("attribute index out of range");
            const uint64_t src = nd.output_attrs[a].index;
            if (!nd.is_join) return {static_cast<uint32_t>(nd.base_table_id), static_cast<uint32_t>(src)};
            if (nd.left >= plan->n_nodes || nd.right >= plan->n_nodes) throw EngineError("join child out of range");
            const uint64_t left_w = plan->nodes[nd.left].n_output_attrs;
            return src < left_w ? resolve(nd.left, src) : resolve(nd.right, src - left_w);
        }
        void walk(uint64_t n, std::set<std::pair<uint32_t, uint32_t>>* w) const {
            if (n >= plan->n_nodes) throw EngineError("plan node index out of range");
            const rj_node_t& nd = plan->nodes[n];
            if (!nd.is_join) return;
            w->insert(resolve(nd.left, nd.left_attr));
            w->insert(resolve(nd.right, nd.right_attr));
            walk(nd.left, w);
            walk(nd.right, w);
        }
    } wk{plan};
    if (plan->root >= plan->n_nodes) throw EngineError("plan root out of range");
    wk.walk(plan->root, wanted);
    const rj_node_t& root = plan->nodes[plan->root];
    for (uint32_t a = 0; a < root.n_output_attrs; ++a) wanted->insert(wk.resolve(plan->root, a));
    for (auto& tc: *wanted) {
        if (tc.first >= plan->n_inputs) throw EngineError("base table out of range");
        if (tc.second >= plan->inputs[tc.first].n_columns) throw EngineError("scan attribute out of range");
    }
}

template <class F>
int guarded(rj_ctx* ctx, F f) {
    if (!ctx) return 1;
    try {
        cudaSetDevice(ctx->device);
        t_home_stream = ctx->stream;
        t_cache = ctx->cache;
        f();
        return 0;
    } catch (const std::exception& e) {
        ctx->err = e.what();
        cudaGetLastError(); // clear a sticky launch error so the next call reports its own
        return 1;
    }
}

// Stage entry points: a caller that passes no stream gets the context's (non-blocking) stream.  Such a caller
// prepares its buffers on the default stream (torch does: a zero-fill may still be queued there), so the context's
// stream is first ordered behind everything the legacy default stream has been given.
cudaStream_t pick_stream(rj_ctx* ctx, void* stream) {
    if (stream) return static_cast<cudaStream_t>(stream);
    static thread_local cudaEvent_t ev[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    cudaEvent_t& e = ev[dev & 63];
    if (!e) RJ_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    RJ_CUDA(cudaEventRecord(e, cudaStreamLegacy));
    RJ_CUDA(cudaStreamWaitEvent(ctx->stream, e, 0));
    return ctx->stream;
}

} // namespace

// ================================================================================================
// C-ABI
// ================================================================================================
extern "C" {

const char* rj_version(void) { return "radix-join_b200 0.1 (sm_100a)"; }

int rj_ctx_create(int device, rj_ctx** out) {
    try {
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
            throw EngineError("no CUDA device: the B200 engine has no CPU fallback");
        }
        if (device < 0 || device >= n) throw EngineError("device index out of range");
        RJ_CUDA(cudaSetDevice(device));
        cudaDeviceProp prop;
        RJ_CUDA(cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10) {
            throw EngineError(std::string("device ") + prop.name + " is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                              "; this engine is built for sm_100a only");
        }
        auto ctx = std::make_unique<rj_ctx>();
        ctx->device = device;
        ctx->sm_count = prop.multiProcessorCount;
        RJ_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->cache = std::make_shared<BlockCache>();
        RJ_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&ctx->err_host), 64, cudaHostAllocMapped));
        std::memset(ctx->err_host, 0, 64);
        RJ_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&ctx->err_dev), ctx->err_host, 0));
        unsigned hc = std::thread::hardware_concurrency();
        ctx->host_threads = hc ? static_cast<int>(std::min(hc, 32u)) : 8;
        *out = ctx.release();
        return 0;
    } catch (const std::exception& e) {
        g_create_error = e.what();
        return 1;
    }
}

void rj_ctx_destroy(rj_ctx* ctx) {
    if (!ctx) return;
    for (rj_ctx* sub: ctx->subs) rj_ctx_destroy(sub);
    ctx->subs.clear();
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    profile_collect(ctx);
    for (auto e: ctx->free_events) cudaEventDestroy(e);
    ctx->pipe.reset(); // joins the workers, frees the pinned rings
    if (ctx->err_host) cudaFreeHost(ctx->err_host);
    if (ctx->h2d_stream) cudaStreamDestroy(ctx->h2d_stream);
    if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
    for (int i = 0; i < 2; ++i)
        if (ctx->up_ev[i]) cudaEventDestroy(ctx->up_ev[i]);
    cudaStreamDestroy(ctx->stream);
    ctx->cache->trim(); // release the cached HBM; blocks still held by live results follow when those are freed
    delete ctx;
}

const char* rj_last_error(const rj_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
int rj_ctx_device(const rj_ctx* ctx) { return ctx->device; }
void* rj_ctx_stream(const rj_ctx* ctx) { return ctx ? ctx->stream : nullptr; }
uint64_t rj_kernel_launch_count(void) { return g_kernel_launches.load(); }
int rj_ctx_sm_count(const rj_ctx* ctx) { return ctx->sm_count; }
int rj_ctx_set_host_threads(rj_ctx* ctx, int n) {
    if (!ctx || n < 1) return 1;
    ctx->host_threads = std::min(n, 256);
    return 0;
}

int rj_inputs_upload(rj_ctx* ctx, const rj_table_t* tables, uint32_t n_tables, rj_inputs** out) {
    return guarded(ctx, [&] {
        auto in = upload_tables(ctx, tables, n_tables, nullptr);
        RJ_CUDA(cudaStreamSynchronize(ctx->stream));
        *out = in.release();
    });
}

int rj_inputs_adopt_device(rj_ctx* ctx, const rj_table_t* tables, uint32_t n_tables, rj_inputs** out) {
    return guarded(ctx, [&] {
        auto in = std::make_unique<rj_inputs>();
        in->tables.resize(n_tables);
        std::vector<Buf> totals;
        for (uint32_t t = 0; t < n_tables; ++t) {
            TableDev& td = in->tables[t];
            td.num_rows = tables[t].num_rows;
            td.cols.resize(tables[t].n_columns);
            for (uint32_t c = 0; c < tables[t].n_columns; ++c) {
                const rj_column_t& col = tables[t].columns[c];
                ColumnDev& cd = td.cols[c];
                cd.type = col.type;
                cd.n_pages = col.n_pages;
                cd.pages = static_cast<const uint8_t*>(col.contiguous);
                if (col.n_pages && !cd.pages) throw EngineError("adopted column needs a contiguous device address");
                Buf tot = dev_alloc_zero(16, ctx->stream);
                Buf rows = dev_alloc(col.n_pages * 4, ctx->stream);
                launch_page_rows(cd.pages, cd.n_pages, cd.type, rows->as<uint32_t>(), tot->as<uint64_t>(), ctx->stream);
                uint64_t h[2] = {0, 0};
                RJ_CUDA(cudaMemcpyAsync(h, tot->p, 16, cudaMemcpyDeviceToHost, ctx->stream));
                RJ_CUDA(cudaStreamSynchronize(ctx->stream));
                cd.page_rows = h[0];
                cd.non_null = h[1];
                check_column_rows(td, cd);
            }
        }
        *out = in.release();
    });
}

int rj_inputs_adopt_dense(rj_ctx* ctx, const rj_dense_table_t* tables, uint32_t n_tables, rj_inputs** out) {
    return guarded(ctx, [&] {
        auto in = std::make_unique<rj_inputs>();
        in->tables.resize(n_tables);
        for (uint32_t t = 0; t < n_tables; ++t) {
            TableDev& td = in->tables[t];
            td.num_rows = tables[t].num_rows;
            td.cols.resize(tables[t].n_columns);
            for (uint32_t c = 0; c < tables[t].n_columns; ++c) {
                const rj_dense_column_t& col = tables[t].columns[c];
                if (col.type != RJ_INT32 && col.type != RJ_INT64 && col.type != RJ_FP64) throw EngineError("dense columns must be fixed-width");
                if (td.num_rows && !col.d_values) throw EngineError("dense column without values");
                ColumnDev& cd = td.cols[c];
                cd.type = col.type;
                cd.dense = true;
                cd.dense_values = col.d_values;
                cd.dense_valid = col.d_valid;
                cd.page_rows = td.num_rows;
                cd.non_null = col.d_valid ? 0 : td.num_rows;
            }
        }
        *out = in.release();
    });
}

void rj_inputs_free(rj_ctx* ctx, rj_inputs* in) {
    if (ctx) cudaSetDevice(ctx->device);
    delete in;
}

int rj_execute_resident(rj_ctx* ctx, const rj_plan_t* plan, const rj_inputs* in, rj_result** out) {
    return guarded(ctx, [&] {
        if (!plan || !in) throw EngineError("null plan or inputs");
        *out = execute_resident(ctx, plan, in).release();
    });
}

int rj_execute(rj_ctx* ctx, const rj_plan_t* plan, rj_result** out) {
    return guarded(ctx, [&] {
        if (!plan) throw EngineError("null plan");
        std::set<std::pair<uint32_t, uint32_t>> wanted;
        collect_wanted(plan, &wanted);
        auto in = upload_tables(ctx, plan->inputs, plan->n_inputs, &wanted);
        *out = execute_resident(ctx, plan, in.get()).release();
    });
}


// ---- streamed execution: host pages in, host pages out, PCIe busy in both directions ----------------
// An inner join distributes over a union of its inputs, so the plan can run once per row window of ONE
// base table (the largest, when a single scan reads it) with every other table resident:
//     gather + upload(k+1)  ||  kernels(k)  ||  download + scatter(k-1)
// The copies run on their own streams; individually allocated pages (the contest's ColumnarTable,
// plan.h:60-68) are gathered into / scattered out of pinned ring buffers by the worker pool
// (host_pipe.h), contiguous pinned buffers are plain DMAs.  Windows are cut on row boundaries, so the
// pages of a window's columns start and end at different rows: the decode shifts each column to the
// window (ColumnDev::skip_rows).
namespace {

struct StreamCol {
    uint32_t              col = 0;
    std::vector<uint64_t> row_prefix; // [n_pages + 1] rows before page p
    bool                  has_null = false;
    Buf                   slot[2];
    uint64_t              slot_pages = 0;
};

// Where result pages go.  deliver() issues the D2H copies of one window's result on `stream` (the
// result stays alive until they have run); finish() returns once every page is in host memory.
struct OutputSink {
    virtual ~OutputSink() = default;
    virtual void deliver(rj_ctx* ctx, rj_result* res, cudaStream_t stream) = 0;
    virtual void finish() = 0;
    virtual void abort() {}
};

// rj_execute_streamed: the caller hands out one contiguous host buffer per (window, column)
struct ContiguousSink: OutputSink {
    rj_page_sink_t sink;
    void*          user;
    ContiguousSink(rj_page_sink_t s, void* u): sink(s), user(u) {}
    void deliver(rj_ctx* ctx, rj_result* res, cudaStream_t stream) override {
        for (uint32_t c = 0; c < res->cols.size(); ++c) {
            const ResultColumn& rc = res->cols[c];
            if (rc.n_pages == 0) continue;
            void* dst = sink(user, c, rc.type, rc.n_pages);
            if (!dst) throw EngineError("rj_execute_streamed: the sink returned no buffer");
            StageScope scope(ctx, RJ_ST_D2H, stream, 1, rc.n_pages * uint64_t(RJ_PAGE_SIZE));
            RJ_CUDA(cudaMemcpyAsync(dst, rc.pages->p, rc.n_pages * size_t(RJ_PAGE_SIZE), cudaMemcpyDeviceToHost, stream));
        }
    }
    void finish() override {}
};

// rj_execute_pages: every result page is an individually allocated page of the caller (`new Page`,
// plan.h:64-68).  Pages are allocated on the calling thread (one allocator arena, so the memory of the
// previous result is reused), filled by the worker pool behind the D2H copies, and handed over column by
// column in window order at the end -- the same order for every column, which keeps rows aligned.
struct PageSink: OutputSink {
    const rj_page_alloc_t* alloc;
    HostPipe*              hp;
    struct Piece {
        uint32_t           column;
        int32_t            type;
        std::vector<void*> pages;
    };
    std::deque<Piece> pieces;
    TaskGroup         group;
    bool              open = true;
    std::atomic<bool> alloc_failed{false};
    // RJ_TRACE: where the host side of the download spends its time (ns, summed over threads)
    std::atomic<uint64_t> ns_alloc{0}, ns_copy{0}, ns_acquire{0};
    PageSink(const rj_page_alloc_t* a, HostPipe* h): alloc(a), hp(h) { group.open(); }
    void deliver(rj_ctx* ctx, rj_result* res, cudaStream_t stream) override {
        using clk = std::chrono::steady_clock;
        auto ns_since = [](clk::time_point t) { return static_cast<uint64_t>(std::chrono::duration_cast<std::chrono::nanoseconds>(clk::now() - t).count()); };
        for (uint32_t c = 0; c < res->cols.size(); ++c) {
            const ResultColumn& rc = res->cols[c];
            if (rc.n_pages == 0) continue;
            pieces.emplace_back();
            Piece& pc = pieces.back();
            pc.column = c;
            pc.type = rc.type;
            pc.pages.assign(rc.n_pages, nullptr);
            StageScope scope(ctx, RJ_ST_D2H, stream, 1, rc.n_pages * uint64_t(RJ_PAGE_SIZE));
            void**         dst = pc.pages.data(); // the deque never moves a Piece, the vector is never resized
            const uint8_t* src = rc.pages->as<uint8_t>();
            const uint64_t per = HostPipe::kBufBytes / RJ_PAGE_SIZE;
            for (uint64_t q = 0; q < rc.n_pages; q += per) {
                const uint64_t m = std::min<uint64_t>(per, rc.n_pages - q);
                const auto t0 = clk::now();
                PinnedBuf* buf = hp->down.acquire(); // blocks while all staging buffers are being scattered
                ns_acquire.fetch_add(ns_since(t0), std::memory_order_relaxed);
                cudaError_t e = cudaMemcpyAsync(buf->p, src + q * RJ_PAGE_SIZE, m * RJ_PAGE_SIZE, cudaMemcpyDeviceToHost, stream);
                if (e == cudaSuccess) e = cudaEventRecord(buf->ev, stream);
                if (e != cudaSuccess) {
                    hp->down.release(buf);
                    throw CudaError(std::string("page download: ") + cudaGetErrorString(e));
                }
                group.add();
                hp->waiter.after(buf->ev, [this, buf, dst, q, m, ns_since] {
                    // the pages of this piece are allocated here, by the worker that fills them: a serial
                    // `new Page` per result page on the calling thread was the longest stage of the pipeline
                    const auto t1 = clk::now();
                    const bool ok = !alloc_failed.load(std::memory_order_relaxed) && alloc->new_pages(alloc->user, m, dst + q) == 0;
                    const auto t2 = clk::now();
                    if (ok) {
                        for (uint64_t i = 0; i < m; ++i) copy_page(dst[q + i], buf->p + i * RJ_PAGE_SIZE);
                        copy_fence();
                    } else {
                        alloc_failed.store(true, std::memory_order_relaxed);
                    }
                    ns_alloc.fetch_add(static_cast<uint64_t>(std::chrono::duration_cast<std::chrono::nanoseconds>(t2 - t1).count()), std::memory_order_relaxed);
                    ns_copy.fetch_add(ns_since(t2), std::memory_order_relaxed);
                    hp->down.release(buf);
                    group.done();
                });
            }
        }
    }
    void drain() {
        if (open) {
            open = false;
            group.seal();
        }
        group.wait();
    }
    void finish() override {
        drain();
        if (getenv("RJ_TRACE") != nullptr)
            fprintf(stderr, "[rj pages] download side: staging-buffer waits %.1f ms (calling thread), page allocation %.1f ms, page copies %.1f ms (summed over workers)\n",
                    ns_acquire.load() / 1e6, ns_alloc.load() / 1e6, ns_copy.load() / 1e6);
        if (alloc_failed.load()) throw EngineError("rj_execute_pages: the page allocator failed");
        while (!pieces.empty()) {
            Piece& pc = pieces.front();
            if (alloc->append(alloc->user, pc.column, pc.type, pc.pages.data(), pc.pages.size()) != 0)
                throw EngineError("rj_execute_pages: the caller refused result pages");
            pieces.pop_front(); // handed over: no longer ours to free
        }
    }
    void abort() override {
        try { drain(); } catch (...) {}
        for (auto& pc: pieces) {
            if (!alloc->free_pages) continue;
            // pages are allocated batch by batch: give back whatever exists
            std::vector<void*> have;
            for (void* p: pc.pages)
                if (p) have.push_back(p);
            if (!have.empty()) alloc->free_pages(alloc->user, have.size(), have.data());
        }
        pieces.clear();
    }
};

struct PendingDownload {
    std::unique_ptr<rj_result> res;
    cudaEvent_t                done = nullptr;
};

uint64_t execute_streamed(rj_ctx* ctx, const rj_plan_t* plan, uint64_t chunk_bytes, OutputSink& out) {
    if (chunk_bytes == 0) {
        // RJ_WINDOW_BYTES: window size of callers that cannot pass one (Contest::execute); tests use it to
        // cut small inputs into many windows
        const char* env = getenv("RJ_WINDOW_BYTES");
        chunk_bytes = env && atoll(env) > 0 ? static_cast<uint64_t>(atoll(env)) : uint64_t(256) << 20;
    }
    ensure_copy_streams(ctx);
    ensure_pipe(ctx);
    std::set<std::pair<uint32_t, uint32_t>> wanted;
    collect_wanted(plan, &wanted);

    // the table to stream: read by exactly one scan, fixed-width columns only, at least two chunks
    std::vector<int> scans(plan->n_inputs, 0);
    for (uint64_t i = 0; i < plan->n_nodes; ++i)
        if (!plan->nodes[i].is_join && plan->nodes[i].base_table_id < plan->n_inputs) ++scans[plan->nodes[i].base_table_id];
    int      pick = -1;
    uint64_t pick_bytes = 0;
    for (uint32_t t = 0; t < plan->n_inputs; ++t) {
        if (scans[t] != 1) continue;
        uint64_t bytes = 0;
        bool     ok = true;
        for (uint32_t c = 0; c < plan->inputs[t].n_columns; ++c) {
            if (!wanted.count({t, c})) continue;
            if (plan->inputs[t].columns[c].type == RJ_VARCHAR) ok = false;
            bytes += plan->inputs[t].columns[c].n_pages * uint64_t(RJ_PAGE_SIZE);
        }
        if (ok && bytes >= 2 * chunk_bytes && bytes > pick_bytes) {
            pick = static_cast<int>(t);
            pick_bytes = bytes;
        }
    }
    if (pick < 0) {
        // nothing worth streaming: one upload, one execute, one download
        auto in  = upload_tables(ctx, plan->inputs, plan->n_inputs, &wanted);
        auto res = execute_resident(ctx, plan, in.get());
        out.deliver(ctx, res.get(), ctx->stream);
        RJ_CUDA(cudaStreamSynchronize(ctx->stream));
        out.finish();
        return res->num_rows;
    }

    const uint32_t    T  = static_cast<uint32_t>(pick);
    const rj_table_t& ht = plan->inputs[T];
    // (Every window joins against ALL tables of the other side -- 2^15 of them at config 2 -- so the fused join sizes
    // its grid by the tables and lets only as many warps emit as the window's probe tuples can feed: measured on
    // config 2, Contest::execute, 256 MiB windows 355 ms, 512 MiB 368 ms.)
    // page headers of the streamed table: rows before every page, per column
    std::vector<StreamCol> cols;
    for (uint32_t c = 0; c < ht.n_columns; ++c) {
        if (!wanted.count({T, c})) continue;
        const rj_column_t& hc = ht.columns[c];
        if (hc.n_pages && !hc.pages && !hc.contiguous) throw EngineError("column has pages but no page pointers");
        StreamCol sc;
        sc.col = c;
        sc.row_prefix.assign(hc.n_pages + 1, 0);
        std::atomic<uint64_t> non_null{0};
        pool_for(ctx, hc.n_pages, 4096, [&](uint64_t b, uint64_t e) {
            uint64_t nn = 0;
            for (uint64_t i = b; i < e; ++i) {
                uint64_t r = 0;
                page_counts(host_page(hc, i), hc.type, &r, &nn);
                sc.row_prefix[i + 1] = r;
            }
            non_null.fetch_add(nn, std::memory_order_relaxed);
        });
        for (uint64_t i = 0; i < hc.n_pages; ++i) sc.row_prefix[i + 1] += sc.row_prefix[i];
        if (sc.row_prefix[hc.n_pages] > ht.num_rows) throw EngineError("row_idx");
        sc.has_null = non_null.load() != ht.num_rows;
        cols.push_back(std::move(sc));
    }

    // row windows: multiples of 4096 rows, about chunk_bytes of pages each
    uint64_t rows_per_chunk = static_cast<uint64_t>(static_cast<double>(ht.num_rows) * chunk_bytes / pick_bytes);
    rows_per_chunk = std::max<uint64_t>(4096, (rows_per_chunk + 4095) & ~uint64_t(4095));
    const uint64_t n_chunks = (ht.num_rows + rows_per_chunk - 1) / rows_per_chunk;
    // pages of column `sc` that hold rows [r0, r1)
    auto page_range = [&](const StreamCol& sc, uint64_t r0, uint64_t r1, uint64_t* p0, uint64_t* p1) {
        const auto& pre = sc.row_prefix;
        const uint64_t np = pre.size() - 1;
        *p0 = std::upper_bound(pre.begin(), pre.end(), r0) - pre.begin();
        *p0 = *p0 ? *p0 - 1 : 0;                     // last page with pre[p] <= r0
        if (*p0 > np) *p0 = np;
        *p1 = std::lower_bound(pre.begin(), pre.end(), r1) - pre.begin(); // first p with pre[p] >= r1
        if (*p1 > np) *p1 = np;
        if (*p1 < *p0) *p1 = *p0;
    };
    for (auto& sc: cols) {
        for (uint64_t k = 0; k < n_chunks; ++k) {
            uint64_t p0, p1;
            page_range(sc, k * rows_per_chunk, std::min(ht.num_rows, (k + 1) * rows_per_chunk), &p0, &p1);
            sc.slot_pages = std::max(sc.slot_pages, p1 - p0);
        }
        for (int i = 0; i < 2; ++i) sc.slot[i] = dev_alloc(std::max<uint64_t>(1, sc.slot_pages) * RJ_PAGE_SIZE, ctx->stream);
    }
    RJ_CUDA(cudaStreamSynchronize(ctx->stream)); // the slots are in place

    // window k's pages: gathered / copied on the upload stream; up_group[k] completes when every copy
    // of the window has been issued, and its last task records up_ev[k & 1] behind them
    std::vector<std::unique_ptr<TaskGroup>> up_group(n_chunks);
    auto issue_upload = [&](uint64_t k) {
        const int      slot = static_cast<int>(k & 1);
        const uint64_t r0 = k * rows_per_chunk, r1 = std::min(ht.num_rows, r0 + rows_per_chunk);
        uint64_t       bytes = 0;
        for (auto& sc: cols) {
            uint64_t p0, p1;
            page_range(sc, r0, r1, &p0, &p1);
            bytes += (p1 - p0) * RJ_PAGE_SIZE;
        }
        StageScope scope(ctx, RJ_ST_H2D, ctx->h2d_stream, cols.size(), bytes);
        up_group[k] = std::make_unique<TaskGroup>();
        TaskGroup* g = up_group[k].get();
        cudaEvent_t  ev = ctx->up_ev[slot];
        cudaStream_t hs = ctx->h2d_stream;
        g->on_complete = [ev, hs] {
            if (cudaEventRecord(ev, hs) != cudaSuccess) throw CudaError("window upload: cudaEventRecord failed");
        };
        g->open();
        try {
            for (auto& sc: cols) {
                uint64_t p0, p1;
                page_range(sc, r0, r1, &p0, &p1);
                upload_pages_async(ctx, ht.columns[sc.col], p0, p1 - p0, sc.slot[slot]->as<uint8_t>(), ctx->h2d_stream, g, nullptr);
            }
        } catch (...) {
            g->seal();
            throw;
        }
        g->seal();
    };

    Exec::DecodedMap             shared;
    std::deque<PendingDownload>  pending;
    uint64_t                     total_rows = 0;
    auto reap = [&](bool all) {
        while (!pending.empty()) {
            if (all) RJ_CUDA(cudaEventSynchronize(pending.front().done));
            else if (cudaEventQuery(pending.front().done) != cudaSuccess) { cudaGetLastError(); break; }
            cudaEventDestroy(pending.front().done);
            pending.pop_front();
        }
    };
    static const bool trace = getenv("RJ_TRACE") != nullptr;
    auto now_ms = [t0 = std::chrono::steady_clock::now()] {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    };
    if (trace) fprintf(stderr, "[rj stream] table %u: %llu rows in %llu windows of %llu rows\n", T, (unsigned long long)ht.num_rows,
                       (unsigned long long)n_chunks, (unsigned long long)rows_per_chunk);
    std::unique_ptr<rj_inputs> in;
    try {
        // the first window's gather runs beside the upload of the resident tables
        issue_upload(0);
        std::set<std::pair<uint32_t, uint32_t>> others;
        for (auto& w: wanted)
            if (w.first != T) others.insert(w);
        in = upload_tables(ctx, plan->inputs, plan->n_inputs, &others);
        for (uint64_t k = 0; k < n_chunks; ++k) {
            const double t_a = now_ms();
            if (k + 1 < n_chunks) issue_upload(k + 1); // its slot was last read by chunk k-1, which has finished
            const double t_b = now_ms();
            const int      slot = static_cast<int>(k & 1);
            const uint64_t r0 = k * rows_per_chunk, r1 = std::min(ht.num_rows, r0 + rows_per_chunk);
            up_group[k]->wait(); // every copy of the window is on the upload stream, the event behind them
            const double t_w = now_ms() - t_b;
            RJ_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->up_ev[slot], 0));
            TableDev& td = in->tables[T];
            td.num_rows = r1 - r0;
            for (auto& sc: cols) {
                uint64_t p0, p1;
                page_range(sc, r0, r1, &p0, &p1);
                ColumnDev& cd = td.cols[sc.col];
                cd.type = ht.columns[sc.col].type;
                cd.pages = sc.slot[slot]->as<uint8_t>();
                cd.n_pages = p1 - p0;
                cd.page_rows = sc.row_prefix[p1] - sc.row_prefix[p0];
                cd.windowed = true;
                cd.skip_rows = std::min(r0, sc.row_prefix[p1]) - sc.row_prefix[p0];
                cd.window_nulls = sc.has_null;
                cd.non_null = 0;
            }
            const double scale = static_cast<double>(ht.num_rows) / static_cast<double>(std::max<uint64_t>(1, r1 - r0));
            auto res = execute_resident(ctx, plan, in.get(), &shared, T, scale); // returns with the stream idle
            const double t_c = now_ms();
            total_rows += res->num_rows;
            PendingDownload pd;
            RJ_CUDA(cudaEventCreateWithFlags(&pd.done, cudaEventDisableTiming));
            out.deliver(ctx, res.get(), ctx->d2h_stream);
            RJ_CUDA(cudaEventRecord(pd.done, ctx->d2h_stream));
            pd.res = std::move(res);
            pending.push_back(std::move(pd));
            reap(false);
            if (trace) fprintf(stderr, "[rj stream] window %llu: issue upload %.2f ms, wait for own upload %.2f ms, execute %.2f ms, issue download + reap %.2f ms, %zu downloads pending\n",
                               (unsigned long long)k, t_b - t_a, t_w, t_c - t_b - t_w, now_ms() - t_c, pending.size());
        }
        reap(true);
        out.finish();
    } catch (...) {
        for (auto& g: up_group)
            if (g) { try { g->wait(); } catch (...) {} }
        cudaStreamSynchronize(ctx->h2d_stream);
        cudaStreamSynchronize(ctx->d2h_stream);
        for (auto& pd: pending) cudaEventDestroy(pd.done);
        out.abort();
        throw;
    }
    RJ_CUDA(cudaStreamSynchronize(ctx->h2d_stream));
    return total_rows;
}

} // namespace

// ================================================================================================
// multi-device execution of ONE plan from ONE process (rj_ctx_create_multi)
// ================================================================================================
// The reference's execute() is a single call from a single thread (tests/read_sql.cpp:1308-1317); to let it
// use several GPUs the same algorithm the torchrun driver runs with one process per GPU
// (radix_join_b200/dist_join.py: distributed_join_fused, pull variant) is run here by one host thread per
// device: each device takes a 1/G slice of the rows of both tables, decodes it, histograms the keys over
// the job's radix digits; from all devices' histograms every device derives where the runs of each pass-1
// digit lie; scatter pass 1 stays LOCAL; the owner of a digit range then runs scatter pass 2 directly out of
// the other devices' pass-1 arrays (peer access, TMA bulk loads over NVLink) and the fused build / probe /
// page-output kernel on the partitions it owns.  The devices' result page lists are appended in device
// order for every column, which keeps the columns row-aligned (include/plan.h:102-105).
// Eligible: the plan is one key / foreign-key join of two scans on an INT32 key, the outputs are the key and
// at most two fixed-width columns per side, and the build side needs two scatter passes (> 512 Ki rows);
// anything else runs on the group's first device as before.
namespace {

std::unique_ptr<rj_result> join_partitioned_impl(rj_ctx* ctx, const rj_part_side_t* build, const rj_part_side_t* probe,
                                                 const uint32_t* d_hist_build, const uint32_t* d_hist_probe, int32_t local_bits,
                                                 int32_t local_pass1_bits, int32_t hash_bits, const rj_part_out_t* outs, uint32_t n_out);

struct FusedShape {
    uint32_t table[2] = {0, 0};     // [0] build side (the smaller table), [1] probe side
    uint32_t key_col[2] = {0, 0};
    std::vector<uint32_t> cols[2];  // carried columns per side
    std::vector<int32_t>  types[2];
    rj_part_out_t outs[kEmitMaxOut];
    uint32_t n_out = 0;
};

bool analyze_fused_root(const rj_plan_t* plan, FusedShape* fs) {
    if (!plan || plan->root >= plan->n_nodes) return false;
    const rj_node_t& nd = plan->nodes[plan->root];
    if (!nd.is_join || nd.left >= plan->n_nodes || nd.right >= plan->n_nodes) return false;
    const rj_node_t &ln = plan->nodes[nd.left], &rn = plan->nodes[nd.right];
    if (ln.is_join || rn.is_join || nd.left == nd.right) return false;
    if (nd.n_output_attrs < 1 || nd.n_output_attrs > static_cast<uint32_t>(kEmitMaxOut)) return false;
    if (nd.left_attr >= ln.n_output_attrs || nd.right_attr >= rn.n_output_attrs) return false;
    if (ln.base_table_id >= plan->n_inputs || rn.base_table_id >= plan->n_inputs || ln.base_table_id == rn.base_table_id) return false;
    const rj_table_t &tl = plan->inputs[ln.base_table_id], &tr = plan->inputs[rn.base_table_id];
    const uint64_t lkey = ln.output_attrs[nd.left_attr].index, rkey = rn.output_attrs[nd.right_attr].index;
    if (lkey >= tl.n_columns || rkey >= tr.n_columns) return false;
    // the key type comes from the build side's declaration (src/execute.cpp:271-273); both physical columns must be INT32
    const rj_node_t& bnode = nd.build_left ? ln : rn;
    if (bnode.output_attrs[nd.build_left ? nd.left_attr : nd.right_attr].type != RJ_INT32) return false;
    if (tl.columns[lkey].type != RJ_INT32 || tr.columns[rkey].type != RJ_INT32) return false;
    if (tl.num_rows == 0 || tr.num_rows == 0 || tl.num_rows >= 0xffffffffull || tr.num_rows >= 0xffffffffull) return false;
    const bool table_left = tl.num_rows <= tr.num_rows;
    fs->table[0] = static_cast<uint32_t>(table_left ? ln.base_table_id : rn.base_table_id);
    fs->table[1] = static_cast<uint32_t>(table_left ? rn.base_table_id : ln.base_table_id);
    fs->key_col[0] = static_cast<uint32_t>(table_left ? lkey : rkey);
    fs->key_col[1] = static_cast<uint32_t>(table_left ? rkey : lkey);
    fs->n_out = nd.n_output_attrs;
    for (uint32_t a = 0; a < nd.n_output_attrs; ++a) {
        const uint64_t src = nd.output_attrs[a].index;
        if (src >= uint64_t(ln.n_output_attrs) + rn.n_output_attrs) return false;
        const bool     from_left = src < ln.n_output_attrs;
        const rj_node_t& sn = from_left ? ln : rn;
        const uint64_t col = sn.output_attrs[from_left ? src : src - ln.n_output_attrs].index;
        const rj_table_t& tb = from_left ? tl : tr;
        if (col >= tb.n_columns) return false;
        const int32_t t = tb.columns[col].type;
        if (t == RJ_VARCHAR || t != nd.output_attrs[a].type) return false;
        const int side = (from_left == table_left) ? 0 : 1;
        if (col == fs->key_col[side]) {
            fs->outs[a] = rj_part_out_t{side, -1};
            continue;
        }
        auto& list = fs->cols[side];
        size_t i = 0;
        while (i < list.size() && list[i] != col) ++i;
        if (i == list.size()) {
            if (list.size() == static_cast<size_t>(kEmitMaxPay)) return false;
            list.push_back(static_cast<uint32_t>(col));
            fs->types[side].push_back(t);
        }
        fs->outs[a] = rj_part_out_t{side, static_cast<int32_t>(i)};
    }
    return true;
}

// a reusable barrier of the device threads; a thread that failed keeps arriving so that nobody waits for ever
class HostBarrier {
public:
    explicit HostBarrier(int n): n_(n) {}
    void wait() {
        std::unique_lock<std::mutex> lk(mu_);
        const uint64_t gen = gen_;
        if (++count_ == n_) {
            count_ = 0;
            ++gen_;
            cv_.notify_all();
        } else {
            cv_.wait(lk, [&] { return gen_ != gen; });
        }
    }
private:
    std::mutex mu_;
    std::condition_variable cv_;
    int n_, count_ = 0;
    uint64_t gen_ = 0;
};

rj_ctx* group_ctx(rj_ctx* ctx, int d) { return d == 0 ? ctx : ctx->subs[d - 1]; }

struct MultiJob {
    int G = 0, g = 0, bits = 0, p1 = 0;
    const rj_plan_t* plan = nullptr;
    FusedShape fs;
    std::vector<StreamCol> cols[2];         // per side: key column first, then the carried columns (row prefixes)
    uint64_t rows_per_dev[2] = {0, 0};
    std::vector<uint32_t> H;                // [G][2][2^bits]
    struct Arrays { uint64_t a[5] = {0, 0, 0, 0, 0}; }; // keys, value 0, value 1, flag 0, flag 1 (device addresses)
    std::vector<Arrays> staging;            // [G * 2]
    std::vector<std::unique_ptr<rj_result>> results;
    std::vector<uint8_t> failed_dup;
    std::mutex mu;
    std::exception_ptr err;
    std::atomic<bool> failed{false};
    void fail(std::exception_ptr e) {
        std::lock_guard<std::mutex> lk(mu);
        if (!err) err = e;
        failed.store(true);
    }
};

// pages of column `sc` that hold rows [r0, r1)
void page_range_of(const StreamCol& sc, uint64_t r0, uint64_t r1, uint64_t* p0, uint64_t* p1) {
    const auto& pre = sc.row_prefix;
    const uint64_t np = pre.size() - 1;
    *p0 = std::upper_bound(pre.begin(), pre.end(), r0) - pre.begin();
    *p0 = *p0 ? *p0 - 1 : 0;
    if (*p0 > np) *p0 = np;
    *p1 = std::lower_bound(pre.begin(), pre.end(), r1) - pre.begin();
    if (*p1 > np) *p1 = np;
    if (*p1 < *p0) *p1 = *p0;
}

void multi_device_thread(rj_ctx* ctx, MultiJob* job, HostBarrier* bar, int me, OutputSink* sink) {
    cudaSetDevice(ctx->device);
    t_home_stream = ctx->stream;
    t_cache = ctx->cache;
    cudaStream_t s = ctx->stream;
    const int G = job->G, bits = job->bits, p1 = job->p1, g = job->g;
    const uint32_t nfin = 1u << bits, ndig = 1u << p1, per = ndig >> g, fin_per_dig = nfin / ndig;
    auto guard = [&](auto fn) {
        if (job->failed.load()) return;
        try {
            fn();
        } catch (...) {
            job->fail(std::current_exception());
        }
    };
    rj_inputs in;
    std::unique_ptr<Exec> ex;
    const DecodedCol* key[2] = {nullptr, nullptr};
    std::vector<const DecodedCol*> carried[2];
    uint64_t n_local[2] = {0, 0};
    Buf stage_buf[2][5], cursor_buf[2];
    // ---- phase 1: upload this device's row windows, decode, histogram ---------------------------------------
    guard([&] {
        in.tables.resize(job->plan->n_inputs);
        TaskGroup up;
        up.open();
        std::vector<Buf> slots;
        for (int side = 0; side < 2; ++side) {
            const uint32_t    T = job->fs.table[side];
            const rj_table_t& ht = job->plan->inputs[T];
            const uint64_t r0 = std::min<uint64_t>(ht.num_rows, uint64_t(me) * job->rows_per_dev[side]);
            const uint64_t r1 = std::min<uint64_t>(ht.num_rows, r0 + job->rows_per_dev[side]);
            n_local[side] = r1 - r0;
            TableDev& td = in.tables[T];
            td.num_rows = r1 - r0;
            td.cols.resize(ht.n_columns);
            for (uint32_t c = 0; c < ht.n_columns; ++c) td.cols[c].type = ht.columns[c].type;
            for (auto& sc: job->cols[side]) {
                uint64_t pa, pb;
                page_range_of(sc, r0, r1, &pa, &pb);
                ColumnDev& cd = td.cols[sc.col];
                cd.n_pages = r1 > r0 ? pb - pa : 0;
                cd.owned = dev_alloc(std::max<uint64_t>(1, cd.n_pages) * RJ_PAGE_SIZE, s);
                cd.pages = cd.owned->as<uint8_t>();
                cd.page_rows = cd.n_pages ? sc.row_prefix[pb] - sc.row_prefix[pa] : 0;
                cd.windowed = true;
                cd.skip_rows = cd.n_pages ? std::min(r0, sc.row_prefix[pb]) - sc.row_prefix[pa] : 0;
                cd.window_nulls = sc.has_null;
                if (cd.n_pages) upload_pages_async(ctx, ht.columns[sc.col], pa, cd.n_pages, cd.owned->as<uint8_t>(), s, &up, nullptr);
            }
        }
        up.seal();
        up.wait(); // every copy is on the stream
        ex = std::make_unique<Exec>(ctx, job->plan, &in);
        Buf hist = dev_alloc_zero(size_t(2) * nfin * 4, s);
        for (int side = 0; side < 2; ++side) {
            const uint32_t T = job->fs.table[side];
            key[side] = &ex->column(T, job->fs.key_col[side]);
            for (uint32_t c: job->fs.cols[side]) carried[side].push_back(&ex->column(T, c));
            launch_radix_histogram(key[side]->values->p, key[side]->valid_ptr(), n_local[side], 4, 0, bits, hist->as<uint32_t>() + side * nfin, ctx->sm_count, s);
        }
        RJ_CUDA(cudaMemcpyAsync(job->H.data() + size_t(me) * 2 * nfin, hist->p, size_t(2) * nfin * 4, cudaMemcpyDeviceToHost, s));
        // the arrays pass 1 writes and the digits' owners read
        for (int side = 0; side < 2; ++side) {
            const uint64_t n = n_local[side];
            MultiJob::Arrays& ar = job->staging[size_t(me) * 2 + side];
            stage_buf[side][0] = dev_alloc(n * 4 + 64, s);
            ar.a[0] = reinterpret_cast<uint64_t>(stage_buf[side][0]->p);
            int nf = 0;
            for (size_t c = 0; c < carried[side].size(); ++c) {
                const int w = job->fs.types[side][c] == RJ_INT32 ? 4 : 8;
                stage_buf[side][1 + c] = dev_alloc(n * w + 64, s);
                ar.a[1 + c] = reinterpret_cast<uint64_t>(stage_buf[side][1 + c]->p);
                if (job->cols[side][1 + c].has_null) {
                    stage_buf[side][3 + nf] = dev_alloc(n + 64, s);
                    ar.a[3 + nf] = reinterpret_cast<uint64_t>(stage_buf[side][3 + nf]->p);
                    ++nf;
                }
            }
        }
        RJ_CUDA(cudaStreamSynchronize(s));
    });
    bar->wait(); // A: every device's histogram and array addresses are known
    // ---- phase 2 + 3: layout from the histograms, scatter pass 1 into this device's own arrays ------------------
    std::vector<uint64_t> cnt(size_t(G) * 2 * ndig, 0); // [q][side][digit]
    guard([&] {
        for (int q = 0; q < G; ++q)
            for (int side = 0; side < 2; ++side) {
                const uint32_t* h = job->H.data() + (size_t(q) * 2 + side) * nfin;
                uint64_t* c = cnt.data() + (size_t(q) * 2 + side) * ndig;
                for (uint32_t d = 0; d < ndig; ++d) {
                    uint64_t sum = 0;
                    for (uint32_t f = 0; f < fin_per_dig; ++f) sum += h[d * fin_per_dig + f];
                    c[d] = sum;
                }
            }
        for (int side = 0; side < 2; ++side) {
            std::vector<uint32_t> cur(ndig);
            uint64_t run = 0;
            const uint64_t* c = cnt.data() + (size_t(me) * 2 + side) * ndig;
            for (uint32_t d = 0; d < ndig; ++d) {
                cur[d] = static_cast<uint32_t>(run);
                run += c[d];
            }
            cursor_buf[side] = dev_alloc(ndig * 4, s);
            RJ_CUDA(cudaMemcpyAsync(cursor_buf[side]->p, cur.data(), ndig * 4, cudaMemcpyHostToDevice, s));
            RJ_CUDA(cudaStreamSynchronize(s)); // `cur` dies with this scope
            CarryScatter c1;
            c1.keys = key[side]->values->as<uint32_t>();
            c1.valid = key[side]->valid_ptr();
            c1.n = n_local[side];
            c1.shift = bits - p1; c1.bits = p1; c1.cursor = cursor_buf[side]->as<uint32_t>();
            c1.keys_out = stage_buf[side][0]->as<uint32_t>();
            for (size_t k = 0; k < carried[side].size(); ++k) {
                c1.val_src[c1.n_val] = carried[side][k]->values->p;
                c1.val_dst[c1.n_val] = stage_buf[side][1 + k]->p;
                c1.val_width[c1.n_val++] = job->fs.types[side][k] == RJ_INT32 ? 4 : 8;
                if (job->cols[side][1 + k].has_null) {
                    if (!carried[side][k]->valid) throw EngineError("internal: a nullable column decoded without a validity bitmap");
                    c1.flag_src[c1.n_flag] = carried[side][k]->valid_ptr();
                    c1.flag_dst[c1.n_flag] = stage_buf[side][3 + c1.n_flag]->as<uint8_t>();
                    ++c1.n_flag;
                }
            }
            StageScope sc(ctx, RJ_ST_SCATTER, s, 1, n_local[side] * 8);
            launch_scatter_carry(c1, ctx->sm_count, s);
        }
        RJ_CUDA(cudaStreamSynchronize(s));
    });
    bar->wait(); // B: every device's first pass is complete
    // ---- phase 4: second pass out of the senders' arrays, join, pages ------------------------------------------
    guard([&] {
        rj_part_side_t sides[2];
        Buf tabs[2][4], lhist[2];
        for (int side = 0; side < 2; ++side) {
            const uint32_t n_sub = per * G;
            std::vector<uint64_t> table(size_t(n_sub) * 5, 0);
            std::vector<uint32_t> start(n_sub + 1, 0), tile(n_sub + 1, 0), group(n_sub, 0);
            const int widths[5] = {4, carried[side].size() > 0 ? (job->fs.types[side][0] == RJ_INT32 ? 4 : 8) : 0,
                                   carried[side].size() > 1 ? (job->fs.types[side][1] == RJ_INT32 ? 4 : 8) : 0, 1, 1};
            // sender-local start of every digit run
            std::vector<uint64_t> run_start(size_t(G) * ndig);
            for (int q = 0; q < G; ++q) {
                uint64_t run = 0;
                for (uint32_t d = 0; d < ndig; ++d) {
                    run_start[size_t(q) * ndig + d] = run;
                    run += cnt[(size_t(q) * 2 + side) * ndig + d];
                }
            }
            uint64_t pos = 0, tiles = 0;
            for (uint32_t j = 0; j < per; ++j) {
                const uint32_t d = static_cast<uint32_t>(me) * per + j;
                for (int q = 0; q < G; ++q) {
                    const uint32_t x = j * G + q;
                    const uint64_t c = cnt[(size_t(q) * 2 + side) * ndig + d];
                    start[x] = static_cast<uint32_t>(pos);
                    tile[x] = static_cast<uint32_t>(tiles);
                    group[x] = j;
                    const int64_t delta = static_cast<int64_t>(run_start[size_t(q) * ndig + d]) - static_cast<int64_t>(pos);
                    const MultiJob::Arrays& ar = job->staging[size_t(q) * 2 + side];
                    for (int a = 0; a < 5; ++a) table[size_t(x) * 5 + a] = ar.a[a] + static_cast<uint64_t>(delta * widths[a]);
                    pos += c;
                    tiles += (c + scatter_tile(4) - 1) / scatter_tile(4);
                }
            }
            start[n_sub] = static_cast<uint32_t>(pos);
            tile[n_sub] = static_cast<uint32_t>(tiles);
            // tuples per final partition of the range this device owns
            std::vector<uint32_t> lh(nfin >> g, 0);
            for (int q = 0; q < G; ++q) {
                const uint32_t* h = job->H.data() + (size_t(q) * 2 + side) * nfin + size_t(me) * (nfin >> g);
                for (uint32_t f = 0; f < (nfin >> g); ++f) lh[f] += h[f];
            }
            tabs[side][0] = dev_alloc(table.size() * 8, s);
            tabs[side][1] = dev_alloc(start.size() * 4, s);
            tabs[side][2] = dev_alloc(tile.size() * 4, s);
            tabs[side][3] = dev_alloc(group.size() * 4, s);
            lhist[side] = dev_alloc(lh.size() * 4, s);
            RJ_CUDA(cudaMemcpyAsync(tabs[side][0]->p, table.data(), table.size() * 8, cudaMemcpyHostToDevice, s));
            RJ_CUDA(cudaMemcpyAsync(tabs[side][1]->p, start.data(), start.size() * 4, cudaMemcpyHostToDevice, s));
            RJ_CUDA(cudaMemcpyAsync(tabs[side][2]->p, tile.data(), tile.size() * 4, cudaMemcpyHostToDevice, s));
            RJ_CUDA(cudaMemcpyAsync(tabs[side][3]->p, group.data(), group.size() * 4, cudaMemcpyHostToDevice, s));
            RJ_CUDA(cudaMemcpyAsync(lhist[side]->p, lh.data(), lh.size() * 4, cudaMemcpyHostToDevice, s));
            RJ_CUDA(cudaStreamSynchronize(s)); // the host vectors die with this scope
            rj_part_side_t& sd = sides[side];
            sd = rj_part_side_t{};
            sd.n = pos;
            sd.n_cols = static_cast<uint32_t>(carried[side].size());
            for (size_t k = 0; k < carried[side].size(); ++k) {
                sd.types[k] = job->fs.types[side][k];
                sd.d_valid_bytes[k] = job->cols[side][1 + k].has_null ? reinterpret_cast<const uint8_t*>(1) : nullptr; // "has NULLs": the bytes come through the table
            }
            sd.n_sub = n_sub;
            sd.d_src_table = tabs[side][0]->as<uint64_t>();
            sd.d_sub_start = tabs[side][1]->as<uint32_t>();
            sd.d_sub_tile = tabs[side][2]->as<uint32_t>();
            sd.d_sub_group = tabs[side][3]->as<uint32_t>();
        }
        auto res = join_partitioned_impl(ctx, &sides[0], &sides[1], lhist[0]->as<uint32_t>(), lhist[1]->as<uint32_t>(), bits - g, p1 - g, bits,
                                         job->fs.outs, job->fs.n_out);
        if (!res) job->failed_dup[me] = 1;
        job->results[me] = std::move(res);
    });
    bar->wait(); // C: nobody reads anybody's arrays any more; the verdict on duplicate keys is in
    bool dup = false;
    for (int q = 0; q < G; ++q) dup = dup || job->failed_dup[q];
    if (!dup) {
        guard([&] {
            // the root's declared column types (they equal the physical ones: analyze_fused_root checked)
            ensure_copy_streams(ctx);
            sink->deliver(ctx, job->results[me].get(), ctx->d2h_stream);
            RJ_CUDA(cudaStreamSynchronize(ctx->d2h_stream));
        });
    }
    ex.reset();
    t_cache.reset();
}

// -> true when the plan ran on the whole group (rows in *total); false: not eligible, run it on one device
bool execute_multi(rj_ctx* ctx, const rj_plan_t* plan, const std::vector<OutputSink*>& sinks, uint64_t* total) {
    const int G = static_cast<int>(ctx->subs.size()) + 1;
    if (G < 2 || (G & (G - 1)) != 0 || getenv("RJ_NO_MULTI") != nullptr) return false;
    MultiJob job;
    if (!analyze_fused_root(plan, &job.fs)) return false;
    int g = 0;
    while ((1 << g) < G) ++g;
    const uint64_t nb = plan->inputs[job.fs.table[0]].num_rows;
    int bits = 0;
    while ((nb >> bits) > kJoinTargetFill && bits < kMaxTotalBits) ++bits;
    if (bits <= kMaxPassBits) return false; // one scatter pass: a small join, one device is enough
    const int p1 = pass1_bits_of(bits);
    if (p1 <= g) return false; // the owners need a second pass of their own
    job.G = G; job.g = g; job.bits = bits; job.p1 = p1; job.plan = plan;
    // page headers of the columns in play: rows before every page
    for (int side = 0; side < 2; ++side) {
        const rj_table_t& ht = plan->inputs[job.fs.table[side]];
        std::vector<uint32_t> want = {job.fs.key_col[side]};
        for (uint32_t c: job.fs.cols[side]) want.push_back(c);
        for (uint32_t c: want) {
            const rj_column_t& hc = ht.columns[c];
            if (hc.n_pages && !hc.pages && !hc.contiguous) throw EngineError("column has pages but no page pointers");
            StreamCol sc;
            sc.col = c;
            sc.row_prefix.assign(hc.n_pages + 1, 0);
            std::atomic<uint64_t> non_null{0};
            pool_for(ctx, hc.n_pages, 4096, [&](uint64_t b, uint64_t e) {
                uint64_t nn = 0;
                for (uint64_t i = b; i < e; ++i) {
                    uint64_t r = 0;
                    page_counts(host_page(hc, i), hc.type, &r, &nn);
                    sc.row_prefix[i + 1] = r;
                }
                non_null.fetch_add(nn, std::memory_order_relaxed);
            });
            for (uint64_t i = 0; i < hc.n_pages; ++i) sc.row_prefix[i + 1] += sc.row_prefix[i];
            if (sc.row_prefix[hc.n_pages] > ht.num_rows) throw EngineError("row_idx");
            sc.has_null = non_null.load() != ht.num_rows;
            job.cols[side].push_back(std::move(sc));
        }
        const uint64_t per_dev = (ht.num_rows + G - 1) / G;
        job.rows_per_dev[side] = (per_dev + 4095) & ~uint64_t(4095);
    }
    job.H.assign(size_t(G) * 2 * (size_t(1) << bits), 0);
    job.staging.resize(size_t(G) * 2);
    job.results.resize(G);
    job.failed_dup.assign(G, 0);
    HostBarrier bar(G);
    std::vector<std::thread> threads;
    for (int d = 1; d < G; ++d) threads.emplace_back(multi_device_thread, group_ctx(ctx, d), &job, &bar, d, sinks[d]);
    multi_device_thread(ctx, &job, &bar, 0, sinks[0]);
    for (auto& t: threads) t.join();
    // this thread's bindings were reset by its own device pass: restore them for the caller
    cudaSetDevice(ctx->device);
    t_home_stream = ctx->stream;
    t_cache = ctx->cache;
    if (job.err) std::rethrow_exception(job.err);
    for (int q = 0; q < G; ++q)
        if (job.failed_dup[q]) return false; // duplicate build keys: nothing was delivered, the general path runs on one device
    uint64_t rows = 0;
    for (auto& r: job.results) rows += r ? r->num_rows : 0;
    *total = rows;
    if (getenv("RJ_TRACE") != nullptr) {
        fprintf(stderr, "[rj multi] %d devices: %d radix bits (%d in pass 1), build %llu + probe %llu rows in slices of %llu / %llu, %llu result rows:",
                G, bits, p1, (unsigned long long)plan->inputs[job.fs.table[0]].num_rows, (unsigned long long)plan->inputs[job.fs.table[1]].num_rows,
                (unsigned long long)job.rows_per_dev[0], (unsigned long long)job.rows_per_dev[1], (unsigned long long)rows);
        for (auto& r: job.results) fprintf(stderr, " %llu", (unsigned long long)(r ? r->num_rows : 0));
        fprintf(stderr, "\n");
    }
    return true;
}

} // namespace

int rj_execute_streamed(rj_ctx* ctx, const rj_plan_t* plan, uint64_t chunk_bytes, rj_page_sink_t sink, void* user,
                        uint64_t* num_rows) {
    return guarded(ctx, [&] {
        if (!plan) throw EngineError("null plan");
        if (!sink) throw EngineError("rj_execute_streamed: no page sink");
        ContiguousSink out(sink, user);
        const uint64_t n = execute_streamed(ctx, plan, chunk_bytes, out);
        if (num_rows) *num_rows = n;
    });
}

int rj_execute_pages(rj_ctx* ctx, const rj_plan_t* plan, uint64_t chunk_bytes, const rj_page_alloc_t* alloc, uint64_t* num_rows) {
    return guarded(ctx, [&] {
        if (!plan) throw EngineError("null plan");
        if (!alloc || !alloc->new_pages || !alloc->append) throw EngineError("rj_execute_pages: no page allocator");
        if (!ctx->subs.empty()) {
            // a device group: one sink per device, handed over in device order (each device's share of the result
            // holds the same rows in every column)
            const int G = static_cast<int>(ctx->subs.size()) + 1;
            std::vector<std::unique_ptr<PageSink>> sinks;
            std::vector<OutputSink*> raw;
            for (int d = 0; d < G; ++d) {
                sinks.push_back(std::make_unique<PageSink>(alloc, ensure_pipe(group_ctx(ctx, d))));
                raw.push_back(sinks.back().get());
            }
            uint64_t n = 0;
            bool done = false;
            try {
                done = execute_multi(ctx, plan, raw, &n);
                if (done)
                    for (auto& sk: sinks) sk->finish();
            } catch (...) {
                for (auto& sk: sinks) sk->abort();
                throw;
            }
            if (done) {
                if (num_rows) *num_rows = n;
                return;
            }
            for (auto& sk: sinks) sk->abort(); // nothing was delivered
        }
        PageSink out(alloc, ensure_pipe(ctx));
        uint64_t n = 0;
        try {
            n = execute_streamed(ctx, plan, chunk_bytes, out);
        } catch (...) {
            out.abort();
            throw;
        }
        if (num_rows) *num_rows = n;
    });
}

int rj_ctx_create_multi(const int* devices, uint32_t n_devices, rj_ctx** out) {
    if (!devices || n_devices == 0 || !out) {
        g_create_error = "rj_ctx_create_multi: no devices";
        return 1;
    }
    rj_ctx* first = nullptr;
    if (rj_ctx_create(devices[0], &first) != 0) return 1;
    try {
        for (uint32_t i = 1; i < n_devices; ++i) {
            rj_ctx* sub = nullptr;
            if (rj_ctx_create(devices[i], &sub) != 0) throw EngineError(g_create_error);
            first->subs.push_back(sub);
        }
        // every device of the group reads every other one's memory (scatter pass 2 pulls its regions over NVLink)
        for (uint32_t i = 0; i < n_devices; ++i) {
            RJ_CUDA(cudaSetDevice(devices[i]));
            for (uint32_t j = 0; j < n_devices; ++j) {
                if (i == j) continue;
                int can = 0;
                RJ_CUDA(cudaDeviceCanAccessPeer(&can, devices[i], devices[j]));
                if (!can) throw EngineError("rj_ctx_create_multi: the devices cannot access each other's memory");
                const cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) RJ_CUDA(e);
                cudaGetLastError();
            }
        }
        // the host cores are shared by the devices' worker pools
        const int per = std::max(2, first->host_threads / static_cast<int>(n_devices));
        first->host_threads = per;
        for (rj_ctx* sub: first->subs) sub->host_threads = per;
        RJ_CUDA(cudaSetDevice(devices[0]));
        *out = first;
        return 0;
    } catch (const std::exception& e) {
        g_create_error = e.what();
        for (rj_ctx* sub: first->subs) rj_ctx_destroy(sub);
        first->subs.clear();
        rj_ctx_destroy(first);
        return 1;
    }
}

int rj_ctx_group_size(const rj_ctx* ctx) { return ctx ? static_cast<int>(ctx->subs.size()) + 1 : 0; }

uint64_t rj_result_num_rows(const rj_result* r) { return r->num_rows; }
uint32_t rj_result_num_columns(const rj_result* r) { return static_cast<uint32_t>(r->cols.size()); }
int32_t  rj_result_column_type(const rj_result* r, uint32_t c) { return r->cols[c].type; }
uint64_t rj_result_column_pages(const rj_result* r, uint32_t c) { return r->cols[c].n_pages; }
uint64_t rj_result_column_device_ptr(const rj_result* r, uint32_t c) {
    return r->cols[c].pages ? reinterpret_cast<uint64_t>(r->cols[c].pages->p) : 0;
}

int rj_result_fetch(rj_ctx* ctx, const rj_result* r, uint32_t col, void* const* dst_pages, void* dst_contiguous) {
    return guarded(ctx, [&] {
        if (col >= r->cols.size()) throw EngineError("result column out of range");
        const ResultColumn& rc = r->cols[col];
        if (rc.n_pages == 0) return;
        if (!dst_pages && !dst_contiguous) throw EngineError("no destination pages");
        StageScope scope(ctx, RJ_ST_D2H, ctx->stream, 1, rc.n_pages * uint64_t(RJ_PAGE_SIZE));
        bool pageable_contiguous = false;
        if (!dst_pages && rc.n_pages * size_t(RJ_PAGE_SIZE) >= (size_t(4) << 20)) {
            // a large pageable destination (numpy): through the pinned ring like individually allocated pages
            cudaPointerAttributes attr{};
            if (cudaPointerGetAttributes(&attr, dst_contiguous) != cudaSuccess) cudaGetLastError();
            pageable_contiguous = attr.type == cudaMemoryTypeUnregistered;
        }
        if (!dst_pages && !pageable_contiguous) {
            // contiguous destination: one DMA into the caller's buffer
            RJ_CUDA(cudaMemcpyAsync(dst_contiguous, rc.pages->p, rc.n_pages * size_t(RJ_PAGE_SIZE), cudaMemcpyDeviceToHost, ctx->stream));
            RJ_CUDA(cudaStreamSynchronize(ctx->stream));
            return;
        }
        uint8_t* const flat = dst_pages ? nullptr : static_cast<uint8_t*>(dst_contiguous);
        // individually allocated destination pages: D2H into pinned ring buffers, scattered by the pool
        HostPipe*      hp = ensure_pipe(ctx);
        TaskGroup      group;
        const uint8_t* src = rc.pages->as<uint8_t>();
        const uint64_t per = HostPipe::kBufBytes / RJ_PAGE_SIZE;
        group.open();
        try {
            for (uint64_t q = 0; q < rc.n_pages; q += per) {
                const uint64_t m = std::min<uint64_t>(per, rc.n_pages - q);
                PinnedBuf* buf = hp->down.acquire();
                cudaError_t e = cudaMemcpyAsync(buf->p, src + q * RJ_PAGE_SIZE, m * RJ_PAGE_SIZE, cudaMemcpyDeviceToHost, ctx->stream);
                if (e == cudaSuccess) e = cudaEventRecord(buf->ev, ctx->stream);
                if (e != cudaSuccess) {
                    hp->down.release(buf);
                    throw CudaError(std::string("page download: ") + cudaGetErrorString(e));
                }
                group.add();
                TaskGroup* g = &group;
                hp->waiter.after(buf->ev, [=] {
                    for (uint64_t i = 0; i < m; ++i) copy_page(flat ? flat + (q + i) * RJ_PAGE_SIZE : dst_pages[q + i], buf->p + i * RJ_PAGE_SIZE);
                    copy_fence();
                    hp->down.release(buf);
                    g->done();
                });
            }
        } catch (...) {
            group.seal();
            try { group.wait(); } catch (...) {}
            throw;
        }
        group.seal();
        group.wait();
    });
}

void rj_result_free(rj_ctx* ctx, rj_result* r) {
    if (ctx) cudaSetDevice(ctx->device);
    delete r;
}

// ---- per-stage entry points ------------------------------------------------------------------------
int rj_page_row_offsets(rj_ctx* ctx, const void* d_pages, uint64_t n_pages, int32_t type, uint64_t* d_page_row_start,
                        uint64_t* d_totals, void* stream) {
    return guarded(ctx, [&] {
        cudaStream_t s = pick_stream(ctx, stream);
        Buf rows = dev_alloc(n_pages * 4, s), tmp = dev_alloc(scan_tmp_bytes(n_pages), s);
        if (d_totals) RJ_CUDA(cudaMemsetAsync(d_totals, 0, 16, s));
        launch_page_rows(d_pages, n_pages, type, rows->as<uint32_t>(), d_totals, s);
        launch_exclusive_scan_u32_u64(rows->as<uint32_t>(), d_page_row_start, n_pages, tmp->p, s);
    });
}

int rj_decode_fixed(rj_ctx* ctx, const void* d_pages, uint64_t n_pages, int32_t type, const uint64_t* d_page_row_start,
                    void* d_values, uint32_t* d_valid, void* stream) {
    return guarded(ctx, [&] {
        if (type == RJ_VARCHAR) throw EngineError("rj_decode_fixed: VARCHAR column");
        launch_decode_fixed(d_pages, n_pages, type, d_page_row_start, d_values, d_valid, ctx->sm_count, pick_stream(ctx, stream));
    });
}

int rj_decode_varchar(rj_ctx* ctx, const void* d_pages, uint64_t n_pages, const uint64_t* d_page_row_start,
                      uint64_t* d_desc, uint32_t* d_valid, void* stream) {
    return guarded(ctx, [&] {
        launch_decode_varchar(d_pages, n_pages, d_page_row_start, d_desc, d_valid, ctx->sm_count, pick_stream(ctx, stream));
    });
}

int rj_radix_histogram(rj_ctx* ctx, const void* d_keys, const uint32_t* d_valid, uint64_t n, int32_t key_bytes,
                       int32_t shift, int32_t bits, uint32_t* d_hist, void* stream) {
    return guarded(ctx, [&] {
        if ((key_bytes != 4 && key_bytes != 8) || bits < 0 || bits > kMaxTotalBits) throw EngineError("rj_radix_histogram: bad arguments");
        launch_radix_histogram(d_keys, d_valid, n, key_bytes, shift, bits, d_hist, ctx->sm_count, pick_stream(ctx, stream));
    });
}

int rj_radix_scatter(rj_ctx* ctx, const void* d_keys, const uint32_t* d_valid, const uint32_t* d_idx_in, uint64_t n,
                     int32_t key_bytes, int32_t shift, int32_t bits, uint32_t* d_cursor, void* d_keys_out,
                     uint32_t* d_idx_out, void* stream) {
    return guarded(ctx, [&] {
        if ((key_bytes != 4 && key_bytes != 8) || bits < 0 || bits > 8) throw EngineError("rj_radix_scatter: bits must be in [0, 8]");
        if (n >= 0xffffffffull) throw EngineError("relation exceeds 2^32-1 rows");
        launch_radix_scatter(d_keys, d_valid, d_idx_in, n, key_bytes, shift, bits, d_cursor, d_keys_out, d_idx_out,
                             ScatterPayload{}, ctx->sm_count, pick_stream(ctx, stream));
    });
}

int rj_radix_scatter_multi(rj_ctx* ctx, const void* d_keys, const uint32_t* d_valid, uint64_t n, int32_t key_bytes,
                           int32_t shift, int32_t bits, uint32_t* d_cursor, const rj_scatter_multi_t* out, void* stream) {
    return guarded(ctx, [&] {
        if ((key_bytes != 4 && key_bytes != 8) || bits < 0 || bits > 3 || !out) throw EngineError("rj_radix_scatter_multi: at most 8 partitions");
        if (out->n_payload > 6) throw EngineError("rj_radix_scatter_multi: at most 6 payload columns");
        if (n >= 0xffffffffull) throw EngineError("relation exceeds 2^32-1 rows");
        launch_radix_scatter_multi(d_keys, d_valid, n, key_bytes, shift, bits, d_cursor, *out, ctx->sm_count, pick_stream(ctx, stream));
    });
}

int rj_join_keys(rj_ctx* ctx, const void* d_build_keys, const uint32_t* d_build_valid, uint64_t n_build,
                 const void* d_probe_keys, const uint32_t* d_probe_valid, uint64_t n_probe, int32_t key_bytes,
                 uint64_t capacity, uint32_t* d_out_build, uint32_t* d_out_probe, uint64_t* n_matches, void* stream) {
    return guarded(ctx, [&] {
        if (key_bytes != 4 && key_bytes != 8) throw EngineError("rj_join_keys: key_bytes must be 4 or 8");
        // runs on the context stream; order it after the caller's stream
        cudaStream_t cs = pick_stream(ctx, stream);
        if (cs != ctx->stream) RJ_CUDA(cudaStreamSynchronize(cs));
        Exec ex(ctx, nullptr, nullptr);
        uint64_t m = 0;
        if (n_build == 0 || n_probe == 0) {
            *n_matches = 0;
            return;
        }
        JoinSide bs, ps;
        bs.keys = d_build_keys; bs.valid = d_build_valid; bs.n = n_build;
        ps.keys = d_probe_keys; ps.valid = d_probe_valid; ps.n = n_probe;
        ex.join_keys(bs, ps, key_bytes, &m);
        *n_matches = m;
        Buf ob = ex.side_rows(bs, m), op = ex.side_rows(ps, m);
        if (m <= capacity && m > 0) {
            RJ_CUDA(cudaMemcpyAsync(d_out_build, ob->p, m * 4, cudaMemcpyDeviceToDevice, ctx->stream));
            RJ_CUDA(cudaMemcpyAsync(d_out_probe, op->p, m * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        }
        RJ_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}

int rj_gather(rj_ctx* ctx, const void* d_src, const uint32_t* d_src_valid, const uint32_t* d_idx, uint64_t n,
              int32_t elem_bytes, void* d_out, uint32_t* d_out_valid, void* stream) {
    return guarded(ctx, [&] {
        if (elem_bytes != 4 && elem_bytes != 8) throw EngineError("rj_gather: elem_bytes must be 4 or 8");
        launch_gather(d_src, d_src_valid, d_idx, n, elem_bytes, d_out, d_out_valid, ctx->sm_count, pick_stream(ctx, stream));
    });
}

int rj_bitmap_to_bytes(rj_ctx* ctx, const uint32_t* d_bits, uint64_t n, uint8_t* d_bytes, void* stream) {
    return guarded(ctx, [&] { launch_bitmap_to_bytes(d_bits, n, d_bytes, ctx->sm_count, pick_stream(ctx, stream)); });
}

int rj_bytes_to_bitmap(rj_ctx* ctx, const uint8_t* d_bytes, uint64_t n, uint32_t* d_bits, void* stream) {
    return guarded(ctx, [&] { launch_bytes_to_bitmap(d_bytes, n, d_bits, ctx->sm_count, pick_stream(ctx, stream)); });
}

uint32_t rj_fixed_rows_per_page(int32_t type) { return type == RJ_INT32 ? 1984u : 1007u; }

int rj_encode_fixed(rj_ctx* ctx, const void* d_values, const uint32_t* d_valid, const uint32_t* d_idx, uint64_t n,
                    int32_t type, void* d_pages_out, void* stream) {
    return guarded(ctx, [&] {
        if (type == RJ_VARCHAR) throw EngineError("rj_encode_fixed: VARCHAR column");
        launch_encode_fixed(d_values, d_valid, nullptr, d_idx, d_idx, n, type, d_pages_out, ctx->sm_count, pick_stream(ctx, stream));
    });
}

struct rj_varchar_layout {
    VarcharLayoutDev L;
    std::vector<Buf> hold;
};

int rj_encode_varchar_plan(rj_ctx* ctx, const void* d_src_pages, const uint64_t* d_desc, const uint32_t* d_valid,
                           const uint32_t* d_idx, uint64_t n, rj_varchar_layout** layout, uint64_t* n_pages_out, void* stream) {
    return guarded(ctx, [&] {
        cudaStream_t s = pick_stream(ctx, stream);
        auto lay = std::make_unique<rj_varchar_layout>();
        VarcharLayoutDev& L = lay->L;
        L.n = n;
        L.src_pages = static_cast<const uint8_t*>(d_src_pages);
        L.desc = d_desc;
        L.valid = d_valid;
        L.idx = d_idx;
        Buf weights = dev_alloc(n * 8, s), wscan = dev_alloc(n * 8, s), marks = dev_alloc(n * 8, s), bscan = dev_alloc(n * 8, s);
        Buf heads = dev_alloc(n * 4, s), page_of = dev_alloc((n + 1) * 8, s), scalars = dev_alloc_zero(32, s);
        Buf tmp = dev_alloc(scan_tmp_bytes(n), s);
        L.weight_scan = wscan->as<uint64_t>();
        L.base_scan = bscan->as<uint64_t>();
        L.head_pages = heads->as<uint32_t>();
        L.page_of = page_of->as<uint64_t>();
        L.scalars = scalars->as<uint64_t>();
        launch_varchar_weights(L, weights->as<uint64_t>(), ctx->sm_count, s);
        launch_inclusive_sum_u64(weights->as<uint64_t>(), L.weight_scan, n, tmp->p, s);
        launch_varchar_marks(L, marks->as<uint64_t>(), ctx->sm_count, s);
        launch_inclusive_max_u64(marks->as<uint64_t>(), L.base_scan, n, tmp->p, s);
        launch_varchar_heads(L, weights->as<uint64_t>(), ctx->sm_count, s);
        launch_exclusive_scan_u32_u64(L.head_pages, L.page_of, n, tmp->p, s);
        uint64_t n_pages = 0;
        RJ_CUDA(cudaMemcpyAsync(&n_pages, L.page_of + n, 8, cudaMemcpyDeviceToHost, s));
        RJ_CUDA(cudaStreamSynchronize(s));
        L.n_pages = n_pages;
        Buf page_row = dev_alloc(n_pages * 4, s);
        L.page_row = page_row->as<uint32_t>();
        launch_varchar_page_rows(L, ctx->sm_count, s);
        lay->hold = {wscan, bscan, heads, page_of, scalars, page_row};
        *n_pages_out = n_pages;
        *layout = lay.release();
    });
}

int rj_encode_varchar_write(rj_ctx* ctx, rj_varchar_layout* layout, void* d_pages_out, void* stream) {
    return guarded(ctx, [&] {
        launch_varchar_write(layout->L, static_cast<uint8_t*>(d_pages_out), ctx->sm_count, pick_stream(ctx, stream));
    });
}

void rj_encode_varchar_free(rj_ctx* ctx, rj_varchar_layout* layout) {
    if (ctx) cudaSetDevice(ctx->device);
    delete layout;
}

int rj_gen_fixed_pages(rj_ctx* ctx, const void* d_values, const uint32_t* d_valid, uint64_t n, int32_t type,
                       void* d_pages_out, uint64_t* n_pages_out, void* stream) {
    return guarded(ctx, [&] {
        if (type == RJ_VARCHAR) throw EngineError("rj_gen_fixed_pages: VARCHAR column");
        const uint32_t rpp = rj_fixed_rows_per_page(type);
        if (n_pages_out) *n_pages_out = (n + rpp - 1) / rpp;
        if (d_pages_out) launch_encode_fixed(d_values, d_valid, nullptr, nullptr, nullptr, n, type, d_pages_out, ctx->sm_count, pick_stream(ctx, stream));
    });
}

// ---- whole-tuple scatter / join on partitioned inputs (multi-GPU: pass 1 is the exchange) -----------------
int rj_scatter_carry(rj_ctx* ctx, const rj_carry_scatter_t* d, void* stream) {
    return guarded(ctx, [&] {
        if (!d) throw EngineError("rj_scatter_carry: null descriptor");
        if (d->n_val > 2 || d->n_flag > 2 || d->n_owners > 8) throw EngineError("rj_scatter_carry: at most two value columns, two flags, eight owners");
        CarryScatter c;
        c.keys = static_cast<const uint32_t*>(d->d_keys);
        c.valid = d->d_valid;
        c.n = d->n;
        c.region_start = d->d_region_start;
        c.tile_start = d->d_tile_start;
        c.n_regions = d->n_regions;
        c.shift = d->shift;
        c.bits = d->bits;
        c.cursor = d->d_cursor;
        c.keys_out = static_cast<uint32_t*>(d->d_keys_out);
        c.n_val = static_cast<int>(d->n_val);
        c.n_flag = static_cast<int>(d->n_flag);
        for (int i = 0; i < 2; ++i) {
            c.val_src[i] = d->val_src[i];
            c.val_dst[i] = d->val_dst[i];
            c.val_width[i] = d->val_width[i];
            c.flag_src[i] = d->flag_src[i];
            c.flag_dst[i] = static_cast<uint8_t*>(d->flag_dst[i]);
        }
        c.src_tab = d->d_src_table;
        c.region_group = d->d_region_group;
        c.n_owners = static_cast<int>(d->n_owners);
        c.owner_shift = d->owner_shift;
        for (int o = 0; o < 8; ++o) {
            c.keys_dst_multi[o] = d->keys_dst_multi[o];
            for (int i = 0; i < 2; ++i) {
                c.val_dst_multi[i][o] = d->val_dst_multi[i][o];
                c.flag_dst_multi[i][o] = d->flag_dst_multi[i][o];
            }
        }
        launch_scatter_carry(c, ctx->sm_count, pick_stream(ctx, stream));
    });
}

namespace {
// second scatter pass + fused join on partitioned inputs; nullptr when a table met a duplicate build key.
// Runs on ctx->stream of the CURRENT device (the caller has bound device, home stream and block cache).
std::unique_ptr<rj_result> join_partitioned_impl(rj_ctx* ctx, const rj_part_side_t* build, const rj_part_side_t* probe,
                                                 const uint32_t* d_hist_build, const uint32_t* d_hist_probe, int32_t local_bits,
                                                 int32_t local_pass1_bits, int32_t hash_bits, const rj_part_out_t* outs, uint32_t n_out) {
    {
        if (local_bits < 0 || local_bits > kMaxTotalBits || local_pass1_bits < 0 || local_pass1_bits > kMaxPassBits ||
            local_bits - local_pass1_bits > kMaxPassBits || (local_pass1_bits && local_bits <= local_pass1_bits))
            throw EngineError("rj_join_partitioned: radix bits out of range");
        if (n_out < 1 || n_out > static_cast<uint32_t>(kEmitMaxOut) || build->n_cols > static_cast<uint32_t>(kEmitMaxPay) ||
            probe->n_cols > static_cast<uint32_t>(kEmitMaxPay))
            throw EngineError("rj_join_partitioned: too many columns for the fused join");
        if (build->n >= 0xffffffffull || probe->n >= 0xffffffffull) throw EngineError("relation exceeds 2^32-1 rows");
        cudaStream_t s = ctx->stream;
        const uint64_t nb = build->n, np = probe->n;
        auto width_of = [](int32_t t) {
            if (t == RJ_INT32) return 4;
            if (t == RJ_INT64 || t == RJ_FP64) return 8;
            throw EngineError("rj_join_partitioned: fixed-width columns only");
        };
        JoinEmitLaunch L;
        L.n_out = static_cast<int>(n_out);
        L.n_bpay = static_cast<int>(build->n_cols);
        L.n_ppay = static_cast<int>(probe->n_cols);
        for (int c = 0; c < L.n_bpay; ++c) {
            L.bwidth[c] = width_of(build->types[c]);
            L.bvalid[c] = build->d_valid_bytes[c];
        }
        for (int c = 0; c < L.n_ppay; ++c) {
            L.pwidth[c] = width_of(probe->types[c]);
            L.pvalid[c] = probe->d_valid_bytes[c];
        }
        auto res = std::make_unique<rj_result>();
        res->cols.resize(n_out);
        for (uint32_t a = 0; a < n_out; ++a) {
            const rj_part_side_t* sd = outs[a].side == 0 ? build : probe;
            if (outs[a].side != 0 && outs[a].side != 1) throw EngineError("rj_join_partitioned: output side is 0 or 1");
            if (outs[a].col < 0) {
                L.out_src[a] = 0;
                L.out_width[a] = 4;
                res->cols[a].type = RJ_INT32;
            } else {
                if (static_cast<uint32_t>(outs[a].col) >= sd->n_cols) throw EngineError("rj_join_partitioned: output column out of range");
                L.out_src[a] = outs[a].side == 0 ? 1 : 2;
                L.out_idx[a] = outs[a].col;
                L.out_width[a] = width_of(sd->types[outs[a].col]);
                L.out_nullable[a] = sd->d_valid_bytes[outs[a].col] != nullptr;
                res->cols[a].type = sd->types[outs[a].col];
            }
        }
        if (!join_emit_fits(L)) throw EngineError("rj_join_partitioned: the columns do not fit shared memory");
        if (nb == 0 || np == 0) return res; // src/execute.cpp:50: an empty side gives typed, page-less columns
        const uint32_t nparts = 1u << local_bits;
        const int      bits2 = local_bits - local_pass1_bits;
        Buf plan_mem = dev_alloc(partition_plan_words(local_bits, local_pass1_bits) * 4, s);
        PartitionPlanDev pl;
        partition_plan_carve(plan_mem->as<uint32_t>(), local_bits, local_pass1_bits, &pl);
        launch_partition_plan(d_hist_build, d_hist_probe, static_cast<uint32_t>(nb), static_cast<uint32_t>(np), local_bits, local_pass1_bits, 4,
                              pl, s, kEmitBuildCap, kEmitProbeChunk);
        struct SideFinal {
            const uint32_t* keys;
            const void*     val[kEmitMaxPay];
            const uint8_t*  ok[kEmitMaxPay];
            Buf             hold[1 + 2 * kEmitMaxPay];
        };
        auto finish_side = [&](const rj_part_side_t& sd, const int* widths, uint32_t* cur, const uint32_t* reg, const uint32_t* tile) {
            SideFinal f{};
            f.keys = static_cast<const uint32_t*>(sd.d_keys);
            for (uint32_t c = 0; c < sd.n_cols; ++c) {
                f.val[c] = sd.d_vals[c];
                f.ok[c] = sd.d_valid_bytes[c];
            }
            if (local_pass1_bits == 0) {
                if (sd.n_sub > 0) throw EngineError("rj_join_partitioned: sub-regions need a second pass");
                return f; // fully partitioned already
            }
            CarryScatter c2;
            c2.keys = f.keys;
            c2.n = sd.n;
            c2.region_start = reg; c2.tile_start = tile; c2.n_regions = 1u << local_pass1_bits;
            if (sd.n_sub > 0) {
                // pull: the regions are read where the senders' first pass left them (peer memory)
                if (!sd.d_src_table || !sd.d_sub_start || !sd.d_sub_tile || !sd.d_sub_group) throw EngineError("rj_join_partitioned: incomplete sub-region description");
                c2.region_start = sd.d_sub_start; c2.tile_start = sd.d_sub_tile; c2.n_regions = sd.n_sub;
                c2.src_tab = sd.d_src_table; c2.region_group = sd.d_sub_group;
            }
            c2.shift = 0; c2.bits = bits2; c2.cursor = cur;
            f.hold[0] = dev_alloc(sd.n * 4 + 64, s);
            c2.keys_out = f.hold[0]->as<uint32_t>();
            for (uint32_t c = 0; c < sd.n_cols; ++c) {
                f.hold[1 + c] = dev_alloc(sd.n * widths[c] + 64, s);
                c2.val_src[c2.n_val] = sd.d_vals[c];
                c2.val_dst[c2.n_val] = f.hold[1 + c]->p;
                c2.val_width[c2.n_val++] = widths[c];
                if (sd.d_valid_bytes[c]) {
                    f.hold[1 + kEmitMaxPay + c] = dev_alloc(sd.n + 64, s);
                    c2.flag_src[c2.n_flag] = sd.d_valid_bytes[c];
                    c2.flag_dst[c2.n_flag++] = f.hold[1 + kEmitMaxPay + c]->as<uint8_t>();
                }
            }
            launch_scatter_carry(c2, ctx->sm_count, s);
            f.keys = f.hold[0]->as<uint32_t>();
            for (uint32_t c = 0; c < sd.n_cols; ++c) {
                f.val[c] = f.hold[1 + c]->p;
                f.ok[c] = sd.d_valid_bytes[c] ? f.hold[1 + kEmitMaxPay + c]->as<uint8_t>() : nullptr;
            }
            return f;
        };
        SideFinal B, P;
        {
            uint64_t carried = 0;
            for (int c = 0; c < L.n_bpay; ++c) carried += 2 * nb * (L.bwidth[c] + (L.bvalid[c] ? 1 : 0));
            for (int c = 0; c < L.n_ppay; ++c) carried += 2 * np * (L.pwidth[c] + (L.pvalid[c] ? 1 : 0));
            StageScope sc(ctx, RJ_ST_SCATTER, s, local_pass1_bits ? 2 : 0, local_pass1_bits ? (nb + np) * 8 + carried : 0);
            B = finish_side(*build, L.bwidth, pl.cur_b, pl.reg_b, pl.tile_b);
            P = finish_side(*probe, L.pwidth, pl.cur_p, pl.reg_p, pl.tile_p);
        }
        const uint64_t max_chunks = join_emit_max_chunks(np, nparts, ctx->sm_count);
        Buf counters = dev_alloc_zero(32, s); // chunk counter @0, abort flag @4, rows @8
        L.bkeys = B.keys;
        L.pkeys = P.keys;
        L.off_b = pl.off_b; L.off_p = pl.off_p; L.unit_start = pl.unit_start; L.unit_cursor = pl.unit_cursor; L.unit_part = pl.unit_part; L.unit_part = pl.unit_part;
        L.nparts = nparts;
        L.part_bits = hash_bits;
        for (int c = 0; c < L.n_bpay; ++c) { L.bpay[c] = B.val[c]; L.bvalid[c] = B.ok[c]; }
        for (int c = 0; c < L.n_ppay; ++c) { L.ppay[c] = P.val[c]; L.pvalid[c] = P.ok[c]; }
        for (int a = 0; a < L.n_out; ++a) {
            ResultColumn& rc = res->cols[a];
            rc.pages = dev_alloc(max_chunks * (L.out_width[a] == 4 ? 1 : 2) * size_t(RJ_PAGE_SIZE), s);
            L.out_pages[a] = rc.pages->as<uint8_t>();
        }
        L.min_chunks_per_warp = kEmitMinChunksResident;
        L.chunk_counter = counters->as<uint32_t>();
        L.abort_flag = counters->as<uint32_t>() + 1;
        L.row_counter = reinterpret_cast<unsigned long long*>(counters->as<uint8_t>() + 8);
        RJ_CUDA(cudaMemsetAsync(pl.unit_cursor, 0, 4, s));
        uint64_t in_bytes = (nb + np) * 4;
        for (int c = 0; c < L.n_bpay; ++c) in_bytes += nb * (L.bwidth[c] + (L.bvalid[c] ? 1 : 0));
        for (int c = 0; c < L.n_ppay; ++c) in_bytes += np * (L.pwidth[c] + (L.pvalid[c] ? 1 : 0));
        {
            StageScope sc(ctx, RJ_ST_JOIN_EMIT, s, 1, in_bytes);
            launch_join_emit(L, np, ctx->sm_count, s);
        }
        uint32_t h[4] = {0, 0, 0, 0};
        RJ_CUDA(cudaMemcpyAsync(h, counters->p, 16, cudaMemcpyDeviceToHost, s));
        RJ_CUDA(cudaStreamSynchronize(s));
        if (h[1] != 0) return nullptr; // duplicate build keys
        const uint64_t chunks = h[0];
        if (chunks > max_chunks) throw EngineError("internal: the fused join produced more chunks than planned");
        res->num_rows = static_cast<uint64_t>(h[2]) | (static_cast<uint64_t>(h[3]) << 32);
        uint64_t out_page_bytes = 0;
        for (int a = 0; a < L.n_out; ++a) {
            ResultColumn& rc = res->cols[a];
            rc.n_pages = chunks * (L.out_width[a] == 4 ? 1 : 2);
            out_page_bytes += rc.n_pages * uint64_t(RJ_PAGE_SIZE);
            if (rc.n_pages == 0) rc.pages.reset();
        }
        if (ctx->profiling) ctx->stats[RJ_ST_JOIN_EMIT].bytes += out_page_bytes;
        return res;
    }
}
} // namespace

int rj_join_partitioned(rj_ctx* ctx, const rj_part_side_t* build, const rj_part_side_t* probe, const uint32_t* d_hist_build,
                        const uint32_t* d_hist_probe, int32_t local_bits, int32_t local_pass1_bits, int32_t hash_bits,
                        const rj_part_out_t* outs, uint32_t n_out, rj_result** out) {
    return guarded(ctx, [&] {
        if (!build || !probe || !outs || !out) throw EngineError("rj_join_partitioned: null argument");
        *out = nullptr;
        auto res = join_partitioned_impl(ctx, build, probe, d_hist_build, d_hist_probe, local_bits, local_pass1_bits, hash_bits, outs, n_out);
        *out = res.release(); // NULL: duplicate build keys
    });
}


int rj_dist_layout(rj_ctx* ctx, const uint32_t* d_hist, int32_t me, int32_t g, int32_t bits, int32_t pass1_bits, const uint64_t* d_ptrs,
                   const int32_t* d_widths, uint32_t* d_cursor, uint64_t* d_table, uint32_t* d_start, uint32_t* d_tile, uint32_t* d_group,
                   uint32_t* d_local_hist, uint64_t* d_scalars, void* stream) {
    return guarded(ctx, [&] {
        launch_dist_layout(d_hist, me, g, bits, pass1_bits, d_ptrs, d_widths, d_cursor, d_table, d_start, d_tile, d_group, d_local_hist,
                           reinterpret_cast<unsigned long long*>(d_scalars), pick_stream(ctx, stream));
    });
}

// ---- pre-filters ----------------------------------------------------------------------------------------
int rj_filter_compare(rj_ctx* ctx, const void* d_values, const uint32_t* d_valid, uint64_t n, int32_t type, int32_t op,
                      int64_t rhs_i, double rhs_d, uint32_t* d_out, void* stream) {
    return guarded(ctx, [&] { launch_filter_compare(d_values, d_valid, n, type, op, rhs_i, rhs_d, d_out, ctx->sm_count, pick_stream(ctx, stream)); });
}

int rj_filter_varchar(rj_ctx* ctx, const void* d_pages, const uint64_t* d_desc, const uint32_t* d_valid, uint64_t n, int32_t op,
                      const char* rhs, uint64_t rhs_len, uint32_t* d_out, void* stream) {
    return guarded(ctx, [&] {
        cudaStream_t st = pick_stream(ctx, stream);
        if (rhs_len > 0xffffffffull) throw EngineError("rj_filter_varchar: literal too long");
        Buf lit = dev_alloc(rhs_len + 16, st);
        if (rhs_len) RJ_CUDA(cudaMemcpyAsync(lit->p, rhs, rhs_len, cudaMemcpyHostToDevice, st)); // pageable source: staged before the call returns
        launch_filter_varchar(d_pages, d_desc, d_valid, n, op, lit->as<uint8_t>(), static_cast<uint32_t>(rhs_len), d_out, ctx->sm_count, st);
    });
}

int rj_filter_null(rj_ctx* ctx, const uint32_t* d_valid, uint64_t n, int32_t is_null, uint32_t* d_out, void* stream) {
    return guarded(ctx, [&] { launch_filter_null(d_valid, n, is_null != 0, d_out, ctx->sm_count, pick_stream(ctx, stream)); });
}

int rj_bitmap_logic(rj_ctx* ctx, const uint32_t* d_a, const uint32_t* d_b, uint64_t n, int32_t op, uint32_t* d_out, void* stream) {
    return guarded(ctx, [&] { launch_bitmap_logic(d_a, d_b, n, op, d_out, ctx->sm_count, pick_stream(ctx, stream)); });
}

namespace {
// row ids of the set bits, ascending; returns their number (synchronises the stream)
uint64_t bitmap_select(rj_ctx* ctx, const uint32_t* bits, uint64_t n, uint32_t* ids, cudaStream_t st) {
    const uint64_t n_words = (n + 31) / 32;
    if (n_words == 0) return 0;
    Buf counts = dev_alloc(n_words * 4, st);
    Buf start = dev_alloc((n_words + 1) * 8, st);
    Buf tmp = dev_alloc(scan_tmp_bytes(n_words), st);
    launch_bitmap_popc(bits, n_words, counts->as<uint32_t>(), ctx->sm_count, st);
    launch_exclusive_scan_u32_u64(counts->as<uint32_t>(), start->as<uint64_t>(), n_words, tmp->p, st);
    launch_bitmap_expand(bits, start->as<uint64_t>(), n_words, ids, ctx->sm_count, st);
    uint64_t total = 0;
    RJ_CUDA(cudaMemcpyAsync(&total, start->as<uint64_t>() + n_words, 8, cudaMemcpyDeviceToHost, st));
    RJ_CUDA(cudaStreamSynchronize(st));
    return total;
}
} // namespace

int rj_bitmap_select(rj_ctx* ctx, const uint32_t* d_bits, uint64_t n, uint32_t* d_row_ids, uint64_t* count, void* stream) {
    return guarded(ctx, [&] {
        const uint64_t m = bitmap_select(ctx, d_bits, n, d_row_ids, pick_stream(ctx, stream));
        if (count) *count = m;
    });
}

int rj_filter_table(rj_ctx* ctx, const rj_table_t* table, const rj_pred_t* prog, uint32_t n_prog, rj_result** out) {
    return guarded(ctx, [&] {
        if (!table || !out) throw EngineError("rj_filter_table: null argument");
        if (n_prog && !prog) throw EngineError("rj_filter_table: null program");
        if (table->num_rows >= 0xffffffffull) throw EngineError("relation exceeds 2^32-1 rows");
        // a one-node plan: Scan(table 0) with every column as output, in order
        std::vector<rj_attr_t> attrs(table->n_columns);
        for (uint32_t c = 0; c < table->n_columns; ++c) attrs[c] = rj_attr_t{c, table->columns[c].type, 0};
        rj_node_t scan{};
        scan.n_output_attrs = table->n_columns;
        scan.output_attrs = attrs.data();
        rj_plan_t plan{};
        plan.n_nodes = 1;
        plan.n_inputs = 1;
        plan.nodes = &scan;
        plan.inputs = table;
        plan.root = 0;
        auto in = upload_tables(ctx, table, 1, nullptr);
        Exec ex(ctx, &plan, in.get());
        cudaStream_t st = ctx->stream;
        const uint64_t n = table->num_rows;
        const uint64_t words = (n + 31) / 32;
        Rel r;
        r.rows = n;
        r.rid[0] = nullptr; // identity: no filter keeps every row
        if (n_prog) {
            std::vector<Buf> stack;
            for (uint32_t i = 0; i < n_prog; ++i) {
                const rj_pred_t& pr = prog[i];
                Buf res = dev_alloc(words * 4 + 16, st);
                if (pr.kind == 0) {
                    if (pr.column >= table->n_columns) throw EngineError("rj_filter_table: predicate column out of range");
                    const DecodedCol& col = ex.column(0, pr.column);
                    if (pr.op == RJ_OP_IS_NULL || pr.op == RJ_OP_IS_NOT_NULL) {
                        launch_filter_null(col.valid_ptr(), n, pr.op == RJ_OP_IS_NULL, res->as<uint32_t>(), ctx->sm_count, st);
                    } else if (col.type == RJ_VARCHAR) {
                        if (pr.lit_type != RJ_VARCHAR) throw EngineError("rj_filter_table: a VARCHAR column is compared with a string literal");
                        if (pr.rhs_s_len > 0xffffffffull) throw EngineError("rj_filter_table: literal too long");
                        Buf lit = dev_alloc(pr.rhs_s_len + 16, st);
                        if (pr.rhs_s_len) RJ_CUDA(cudaMemcpyAsync(lit->p, pr.rhs_s, pr.rhs_s_len, cudaMemcpyHostToDevice, st));
                        launch_filter_varchar(col.pages, col.values->as<uint64_t>(), col.valid_ptr(), n, pr.op, lit->as<uint8_t>(),
                                              static_cast<uint32_t>(pr.rhs_s_len), res->as<uint32_t>(), ctx->sm_count, st);
                    } else {
                        // std::get<int64_t> / std::get<double> of the literal (statement.cpp:55,74,93): the wrong
                        // alternative throws in the reference
                        if (col.type == RJ_FP64 ? pr.lit_type != RJ_FP64 : pr.lit_type != RJ_INT64) throw EngineError("bad_variant_access");
                        launch_filter_compare(col.values->p, col.valid_ptr(), n, col.type, pr.op, pr.rhs_i, pr.rhs_d, res->as<uint32_t>(), ctx->sm_count, st);
                    }
                } else if (pr.kind == 1) {
                    const size_t need = pr.op == RJ_LOGIC_NOT ? 1 : 2;
                    if (stack.size() < need) throw EngineError("rj_filter_table: malformed program (stack underflow)");
                    Buf b = pr.op == RJ_LOGIC_NOT ? nullptr : stack.back();
                    if (need == 2) stack.pop_back();
                    Buf a = stack.back();
                    stack.pop_back();
                    launch_bitmap_logic(a->as<uint32_t>(), b ? b->as<uint32_t>() : nullptr, n, pr.op, res->as<uint32_t>(), ctx->sm_count, st);
                } else {
                    throw EngineError("rj_filter_table: unknown program entry");
                }
                stack.push_back(std::move(res));
            }
            if (stack.size() != 1) throw EngineError("rj_filter_table: malformed program (one bitmap must remain)");
            Buf ids = dev_alloc(std::max<uint64_t>(n, 1) * 4, st);
            r.rows = bitmap_select(ctx, stack.back()->as<uint32_t>(), n, ids->as<uint32_t>(), st);
            r.rid[0] = ids;
        }
        auto res = ex.root(0, r);
        RJ_CUDA(cudaStreamSynchronize(st));
        *out = res.release();
    });
}

int rj_varchar_descriptors(rj_ctx* ctx, const uint64_t* d_offsets, uint64_t n, uint64_t* d_desc, void* stream) {
    return guarded(ctx, [&] { launch_varchar_desc_from_offsets(d_offsets, n, d_desc, ctx->sm_count, pick_stream(ctx, stream)); });
}

// ---- result validation ------------------------------------------------------------------------------------
int rj_tables_equal(rj_ctx* ctx, const rj_table_t* a, const rj_table_t* b, int32_t* equal, uint64_t* mismatches) {
    return guarded(ctx, [&] {
        if (!a || !b || !equal) throw EngineError("rj_tables_equal: null argument");
        *equal = 0;
        if (mismatches) *mismatches = 0;
        if (a->n_columns != b->n_columns || a->num_rows != b->num_rows) return;
        for (uint32_t c = 0; c < a->n_columns; ++c)
            if (a->columns[c].type != b->columns[c].type) return;
        const uint64_t n = a->num_rows;
        if (n >= 0xffffffffull) throw EngineError("relation exceeds 2^32-1 rows");
        if (n == 0 || a->n_columns == 0) {
            *equal = 1;
            return;
        }
        rj_table_t tabs[2] = {*a, *b};
        auto in = upload_tables(ctx, tabs, 2, nullptr);
        rj_plan_t none{};
        Exec ex(ctx, &none, in.get());
        cudaStream_t s = ctx->stream;
        Buf h[2], idx[2];
        Buf alt_k = dev_alloc(n * 8, s), alt_v = dev_alloc(n * 4, s);
        Buf counts = dev_alloc(sort_tmp_words(n) * 4, s), base = dev_alloc((sort_tmp_words(n) + 1) * 8, s);
        Buf scan_tmp = dev_alloc(scan_tmp_bytes(std::max<uint64_t>(sort_tmp_words(n), n)), s);
        Buf cell = dev_alloc(n * 8, s);
        for (uint32_t t = 0; t < 2; ++t) {
            h[t] = dev_alloc_zero(n * 8, s);
            idx[t] = dev_alloc(n * 4, s);
            for (uint32_t c = 0; c < a->n_columns; ++c) {
                const DecodedCol& col = ex.column(t, c);
                if (col.type == RJ_VARCHAR) {
                    launch_varchar_hash(col.pages, col.values->as<uint64_t>(), col.valid_ptr(), n, cell->as<uint64_t>(), ctx->sm_count, s);
                    launch_hash_combine(cell->as<uint64_t>(), col.valid_ptr(), n, h[t]->as<uint64_t>(), ctx->sm_count, s);
                } else {
                    launch_hash_fixed_cells(col.values->p, col.valid_ptr(), n, col.type == RJ_INT32 ? 4 : 8, h[t]->as<uint64_t>(), ctx->sm_count, s);
                }
            }
            launch_iota_u32(idx[t]->as<uint32_t>(), n, ctx->sm_count, s);
            launch_radix_sort_u64(h[t]->as<uint64_t>(), idx[t]->as<uint32_t>(), alt_k->as<uint64_t>(), alt_v->as<uint32_t>(), n,
                                  counts->as<uint32_t>(), base->as<uint64_t>(), scan_tmp->p, s);
        }
        Buf bad = dev_alloc_zero(8, s);
        auto* d_bad = bad->as<unsigned long long>();
        launch_keys_differ(h[0]->as<uint64_t>(), h[1]->as<uint64_t>(), n, d_bad, ctx->sm_count, s);
        Buf keep = dev_alloc(n * 4, s);
        for (uint32_t c = 0; c < a->n_columns; ++c) {
            const DecodedCol& ca = ex.column(0, c);
            const DecodedCol& cb = ex.column(1, c);
            if (ca.type == RJ_VARCHAR) {
                launch_varchar_pairs_equal(ca.pages, ca.values->as<uint64_t>(), idx[0]->as<uint32_t>(), cb.pages, cb.values->as<uint64_t>(),
                                           idx[1]->as<uint32_t>(), n, keep->as<uint32_t>(), ctx->sm_count, s);
                launch_pairs_equal_varchar_finish(ca.valid_ptr(), idx[0]->as<uint32_t>(), cb.valid_ptr(), idx[1]->as<uint32_t>(), keep->as<uint32_t>(), n,
                                                  d_bad, ctx->sm_count, s);
            } else {
                launch_pairs_equal_fixed(ca.values->p, ca.valid_ptr(), idx[0]->as<uint32_t>(), cb.values->p, cb.valid_ptr(), idx[1]->as<uint32_t>(), n,
                                         ca.type == RJ_INT32 ? 4 : 8, d_bad, ctx->sm_count, s);
            }
        }
        unsigned long long n_bad = 0;
        RJ_CUDA(cudaMemcpyAsync(&n_bad, d_bad, 8, cudaMemcpyDeviceToHost, s));
        RJ_CUDA(cudaStreamSynchronize(s));
        if (mismatches) *mismatches = n_bad;
        *equal = n_bad == 0 ? 1 : 0;
    });
}

// ---- profiling -------------------------------------------------------------------------------------
int rj_profile_enable(rj_ctx* ctx, int on) {
    if (!ctx) return 1;
    ctx->profiling = on != 0;
    return 0;
}

int rj_profile_reset(rj_ctx* ctx) {
    if (!ctx) return 1;
    profile_collect(ctx);
    std::memset(ctx->stats, 0, sizeof ctx->stats);
    return 0;
}

int rj_profile_read(rj_ctx* ctx, rj_stage_stat_t* stats) {
    if (!ctx || !stats) return 1;
    cudaSetDevice(ctx->device);
    profile_collect(ctx);
    std::memcpy(stats, ctx->stats, sizeof ctx->stats);
    return 0;
}

const char* rj_stage_name(int stage) {
    static const char* names[RJ_ST_COUNT] = {"h2d", "row_offsets", "decode", "histogram", "scatter",
                                             "join", "gather", "encode", "d2h", "join_emit"};
    return stage >= 0 && stage < RJ_ST_COUNT ? names[stage] : "?";
}

} // extern "C"
