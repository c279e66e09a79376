// The drop-in: Contest::build_context / destroy_context / execute with the exact signatures of the
// reference's include/plan.h:337-344, implemented on the B200 engine's C-ABI (include/rj_b200.h).
//
// A contest harness (tests/read_sql.cpp:1224-1249, tests/unit_tests.cpp) links this translation unit
// INSTEAD of the reference's src/execute.cpp; nothing else changes.  It is compiled against the
// reference's own headers (-I$REF/include), so Plan / ColumnarTable / Column / Page are byte-compatible
// by construction.  The adapter only flattens the plan into plain structs and wraps the result pages:
//   * inputs are borrowed (`const Plan&`): page pointers are passed through, nothing is copied here;
//   * every output page is `new Page` because Column::~Column deletes them (plan.h:64-68,95-99): the
//     engine asks for them through rj_page_alloc_t and fills them while the transfers are running
//     (rj_execute_pages: gather + H2D, kernels, D2H + scatter all overlap);
//   * engine errors surface as std::runtime_error, the reference's error contract
//     (src/execute.cpp:280; the harness catches std::exception, tests/read_sql.cpp:1329-1332).
#include <malloc.h>

#include <cstdlib>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

#include <plan.h>

#include "rj_b200.h"

namespace Contest {

void* build_context() {
    // Result pages are individually `new`-ed 8 KB objects, millions per large result, freed by the
    // caller before the next query.  Keep that memory mapped between queries instead of returning it to
    // the kernel and faulting it in again: the first touch of 11 GB costs more than the join.
    mallopt(M_TRIM_THRESHOLD, std::numeric_limits<int>::max());
    mallopt(M_TOP_PAD, 256 << 20);
    mallopt(M_MMAP_THRESHOLD, 1 << 30);
    rj_ctx* ctx = nullptr;
    // RJ_GPUS=N (a power of two): one context over devices 0..N-1; large key / foreign-key joins of two scans
    // then run on all of them (rj_ctx_create_multi), everything else on device 0
    int n_gpus = 1;
    if (const char* env = std::getenv("RJ_GPUS")) n_gpus = std::atoi(env);
    if (n_gpus > 1) {
        std::vector<int> devices(n_gpus);
        for (int i = 0; i < n_gpus; ++i) devices[i] = i;
        if (rj_ctx_create_multi(devices.data(), static_cast<uint32_t>(n_gpus), &ctx) != 0) {
            throw std::runtime_error(std::string("rj_ctx_create_multi: ") + rj_last_error(nullptr));
        }
        return ctx;
    }
    if (rj_ctx_create(0, &ctx) != 0) {
        throw std::runtime_error(std::string("rj_ctx_create: ") + rj_last_error(nullptr));
    }
    return ctx;
}

void destroy_context(void* context) { rj_ctx_destroy(static_cast<rj_ctx*>(context)); }

ColumnarTable execute(const Plan& plan, void* context) {
    auto* ctx = static_cast<rj_ctx*>(context);
    if (ctx == nullptr) {
        throw std::runtime_error("Contest::execute: context is null (call build_context first)");
    }
    // ---- flatten ---------------------------------------------------------------------------------
    std::vector<std::vector<rj_attr_t>> attrs(plan.nodes.size());
    std::vector<rj_node_t>              nodes(plan.nodes.size());
    for (size_t i = 0; i < plan.nodes.size(); ++i) {
        const PlanNode& n = plan.nodes[i];
        for (auto [idx, type]: n.output_attrs) {
            attrs[i].push_back(rj_attr_t{idx, static_cast<int32_t>(type), 0});
        }
        rj_node_t& f     = nodes[i];
        f                = rj_node_t{};
        f.n_output_attrs = static_cast<uint32_t>(attrs[i].size());
        f.output_attrs   = attrs[i].data();
        if (const auto* j = std::get_if<JoinNode>(&n.data)) {
            f.is_join    = 1;
            f.build_left = j->build_left ? 1 : 0;
            f.left       = j->left;
            f.right      = j->right;
            f.left_attr  = j->left_attr;
            f.right_attr = j->right_attr;
        } else {
            f.base_table_id = std::get<ScanNode>(n.data).base_table_id;
        }
    }
    std::vector<std::vector<rj_column_t>> cols(plan.inputs.size());
    std::vector<rj_table_t>               tables(plan.inputs.size());
    for (size_t t = 0; t < plan.inputs.size(); ++t) {
        const ColumnarTable& in = plan.inputs[t];
        for (const Column& c: in.columns) {
            rj_column_t f{};
            f.type    = static_cast<int32_t>(c.type);
            f.n_pages = c.pages.size();
            // std::vector<Page*> is an array of page pointers; Page::data is its first member
            f.pages = reinterpret_cast<const void* const*>(c.pages.data());
            cols[t].push_back(f);
        }
        tables[t].num_rows  = in.num_rows;
        tables[t].n_columns = static_cast<uint32_t>(cols[t].size());
        tables[t].columns   = cols[t].data();
    }
    rj_plan_t flat{};
    flat.n_nodes  = static_cast<uint32_t>(nodes.size());
    flat.n_inputs = static_cast<uint32_t>(tables.size());
    flat.nodes    = nodes.data();
    flat.inputs   = tables.data();
    flat.root     = plan.root;

    // ---- run: every result page is a `new Page` (Column::~Column deletes them, plan.h:95-99) ---------
    // The typed, possibly page-less result columns exist before the engine runs (tests/unit_tests.cpp:24-27).
    struct Sink {
        ColumnarTable out;
        static int new_pages(void*, uint64_t n, void** pages) {
            uint64_t i = 0;
            try {
                for (; i < n; ++i) pages[i] = new Page;
            } catch (...) {
                while (i > 0) delete static_cast<Page*>(pages[--i]);
                return 1;
            }
            return 0;
        }
        static int append(void* user, uint32_t column, int32_t, void* const* pages, uint64_t n) {
            auto& cols = static_cast<Sink*>(user)->out.columns;
            if (column >= cols.size()) return 1;
            auto& v = cols[column].pages;
            v.reserve(v.size() + n);
            for (uint64_t i = 0; i < n; ++i) v.push_back(static_cast<Page*>(pages[i]));
            return 0;
        }
        static void free_pages(void*, uint64_t n, void* const* pages) {
            for (uint64_t i = 0; i < n; ++i) delete static_cast<Page*>(pages[i]);
        }
    } sink;
    if (plan.root >= plan.nodes.size()) {
        throw std::runtime_error("plan root out of range");
    }
    for (auto [idx, type]: plan.nodes[plan.root].output_attrs) {
        sink.out.columns.emplace_back(type);
    }
    rj_page_alloc_t alloc{&sink, &Sink::new_pages, &Sink::append, &Sink::free_pages};
    uint64_t        num_rows = 0;
    if (rj_execute_pages(ctx, &flat, 0, &alloc, &num_rows) != 0) {
        throw std::runtime_error(rj_last_error(ctx)); // ~Sink deletes whatever was appended
    }
    sink.out.num_rows = num_rows;
    return std::move(sink.out);
}

} // namespace Contest
