// The drop-in: Contest::build_context / destroy_context / execute with the exact signatures of the
// reference's include/plan.h:337-344, implemented on the B200 engine's C-ABI (include/rj_b200.h).
//
// A contest harness (tests/read_sql.cpp:1224-1249, tests/unit_tests.cpp) links this translation unit
// INSTEAD of the reference's src/execute.cpp; nothing else changes.  It is compiled against the
// reference's own headers (-I$REF/include), so Plan / ColumnarTable / Column / Page are byte-compatible
// by construction.  The adapter only flattens the plan into plain structs and wraps the result pages:
//   * inputs are borrowed (`const Plan&`): page pointers are passed through, nothing is copied here;
//   * every output page is `new Page` because Column::~Column deletes them (plan.h:64-68,95-99);
//   * engine errors surface as std::runtime_error, the reference's error contract
//     (src/execute.cpp:280; the harness catches std::exception, tests/read_sql.cpp:1329-1332).
#include <stdexcept>
#include <string>
#include <vector>

#include <plan.h>

#include "rj_b200.h"

namespace Contest {

void* build_context() {
    rj_ctx* ctx = nullptr;
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

    // ---- run ---------------------------------------------------------------------------------------
    rj_result* res = nullptr;
    if (rj_execute(ctx, &flat, &res) != 0) {
        throw std::runtime_error(rj_last_error(ctx));
    }
    // ---- wrap: typed columns of freshly allocated pages ---------------------------------------------
    ColumnarTable out;
    out.num_rows = rj_result_num_rows(res);
    std::string error;
    for (uint32_t c = 0; c < rj_result_num_columns(res); ++c) {
        out.columns.emplace_back(static_cast<DataType>(rj_result_column_type(res, c)));
        Column&  col     = out.columns.back();
        uint64_t n_pages = rj_result_column_pages(res, c);
        col.pages.reserve(n_pages);
        for (uint64_t p = 0; p < n_pages; ++p) {
            col.new_page();
        }
        if (n_pages != 0
            && rj_result_fetch(ctx, res, c, reinterpret_cast<void* const*>(col.pages.data()), nullptr) != 0) {
            error = rj_last_error(ctx);
            break;
        }
    }
    rj_result_free(ctx, res);
    if (!error.empty()) {
        throw std::runtime_error(error);
    }
    return out;
}

} // namespace Contest
