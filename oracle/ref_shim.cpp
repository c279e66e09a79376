// TEST INFRASTRUCTURE ONLY -- C-ABI wrapper around the UNMODIFIED reference.
//
// Compiled by oracle/Makefile together with /root/reference/src/{execute,build_table,statement,
// csv_parser}.cpp (read where they lie; nothing is copied into this repo) into
// oracle/_ref/libref_oracle.so.  It lets a Python test hand the same flattened plan
// (include/rj_b200.h) to the reference's own Contest::execute (src/execute.cpp:316-324) and to the
// CUDA engine, and exposes the reference's page codec (Table::to_columnar / Table::from_columnar,
// src/build_table.cpp:312-681) so the C restatement can be pinned byte for byte.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include <omp.h>

#include <inner_column.h>
#include <plan.h>
#include <table.h>

#include "oracle.h"

struct orc_result {
    uint64_t                          num_rows = 0;
    std::vector<int32_t>              types;
    std::vector<std::vector<uint8_t>> data; // contiguous pages per column
};

namespace {

thread_local double g_last_seconds = 0.0;

void set_err(char* err, size_t errlen, const char* msg) {
    if (err && errlen) {
        std::snprintf(err, errlen, "%s", msg);
    }
}

const void* page_ptr(const rj_column_t& c, uint64_t i) {
    if (c.pages) {
        return c.pages[i];
    }
    return static_cast<const uint8_t*>(c.contiguous) + i * PAGE_SIZE;
}

ColumnarTable make_table(const rj_table_t& t) {
    ColumnarTable out;
    out.num_rows = t.num_rows;
    for (uint32_t c = 0; c < t.n_columns; ++c) {
        const rj_column_t& col = t.columns[c];
        out.columns.emplace_back(static_cast<DataType>(col.type));
        for (uint64_t p = 0; p < col.n_pages; ++p) {
            Page* page = out.columns.back().new_page();
            std::memcpy(page->data, page_ptr(col, p), PAGE_SIZE);
        }
    }
    return out;
}

orc_result* flatten(const ColumnarTable& t) {
    auto* r     = new orc_result;
    r->num_rows = t.num_rows;
    for (auto& col: t.columns) {
        r->types.push_back(static_cast<int32_t>(col.type));
        std::vector<uint8_t> bytes(col.pages.size() * PAGE_SIZE);
        for (size_t p = 0; p < col.pages.size(); ++p) {
            std::memcpy(bytes.data() + p * PAGE_SIZE, col.pages[p]->data, PAGE_SIZE);
        }
        r->data.emplace_back(std::move(bytes));
    }
    return r;
}

Table cells_to_table(const orc_cells_t* cols, uint32_t n_cols, uint64_t n_rows) {
    std::vector<std::vector<Data>> rows(n_rows, std::vector<Data>(n_cols, std::monostate{}));
    std::vector<DataType>          types;
    for (uint32_t c = 0; c < n_cols; ++c) {
        const orc_cells_t& col = cols[c];
        types.push_back(static_cast<DataType>(col.type));
        for (uint64_t i = 0; i < n_rows; ++i) {
            if (!col.valid[i]) {
                continue;
            }
            switch (col.type) {
            case RJ_INT32: rows[i][c] = static_cast<const int32_t*>(col.values)[i]; break;
            case RJ_INT64: rows[i][c] = static_cast<const int64_t*>(col.values)[i]; break;
            case RJ_FP64:  rows[i][c] = static_cast<const double*>(col.values)[i]; break;
            default:
                rows[i][c] = std::string(col.chars + col.str_off[i], col.chars + col.str_off[i + 1]);
            }
        }
    }
    return Table(std::move(rows), std::move(types));
}

Table decode_one(const rj_column_t* col, uint64_t num_rows) {
    rj_table_t t{};
    t.num_rows  = num_rows;
    t.n_columns = 1;
    t.columns   = col;
    return Table::from_columnar(make_table(t));
}

} // namespace

extern "C" {

int ref_execute(const rj_plan_t* p, int n_threads, orc_result** out, char* err, size_t errlen) {
    try {
        Plan plan;
        for (uint32_t i = 0; i < p->n_inputs; ++i) {
            plan.new_input(make_table(p->inputs[i]));
        }
        for (uint32_t i = 0; i < p->n_nodes; ++i) {
            const rj_node_t&                          n = p->nodes[i];
            std::vector<std::tuple<size_t, DataType>> attrs;
            for (uint32_t a = 0; a < n.n_output_attrs; ++a) {
                attrs.emplace_back(n.output_attrs[a].index,
                    static_cast<DataType>(n.output_attrs[a].type));
            }
            if (n.is_join) {
                plan.new_join_node(n.build_left != 0, n.left, n.right, n.left_attr, n.right_attr,
                    std::move(attrs));
            } else {
                plan.new_scan_node(n.base_table_id, std::move(attrs));
            }
        }
        plan.root = p->root;
        if (n_threads > 0) {
            omp_set_num_threads(n_threads);
        }
        void* ctx = Contest::build_context();
        auto  t0  = std::chrono::steady_clock::now();
        auto  res = Contest::execute(plan, ctx);
        auto  t1  = std::chrono::steady_clock::now();
        Contest::destroy_context(ctx);
        g_last_seconds = std::chrono::duration<double>(t1 - t0).count();
        *out           = flatten(res);
        return 0;
    } catch (const std::exception& e) {
        set_err(err, errlen, e.what());
        return 1;
    }
}

double ref_last_execute_seconds(void) { return g_last_seconds; }

int ref_encode(const orc_cells_t* cols, uint32_t n_cols, uint64_t n_rows, orc_result** out,
    char* err, size_t errlen) {
    try {
        *out = flatten(cells_to_table(cols, n_cols, n_rows).to_columnar());
        return 0;
    } catch (const std::exception& e) {
        set_err(err, errlen, e.what());
        return 1;
    }
}

uint64_t ref_result_num_rows(const orc_result* r) { return r->num_rows; }
uint32_t ref_result_num_columns(const orc_result* r) { return static_cast<uint32_t>(r->types.size()); }
int32_t  ref_result_column_type(const orc_result* r, uint32_t c) { return r->types[c]; }
uint64_t ref_result_column_pages(const orc_result* r, uint32_t c) { return r->data[c].size() / PAGE_SIZE; }
const void* ref_result_column_data(const orc_result* r, uint32_t c) { return r->data[c].data(); }
void ref_result_free(orc_result* r) { delete r; }

int ref_decode_sizes(const rj_column_t* col, uint64_t num_rows, uint64_t* n_chars, char* err,
    size_t errlen) {
    try {
        uint64_t chars = 0;
        if (col->type == RJ_VARCHAR) {
            Table t = decode_one(col, num_rows);
            for (auto& row: t.table()) {
                if (auto* s = std::get_if<std::string>(&row[0])) {
                    chars += s->size();
                }
            }
        }
        *n_chars = chars;
        return 0;
    } catch (const std::exception& e) {
        set_err(err, errlen, e.what());
        return 1;
    }
}

int ref_decode_fill(const rj_column_t* col, uint64_t num_rows, uint8_t* valid, void* values,
    uint64_t* str_off, char* chars, char* err, size_t errlen) {
    try {
        Table    t   = decode_one(col, num_rows);
        uint64_t pos = 0;
        for (uint64_t i = 0; i < num_rows; ++i) {
            const Data& d = t.table()[i][0];
            valid[i]      = std::holds_alternative<std::monostate>(d) ? 0 : 1;
            if (col->type == RJ_VARCHAR) {
                str_off[i] = pos;
                if (auto* s = std::get_if<std::string>(&d)) {
                    std::memcpy(chars + pos, s->data(), s->size());
                    pos += s->size();
                }
            } else if (col->type == RJ_INT32) {
                auto* v = std::get_if<int32_t>(&d);
                static_cast<int32_t*>(values)[i] = v ? *v : 0;
            } else if (col->type == RJ_INT64) {
                auto* v = std::get_if<int64_t>(&d);
                static_cast<int64_t*>(values)[i] = v ? *v : 0;
            } else {
                auto* v = std::get_if<double>(&d);
                static_cast<double*>(values)[i] = v ? *v : 0.0;
            }
        }
        if (col->type == RJ_VARCHAR) {
            str_off[num_rows] = pos;
        }
        return 0;
    } catch (const std::exception& e) {
        set_err(err, errlen, e.what());
        return 1;
    }
}

// Pre-filter oracle: the reference's own COLUMN-WISE evaluation (Comparison::eval / LogicalOperation::eval
// over InnerColumns, src/statement.cpp:46-133,186-200 -- what Table::from_csv runs, build_table.cpp:255)
// on the decoded table.  `prog` is the postfix program of include/rj_b200.h; selected[i] = result bit i.
int ref_filter(const rj_table_t* t, const rj_pred_t* prog, uint32_t n_prog, uint8_t* selected, char* err, size_t errlen) {
    try {
        Table tab = Table::from_columnar(make_table(*t));
        std::vector<std::unique_ptr<InnerColumnBase>> owned;
        std::vector<const InnerColumnBase*>           cols;
        for (uint32_t c = 0; c < t->n_columns; ++c) {
            switch (static_cast<DataType>(t->columns[c].type)) {
            case DataType::INT32: {
                auto col = std::make_unique<InnerColumn<int32_t>>();
                for (auto& row: tab.table()) { if (auto* v = std::get_if<int32_t>(&row[c])) col->push_back(*v); else col->push_back_null(); }
                owned.push_back(std::move(col));
                break;
            }
            case DataType::INT64: {
                auto col = std::make_unique<InnerColumn<int64_t>>();
                for (auto& row: tab.table()) { if (auto* v = std::get_if<int64_t>(&row[c])) col->push_back(*v); else col->push_back_null(); }
                owned.push_back(std::move(col));
                break;
            }
            case DataType::FP64: {
                auto col = std::make_unique<InnerColumn<double>>();
                for (auto& row: tab.table()) { if (auto* v = std::get_if<double>(&row[c])) col->push_back(*v); else col->push_back_null(); }
                owned.push_back(std::move(col));
                break;
            }
            case DataType::VARCHAR: {
                auto col = std::make_unique<InnerColumn<std::string>>();
                for (auto& row: tab.table()) { if (auto* v = std::get_if<std::string>(&row[c])) col->push_back(*v); else col->push_back_null(); }
                owned.push_back(std::move(col));
                break;
            }
            }
            cols.push_back(owned.back().get());
        }
        std::vector<std::unique_ptr<Statement>> stack;
        for (uint32_t i = 0; i < n_prog; ++i) {
            const rj_pred_t& e = prog[i];
            if (e.kind == 0) {
                Literal lit = std::monostate{};
                if (e.lit_type == RJ_INT64) lit = e.rhs_i;
                else if (e.lit_type == RJ_FP64) lit = e.rhs_d;
                else if (e.lit_type == RJ_VARCHAR) lit = std::string(e.rhs_s ? e.rhs_s : "", e.rhs_s_len);
                stack.push_back(std::make_unique<Comparison>(e.column, static_cast<Comparison::Op>(e.op), std::move(lit)));
            } else {
                if (e.op == RJ_LOGIC_NOT) {
                    if (stack.empty()) throw std::runtime_error("malformed program");
                    auto c = std::move(stack.back());
                    stack.pop_back();
                    stack.push_back(LogicalOperation::makeNot(std::move(c)));
                } else {
                    if (stack.size() < 2) throw std::runtime_error("malformed program");
                    auto r = std::move(stack.back());
                    stack.pop_back();
                    auto l = std::move(stack.back());
                    stack.pop_back();
                    stack.push_back(e.op == RJ_LOGIC_AND ? LogicalOperation::makeAnd(std::move(l), std::move(r))
                                                         : LogicalOperation::makeOr(std::move(l), std::move(r)));
                }
            }
        }
        if (stack.size() != 1) throw std::runtime_error("malformed program");
        const std::vector<uint8_t> bits = stack.back()->eval(cols);
        for (uint64_t i = 0; i < t->num_rows; ++i) selected[i] = (bits[i / 8] >> (i % 8)) & 1u;
        return 0;
    } catch (const std::exception& e) {
        set_err(err, errlen, e.what());
        return 1;
    }
}

} // extern "C"
