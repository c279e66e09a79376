/* Plain-C use of the drop-in boundary (include/rj_b200.h): the plan of the reference's unit test
 * tests/unit_tests.cpp:12-14 -- two scans and one join -- built from flat structs, executed once with
 * rj_execute + rj_result_fetch and once with rj_execute_streamed.
 *
 *   gcc -std=c11 -Iinclude examples/c_abi_join.c -Lradix-join_b200 -lrj_b200 -Wl,-rpath,$PWD/radix-join_b200
 *
 * The caller provides the 8 KB pages (here: one INT32 column per table, contiguous in host memory). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rj_b200.h"

struct sink_state {
    void*    pages[8];
    uint64_t n_pages[8];
};

/* rj_execute_streamed asks for a host buffer every time result pages of a column are ready */
static void* take_pages(void* user, uint32_t column, int32_t type, uint64_t n_pages) {
    struct sink_state* st = (struct sink_state*)user;
    (void)type;
    if (column >= 8 || st->pages[column]) return NULL; /* this example expects one window */
    st->pages[column] = malloc(n_pages * RJ_PAGE_SIZE);
    st->n_pages[column] = n_pages;
    return st->pages[column];
}

int join_int32_columns(const void* build_pages, uint64_t build_n_pages, uint64_t build_rows,
                       const void* probe_pages, uint64_t probe_n_pages, uint64_t probe_rows) {
    rj_ctx* ctx = NULL;
    if (rj_ctx_create(0, &ctx)) { /* Contest::build_context */
        fprintf(stderr, "no engine: %s\n", rj_last_error(NULL));
        return 1;
    }

    rj_column_t cols[2];
    memset(cols, 0, sizeof cols);
    cols[0].type = RJ_INT32; cols[0].n_pages = build_n_pages; cols[0].contiguous = build_pages;
    cols[1].type = RJ_INT32; cols[1].n_pages = probe_n_pages; cols[1].contiguous = probe_pages;
    rj_table_t tables[2];
    memset(tables, 0, sizeof tables);
    tables[0].num_rows = build_rows; tables[0].n_columns = 1; tables[0].columns = &cols[0];
    tables[1].num_rows = probe_rows; tables[1].n_columns = 1; tables[1].columns = &cols[1];

    /* Scan(0){(0,INT32)}, Scan(1){(0,INT32)}, Join(build_left, 0, 1, 0, 0){(0,INT32),(1,INT32)} */
    rj_attr_t scan_attr = {0, RJ_INT32, 0};
    rj_attr_t join_attrs[2] = {{0, RJ_INT32, 0}, {1, RJ_INT32, 0}};
    rj_node_t nodes[3];
    memset(nodes, 0, sizeof nodes);
    nodes[0].base_table_id = 0; nodes[0].n_output_attrs = 1; nodes[0].output_attrs = &scan_attr;
    nodes[1].base_table_id = 1; nodes[1].n_output_attrs = 1; nodes[1].output_attrs = &scan_attr;
    nodes[2].is_join = 1; nodes[2].build_left = 1; nodes[2].left = 0; nodes[2].right = 1;
    nodes[2].left_attr = 0; nodes[2].right_attr = 0; nodes[2].n_output_attrs = 2; nodes[2].output_attrs = join_attrs;
    rj_plan_t plan;
    memset(&plan, 0, sizeof plan);
    plan.nodes = nodes; plan.n_nodes = 3; plan.inputs = tables; plan.n_inputs = 2; plan.root = 2;

    /* Contest::execute: upload, decode, join, encode; then copy the result pages out */
    rj_result* res = NULL;
    if (rj_execute(ctx, &plan, &res)) {
        fprintf(stderr, "execute: %s\n", rj_last_error(ctx));
        rj_ctx_destroy(ctx);
        return 1;
    }
    printf("%llu rows, %u columns\n", (unsigned long long)rj_result_num_rows(res), rj_result_num_columns(res));
    for (uint32_t c = 0; c < rj_result_num_columns(res); ++c) {
        const uint64_t n = rj_result_column_pages(res, c);
        void* out = malloc((n ? n : 1) * RJ_PAGE_SIZE);
        if (rj_result_fetch(ctx, res, c, NULL, out)) fprintf(stderr, "fetch: %s\n", rj_last_error(ctx));
        free(out);
    }
    rj_result_free(ctx, res);

    /* the same plan, streamed: upload, kernels and download overlap for large inputs */
    struct sink_state st;
    memset(&st, 0, sizeof st);
    uint64_t rows = 0;
    if (rj_execute_streamed(ctx, &plan, 0, take_pages, &st, &rows)) fprintf(stderr, "streamed: %s\n", rj_last_error(ctx));
    printf("streamed: %llu rows\n", (unsigned long long)rows);
    for (int c = 0; c < 8; ++c) free(st.pages[c]);

    rj_ctx_destroy(ctx); /* Contest::destroy_context */
    return 0;
}
