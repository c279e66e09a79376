/*
 * oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Shared declarations of the two CPU checkers of the execute() path:
 *   - oracle/rj_oracle.c        : a plain-C restatement of the reference's algorithm ("port"),
 *   - oracle/ref_shim.cpp       : a C-ABI wrapper around the UNMODIFIED reference sources, compiled
 *                                 where they lie under /root/reference into oracle/_ref/ ("reference").
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * these libraries; the product (radix-join_b200/) never does.
 *
 * Both consume the flattened plan of include/rj_b200.h and return host pages, so a test can hand the
 * same rj_plan_t to the CUDA engine, to the port and to the reference.
 */
#ifndef RJ_ORACLE_H
#define RJ_ORACLE_H

#include "../include/rj_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* One column of decoded cells (the columnar equivalent of the reference's vector<vector<Data>>,
 * include/statement.h:13).  valid[i] = 0 is std::monostate (NULL). */
typedef struct orc_cells_t {
    int32_t         type;    /* rj_data_type */
    uint32_t        reserved;
    uint64_t        rows;
    const uint8_t*  valid;   /* byte per row */
    const void*     values;  /* int32[rows] | int64[rows] | double[rows]; unused for VARCHAR */
    const uint64_t* str_off; /* VARCHAR: rows+1 offsets into chars */
    const char*     chars;
} orc_cells_t;

/* A host result: typed columns of contiguous 8 KB pages. */
typedef struct orc_result orc_result;

/* ---- plain-C restatement (rj_oracle.c) ---- */
int      orc_execute(const rj_plan_t* plan, int n_threads, orc_result** out, char* err, size_t errlen);
int      orc_encode(const orc_cells_t* cols, uint32_t n_cols, uint64_t n_rows, orc_result** out,
                    char* err, size_t errlen);
uint64_t orc_result_num_rows(const orc_result* r);
uint32_t orc_result_num_columns(const orc_result* r);
int32_t  orc_result_column_type(const orc_result* r, uint32_t col);
uint64_t orc_result_column_pages(const orc_result* r, uint32_t col);
const void* orc_result_column_data(const orc_result* r, uint32_t col);
void     orc_result_free(orc_result* r);
/* decode one column; two-call protocol: sizes first, then fill caller buffers */
int orc_decode_sizes(const rj_column_t* col, uint64_t num_rows, uint64_t* n_chars, char* err,
                     size_t errlen);
int orc_decode_fill(const rj_column_t* col, uint64_t num_rows, uint8_t* valid, void* values,
                    uint64_t* str_off, char* chars, char* err, size_t errlen);
/* the reference's hash (src/execute.cpp:16-41) and bucket count (:85-92), exposed for tests */
uint64_t orc_hash_int(int64_t key);
uint64_t orc_hash_bytes(const char* s, uint64_t n);
uint64_t orc_num_buckets(uint64_t build_rows, uint64_t key_bytes);

/* ---- unmodified reference behind the same interface (ref_shim.cpp -> oracle/_ref/libref_oracle.so) */
int      ref_execute(const rj_plan_t* plan, int n_threads, orc_result** out, char* err, size_t errlen);
int      ref_encode(const orc_cells_t* cols, uint32_t n_cols, uint64_t n_rows, orc_result** out,
                    char* err, size_t errlen);
uint64_t ref_result_num_rows(const orc_result* r);
uint32_t ref_result_num_columns(const orc_result* r);
int32_t  ref_result_column_type(const orc_result* r, uint32_t col);
uint64_t ref_result_column_pages(const orc_result* r, uint32_t col);
const void* ref_result_column_data(const orc_result* r, uint32_t col);
void     ref_result_free(orc_result* r);
int ref_decode_sizes(const rj_column_t* col, uint64_t num_rows, uint64_t* n_chars, char* err,
                     size_t errlen);
int ref_decode_fill(const rj_column_t* col, uint64_t num_rows, uint8_t* valid, void* values,
                    uint64_t* str_off, char* chars, char* err, size_t errlen);
/* seconds spent inside the reference's Contest::execute during the last ref_execute on this thread
 * (steady_clock around execute only, as tests/read_sql.cpp:1234-1236 times it) */
double   ref_last_execute_seconds(void);

#ifdef __cplusplus
}
#endif
#endif
