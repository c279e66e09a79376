/*
 * rj_oracle.c -- TEST INFRASTRUCTURE ONLY ("port" oracle).
 *
 * A plain-C, columnar restatement of the reference's execute() path:
 *     Contest::execute            /root/reference/src/execute.cpp:316-324
 *     execute_impl / execute_scan / execute_hash_join / hash_join_omp   src/execute.cpp:43-314
 *     Table::from_columnar        src/build_table.cpp:312-436
 *     Table::to_columnar          src/build_table.cpp:456-681
 * Every function cites the lines it follows.  Cells are kept as typed columns instead of
 * vector<vector<variant>>, everything else (bucket count, hash, histogram, scatter of row ids,
 * per-bucket open addressing with duplicate lists, greedy page fill) follows the reference step by
 * step.  Parity is PINNED: tests/test_oracle.py checks this file against the reference's own golden
 * vectors (tests/unit_tests.cpp) and against the unmodified reference compiled into
 * oracle/_ref/libref_oracle.so, on the codec (page for page) and on execute() (multiset of rows).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may load this.
 */
#define _POSIX_C_SOURCE 200809L
#include <omp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

#define PAGE RJ_PAGE_SIZE

/* ------------------------------------------------------------------------------------------------ */
/* small utilities                                                                                  */
/* ------------------------------------------------------------------------------------------------ */

typedef struct {
    char msg[256];
    int  failed;
} err_t;

static void fail(err_t* e, const char* fmt, ...) {
    if (e->failed) {
        return;
    }
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(e->msg, sizeof e->msg, fmt, ap);
    va_end(ap);
    e->failed = 1;
}

static void* xcalloc(size_t n, size_t sz) {
    void* p = calloc(n ? n : 1, sz ? sz : 1);
    if (!p) {
        fprintf(stderr, "rj_oracle: out of memory (%zu x %zu)\n", n, sz);
        abort();
    }
    return p;
}

static void* xrealloc(void* p, size_t sz) {
    void* q = realloc(p, sz ? sz : 1);
    if (!q) {
        fprintf(stderr, "rj_oracle: out of memory (%zu)\n", sz);
        abort();
    }
    return q;
}

static uint16_t rd16(const uint8_t* p) {
    uint16_t v;
    memcpy(&v, p, 2);
    return v;
}

static void wr16(uint8_t* p, uint16_t v) { memcpy(p, &v, 2); }

static size_t type_width(int32_t t) { return t == RJ_INT32 ? 4 : 8; }

/* a decoded column: the cells of one attribute for every row */
typedef struct {
    int32_t   type; /* physical type = the variant alternative of the non-NULL cells */
    uint64_t  rows;
    uint8_t*  valid;
    uint8_t*  values; /* rows * width for fixed types */
    uint64_t* soff;   /* rows + 1 for VARCHAR */
    char*     chars;
} col_t;

typedef struct {
    uint64_t rows;
    uint32_t ncols;
    col_t*   cols;
} rel_t;

static void col_free(col_t* c) {
    free(c->valid);
    free(c->values);
    free(c->soff);
    free(c->chars);
    memset(c, 0, sizeof *c);
}

static void rel_free(rel_t* r) {
    for (uint32_t i = 0; i < r->ncols; ++i) {
        col_free(&r->cols[i]);
    }
    free(r->cols);
    memset(r, 0, sizeof *r);
}

/* out[i] = src[idx[i]]  -- the per-row copies of execute.cpp:291-298 (scan projection, idx == NULL)
 * and execute.cpp:236-242 (join output row assembly) in columnar form */
static void col_gather(const col_t* src, const uint32_t* idx, uint64_t n, col_t* out) {
    out->type  = src->type;
    out->rows  = n;
    out->valid = xcalloc(n, 1);
    if (src->type == RJ_VARCHAR) {
        out->soff    = xcalloc(n + 1, sizeof(uint64_t));
        uint64_t pos = 0;
        for (uint64_t i = 0; i < n; ++i) {
            uint64_t r   = idx ? idx[i] : i;
            out->soff[i] = pos;
            pos += src->soff[r + 1] - src->soff[r];
        }
        out->soff[n] = pos;
        out->chars   = xcalloc(pos, 1);
#pragma omp parallel for schedule(static)
        for (uint64_t i = 0; i < n; ++i) {
            uint64_t r    = idx ? idx[i] : i;
            out->valid[i] = src->valid[r];
            memcpy(out->chars + out->soff[i], src->chars + src->soff[r], src->soff[r + 1] - src->soff[r]);
        }
    } else {
        size_t w    = type_width(src->type);
        out->values = xcalloc(n, w);
#pragma omp parallel for schedule(static)
        for (uint64_t i = 0; i < n; ++i) {
            uint64_t r    = idx ? idx[i] : i;
            out->valid[i] = src->valid[r];
            memcpy(out->values + i * w, src->values + r * w, w);
        }
    }
}

static const uint8_t* page_at(const rj_column_t* c, uint64_t i) {
    if (c->pages) {
        return (const uint8_t*)c->pages[i];
    }
    return (const uint8_t*)c->contiguous + i * (uint64_t)PAGE;
}

/* build_table.cpp:306-310 */
static int get_bitmap(const uint8_t* bitmap, uint32_t idx) { return (bitmap[idx / 8] >> (idx % 8)) & 1; }

/* ------------------------------------------------------------------------------------------------ */
/* Table::from_columnar -- src/build_table.cpp:312-436, one column                                   */
/* ------------------------------------------------------------------------------------------------ */
static void decode_column(const rj_column_t* c, uint64_t num_rows, col_t* out, err_t* e) {
    memset(out, 0, sizeof *out);
    out->type  = c->type;
    out->rows  = num_rows;
    out->valid = xcalloc(num_rows, 1); /* results start as monostate, :314-315 */
    uint64_t row_idx = 0;
    if (c->type != RJ_VARCHAR) {
        /* :325-381 -- header n_r @0, values @4 (INT32) or @8 (INT64/FP64), bitmap = last (n_r+7)/8 bytes */
        size_t w      = type_width(c->type);
        size_t begin  = c->type == RJ_INT32 ? 4 : 8;
        out->values   = xcalloc(num_rows, w);
        for (uint64_t p = 0; p < c->n_pages && !e->failed; ++p) {
            const uint8_t* page     = page_at(c, p);
            uint32_t       n_r      = rd16(page);
            const uint8_t* bitmap   = page + PAGE - (n_r + 7) / 8;
            uint32_t       data_idx = 0;
            for (uint32_t i = 0; i < n_r; ++i) {
                if (get_bitmap(bitmap, i)) {
                    if (row_idx >= num_rows) {
                        fail(e, "row_idx"); /* :334-336 */
                        break;
                    }
                    memcpy(out->values + row_idx * w, page + begin + (size_t)data_idx * w, w);
                    out->valid[row_idx] = 1;
                    ++data_idx;
                    ++row_idx;
                } else {
                    ++row_idx; /* :338-340 (no bound check for NULLs in the reference either) */
                }
            }
        }
        return;
    }
    /* VARCHAR :382-429 */
    uint64_t* len  = xcalloc(num_rows, sizeof(uint64_t));
    size_t    cap  = 1 << 16, pos = 0;
    char*     buf  = xcalloc(cap, 1);
    for (uint64_t p = 0; p < c->n_pages && !e->failed; ++p) {
        const uint8_t* page = page_at(c, p);
        uint32_t       n_r  = rd16(page);
        uint32_t       n_v  = rd16(page + 2);
        if (n_r == 0xffff || n_r == 0xfffe) {
            /* :384-405 -- one (piece of a) long string: n_v chars @4 */
            if (n_r == 0xffff) {
                if (row_idx >= num_rows) {
                    fail(e, "row_idx");
                    break;
                }
                out->valid[row_idx] = 1;
                ++row_idx;
            } else if (row_idx == 0 || row_idx > num_rows || !out->valid[row_idx - 1]) {
                fail(e, "long string page 0xfffe must follows a string"); /* :401-402 */
                break;
            }
            if (pos + n_v > cap) {
                while (pos + n_v > cap) cap *= 2;
                buf = xrealloc(buf, cap);
            }
            memcpy(buf + pos, page + 4, n_v);
            pos += n_v;
            len[row_idx - 1] += n_v;
            continue;
        }
        /* :406-427 -- n_v u16 END offsets @4, chars @4+2*n_v, bitmap at the page end */
        const uint8_t* offsets  = page + 4;
        const uint8_t* data     = page + 4 + 2 * (size_t)n_v;
        const uint8_t* bitmap   = page + PAGE - (n_r + 7) / 8;
        uint32_t       data_idx = 0, prev = 0;
        for (uint32_t i = 0; i < n_r; ++i) {
            if (get_bitmap(bitmap, i)) {
                uint32_t end = rd16(offsets + 2 * (size_t)data_idx++);
                if (row_idx >= num_rows) {
                    fail(e, "row_idx");
                    break;
                }
                uint32_t l = end - prev;
                if (pos + l > cap) {
                    while (pos + l > cap) cap *= 2;
                    buf = xrealloc(buf, cap);
                }
                memcpy(buf + pos, data + prev, l);
                pos += l;
                prev                = end;
                len[row_idx]        = l;
                out->valid[row_idx] = 1;
                ++row_idx;
            } else {
                ++row_idx;
            }
        }
    }
    out->soff    = xcalloc(num_rows + 1, sizeof(uint64_t));
    uint64_t acc = 0;
    for (uint64_t i = 0; i < num_rows; ++i) {
        out->soff[i] = acc;
        acc += len[i];
    }
    out->soff[num_rows] = acc;
    out->chars          = buf;
    free(len);
}

/* ------------------------------------------------------------------------------------------------ */
/* Table::to_columnar -- src/build_table.cpp:456-681, one column                                     */
/* ------------------------------------------------------------------------------------------------ */
typedef struct {
    uint8_t* data;
    uint64_t n_pages, cap_pages;
} pages_t;

static uint8_t* new_page(pages_t* pg) {
    if (pg->n_pages == pg->cap_pages) {
        pg->cap_pages = pg->cap_pages ? pg->cap_pages * 2 : 4;
        pg->data      = xrealloc(pg->data, pg->cap_pages * (size_t)PAGE);
    }
    uint8_t* p = pg->data + pg->n_pages++ * (size_t)PAGE;
    memset(p, 0, PAGE); /* the reference leaves `new Page` bytes indeterminate; the port zeroes them */
    return p;
}

/* `declared` is the column type handed to Table{...}; a non-NULL cell whose alternative differs is
 * silently skipped for the fixed types (:484-501) and throws for VARCHAR (:667-669) */
static void encode_column(const col_t* c, int32_t declared, pages_t* pg, err_t* e) {
    memset(pg, 0, sizeof *pg);
    uint8_t  bitmap[PAGE];
    uint32_t num_rows = 0;
    if (declared != RJ_VARCHAR) {
        size_t   w     = type_width(declared);
        size_t   begin = declared == RJ_INT32 ? 4 : 8;
        uint8_t  data[PAGE];
        uint32_t n_v = 0;
        memset(bitmap, 0, sizeof bitmap);
#define SAVE_FIXED()                                                                   \
    do {                                                                               \
        uint8_t* page = new_page(pg);                                                  \
        wr16(page, (uint16_t)num_rows);                                                \
        wr16(page + 2, (uint16_t)n_v);                                                 \
        memcpy(page + begin, data, (size_t)n_v * w);                                   \
        memcpy(page + PAGE - (num_rows + 7) / 8, bitmap, (num_rows + 7) / 8);          \
        num_rows = 0;                                                                  \
        n_v      = 0;                                                                  \
        memset(bitmap, 0, sizeof bitmap);                                              \
    } while (0)
        for (uint64_t i = 0; i < c->rows; ++i) {
            if (c->valid[i]) {
                if (c->type != declared) {
                    continue; /* no branch of the visitor matches: cell dropped */
                }
                /* :488 / :531 / :574 */
                if (begin + ((size_t)n_v + 1) * w + (num_rows / 8 + 1) > PAGE) {
                    SAVE_FIXED();
                }
                bitmap[num_rows / 8] |= (uint8_t)(1u << (num_rows % 8));
                memcpy(data + (size_t)n_v * w, c->values + i * w, w);
                ++n_v;
                ++num_rows;
            } else {
                /* :495 / :538 / :581 */
                if (begin + (size_t)n_v * w + (num_rows / 8 + 1) > PAGE) {
                    SAVE_FIXED();
                }
                ++num_rows; /* bit stays 0 */
            }
        }
        if (num_rows != 0) {
            SAVE_FIXED();
        }
#undef SAVE_FIXED
        return;
    }
    /* VARCHAR :595-677 */
    uint8_t  data[PAGE];
    uint16_t offsets[PAGE / 2];
    uint32_t n_v = 0, data_size = 0;
    memset(bitmap, 0, sizeof bitmap);
#define SAVE_VARCHAR()                                                                 \
    do {                                                                               \
        uint8_t* page = new_page(pg);                                                  \
        wr16(page, (uint16_t)num_rows);                                                \
        wr16(page + 2, (uint16_t)n_v);                                                 \
        memcpy(page + 4, offsets, (size_t)n_v * 2);                                    \
        memcpy(page + 4 + (size_t)n_v * 2, data, data_size);                           \
        memcpy(page + PAGE - (num_rows + 7) / 8, bitmap, (num_rows + 7) / 8);          \
        num_rows  = 0;                                                                 \
        n_v       = 0;                                                                 \
        data_size = 0;                                                                 \
        memset(bitmap, 0, sizeof bitmap);                                              \
    } while (0)
    for (uint64_t i = 0; i < c->rows; ++i) {
        if (c->valid[i]) {
            if (c->type != RJ_VARCHAR) {
                fail(e, "not string or null"); /* :667-669 */
                break;
            }
            const char* s   = c->chars + c->soff[i];
            uint64_t    len = c->soff[i + 1] - c->soff[i];
            if (len > PAGE - 7) {
                /* :644-648 then save_long_string :603-619 */
                if (num_rows > 0) {
                    SAVE_VARCHAR();
                }
                uint64_t off   = 0;
                int      first = 1;
                while (off < len) {
                    uint8_t* page = new_page(pg);
                    wr16(page, first ? 0xffff : 0xfffe);
                    first         = 0;
                    uint64_t take = len - off < PAGE - 4 ? len - off : PAGE - 4;
                    wr16(page + 2, (uint16_t)take);
                    memcpy(page + 4, s + off, take);
                    off += take;
                }
            } else {
                /* :650-658 */
                if (4 + ((size_t)n_v + 1) * 2 + (data_size + len) + (num_rows / 8 + 1) > PAGE) {
                    SAVE_VARCHAR();
                }
                bitmap[num_rows / 8] |= (uint8_t)(1u << (num_rows % 8));
                memcpy(data + data_size, s, len);
                data_size += (uint32_t)len;
                offsets[n_v++] = (uint16_t)data_size;
                ++num_rows;
            }
        } else {
            /* :660-666 */
            if (4 + (size_t)n_v * 2 + data_size + (num_rows / 8 + 1) > PAGE) {
                SAVE_VARCHAR();
            }
            ++num_rows;
        }
    }
    if (num_rows != 0 && !e->failed) {
        SAVE_VARCHAR();
    }
#undef SAVE_VARCHAR
}

/* ------------------------------------------------------------------------------------------------ */
/* HashUtil<K>::hash -- src/execute.cpp:16-41                                                        */
/* ------------------------------------------------------------------------------------------------ */
uint64_t orc_hash_int(int64_t key) {
    uint64_t k = (uint64_t)key; /* int32 keys are sign-extended by static_cast<uint64_t>, :21 */
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}

uint64_t orc_hash_bytes(const char* s, uint64_t n) {
    uint64_t h = 14695981039346656037ULL; /* :33-38; `char` is signed on this ABI and is
                                             sign-extended by static_cast<size_t>(c) */
    for (uint64_t i = 0; i < n; ++i) {
        h ^= (uint64_t)(int64_t)(signed char)s[i];
        h *= 1099511628211ULL;
    }
    return h;
}

/* src/execute.cpp:85-92 with SPC__LEVEL2_CACHE_SIZE = 1 MiB (include/hardware.h:44) */
uint64_t orc_num_buckets(uint64_t build_rows, uint64_t key_bytes) {
    const uint64_t l2     = 1048576;
    uint64_t       approx = (build_rows * (key_bytes + 4) + l2 - 1) / l2;
    if (approx < 1) approx = 1;
    if (approx > 128) approx = 128;
    uint64_t nb = 1;
    while (nb < approx) nb <<= 1;
    return nb;
}

/* ------------------------------------------------------------------------------------------------ */
/* hash_join_omp<K> -- src/execute.cpp:43-262                                                        */
/* ------------------------------------------------------------------------------------------------ */
typedef struct {
    uint32_t* l;
    uint32_t* r;
    uint64_t  n, cap;
} pairs_t;

static void pairs_push(pairs_t* p, uint32_t l, uint32_t r) {
    if (p->n == p->cap) {
        p->cap = p->cap ? p->cap * 2 : 1024;
        p->l   = xrealloc(p->l, p->cap * sizeof(uint32_t));
        p->r   = xrealloc(p->r, p->cap * sizeof(uint32_t));
    }
    p->l[p->n] = l;
    p->r[p->n] = r;
    ++p->n;
}

static int key_equal(const col_t* a, uint64_t i, const col_t* b, uint64_t j) {
    if (a->type == RJ_VARCHAR) {
        uint64_t la = a->soff[i + 1] - a->soff[i], lb = b->soff[j + 1] - b->soff[j];
        return la == lb && memcmp(a->chars + a->soff[i], b->chars + b->soff[j], la) == 0;
    }
    size_t w = type_width(a->type);
    return memcmp(a->values + i * w, b->values + j * w, w) == 0;
}

/* step 2 (:61-83): a row is valid iff its cell holds exactly KeyType; the hash is computed once
 * here instead of at each of the reference's three call sites (same value) */
static void extract_keys(const col_t* c, int32_t key_type, uint8_t* valid, uint64_t* hash) {
#pragma omp parallel for schedule(static)
    for (uint64_t i = 0; i < c->rows; ++i) {
        valid[i] = (uint8_t)(c->valid[i] && c->type == key_type);
        if (!valid[i]) {
            hash[i] = 0;
        } else if (key_type == RJ_INT32) {
            int32_t v;
            memcpy(&v, c->values + i * 4, 4);
            hash[i] = orc_hash_int(v);
        } else if (key_type == RJ_INT64) {
            int64_t v;
            memcpy(&v, c->values + i * 8, 8);
            hash[i] = orc_hash_int(v);
        } else {
            hash[i] = orc_hash_bytes(c->chars + c->soff[i], c->soff[i + 1] - c->soff[i]);
        }
    }
}

static void hash_join(const rel_t* left, const rel_t* right, int build_left, uint64_t left_attr,
    uint64_t right_attr, int32_t key_type, int n_threads, pairs_t* out, err_t* e) {
    memset(out, 0, sizeof *out);
    if (key_type == RJ_FP64) {
        /* HashUtil<double>::hash calls itself on the reinterpreted bits (:28-31): the reference
         * never returns for an FP64 join key, so there is no behaviour to restate */
        fail(e, "FP64 join key: the reference's HashUtil<double>::hash recurses forever");
        return;
    }
    const rel_t* build     = build_left ? left : right;
    const rel_t* probe     = build_left ? right : left;
    uint64_t     build_col = build_left ? left_attr : right_attr;   /* :55 */
    uint64_t     probe_col = build_left ? right_attr : left_attr;   /* :56 */
    if (build_col >= build->ncols || probe_col >= probe->ncols) {
        fail(e, "join attribute out of range");
        return;
    }
    const col_t* bk = &build->cols[build_col];
    const col_t* pk = &probe->cols[probe_col];
    uint64_t     B = build->rows, P = probe->rows;

    uint8_t*  bvalid = xcalloc(B, 1);
    uint8_t*  pvalid = xcalloc(P, 1);
    uint64_t* bhash  = xcalloc(B, sizeof(uint64_t));
    uint64_t* phash  = xcalloc(P, sizeof(uint64_t));
    extract_keys(bk, key_type, bvalid, bhash);
    extract_keys(pk, key_type, pvalid, phash);

    /* step 3 (:85-92); sizeof(std::string) is 32 with libstdc++ */
    uint64_t key_bytes   = key_type == RJ_INT32 ? 4 : key_type == RJ_INT64 ? 8 : 32;
    uint64_t num_buckets = orc_num_buckets(B, key_bytes);
    uint64_t bucket_mask = num_buckets - 1;

    /* step 4 (:124-132) serial histograms, (:169-184) exclusive prefix + scatter of ROW IDS */
    uint32_t* build_off = xcalloc(num_buckets + 1, sizeof(uint32_t));
    uint32_t* probe_off = xcalloc(num_buckets + 1, sizeof(uint32_t));
    for (uint64_t i = 0; i < B; ++i)
        if (bvalid[i]) build_off[(bhash[i] & bucket_mask) + 1]++;
    for (uint64_t i = 0; i < P; ++i)
        if (pvalid[i]) probe_off[(phash[i] & bucket_mask) + 1]++;
    for (uint64_t b = 0; b < num_buckets; ++b) {
        build_off[b + 1] += build_off[b];
        probe_off[b + 1] += probe_off[b];
    }
    uint32_t* build_buf = xcalloc(B, sizeof(uint32_t));
    uint32_t* probe_buf = xcalloc(P, sizeof(uint32_t));
    uint32_t* bo        = xcalloc(num_buckets + 1, sizeof(uint32_t));
    uint32_t* po        = xcalloc(num_buckets + 1, sizeof(uint32_t));
    memcpy(bo, build_off, (num_buckets + 1) * sizeof(uint32_t));
    memcpy(po, probe_off, (num_buckets + 1) * sizeof(uint32_t));
    for (uint64_t i = 0; i < B; ++i)
        if (bvalid[i]) build_buf[bo[bhash[i] & bucket_mask]++] = (uint32_t)i;
    for (uint64_t i = 0; i < P; ++i)
        if (pvalid[i]) probe_buf[po[phash[i] & bucket_mask]++] = (uint32_t)i;

    /* step 5 (:187-250): buckets in parallel, dynamic schedule, thread-local outputs */
    int nthreads = n_threads > 0 ? n_threads : omp_get_max_threads();
    pairs_t* thread_out = xcalloc((size_t)nthreads, sizeof(pairs_t));
#pragma omp parallel num_threads(nthreads)
    {
        pairs_t* local = &thread_out[omp_get_thread_num()];
#pragma omp for schedule(dynamic, 1)
        for (uint64_t b = 0; b < num_buckets; ++b) {
            uint32_t bs = build_off[b], be = build_off[b + 1];
            uint32_t ps = probe_off[b], pe = probe_off[b + 1];
            uint64_t cnt = be - bs;
            if (cnt == 0 || ps == pe) continue; /* :200 */
            /* :203-208 -- cap = pow2 >= 2*cnt; slot_key -> slot_row (a representative build row),
             * slot_idxs[h] (vector of duplicates) -> singly linked list head/next in insertion order */
            uint64_t cap = 1;
            while (cap < cnt * 2) cap <<= 1;
            uint64_t  mask      = cap - 1;
            uint32_t* slot_row  = xcalloc(cap, sizeof(uint32_t));
            uint8_t*  slot_used = xcalloc(cap, 1);
            uint32_t* head      = xcalloc(cap, sizeof(uint32_t));
            uint32_t* tail      = xcalloc(cap, sizeof(uint32_t));
            uint32_t* next      = xcalloc(cnt, sizeof(uint32_t)); /* indexed by position in bucket */
            /* build :211-223 */
            for (uint32_t idx = bs; idx < be; ++idx) {
                uint32_t row = build_buf[idx];
                uint64_t h   = bhash[row] & mask;
                while (slot_used[h] && !key_equal(bk, slot_row[h], bk, row)) h = (h + 1) & mask;
                uint32_t node = idx - bs + 1; /* 0 = end of list */
                if (!slot_used[h]) {
                    slot_used[h] = 1;
                    slot_row[h]  = row;
                    head[h]      = node;
                } else {
                    next[tail[h] - 1] = node;
                }
                tail[h] = node;
            }
            /* probe :226-248 */
            for (uint32_t idx = ps; idx < pe; ++idx) {
                uint32_t prow = probe_buf[idx];
                uint64_t h    = phash[prow] & mask;
                while (slot_used[h]) {
                    if (key_equal(bk, slot_row[h], pk, prow)) {
                        for (uint32_t node = head[h]; node; node = next[node - 1]) {
                            uint32_t bi = build_buf[bs + node - 1];
                            /* :233-234 */
                            pairs_push(local, build_left ? bi : prow, build_left ? prow : bi);
                        }
                        break;
                    }
                    h = (h + 1) & mask;
                }
            }
            free(slot_row);
            free(slot_used);
            free(head);
            free(tail);
            free(next);
        }
    }
    /* step 6 (:252-259) merge in thread order */
    uint64_t total = 0;
    for (int t = 0; t < nthreads; ++t) total += thread_out[t].n;
    out->l = xcalloc(total, sizeof(uint32_t));
    out->r = xcalloc(total, sizeof(uint32_t));
    out->n = out->cap = total;
    uint64_t pos      = 0;
    for (int t = 0; t < nthreads; ++t) {
        memcpy(out->l + pos, thread_out[t].l, thread_out[t].n * sizeof(uint32_t));
        memcpy(out->r + pos, thread_out[t].r, thread_out[t].n * sizeof(uint32_t));
        pos += thread_out[t].n;
        free(thread_out[t].l);
        free(thread_out[t].r);
    }
    free(thread_out);
    free(bvalid);
    free(pvalid);
    free(bhash);
    free(phash);
    free(build_off);
    free(probe_off);
    free(build_buf);
    free(probe_buf);
    free(bo);
    free(po);
}

/* ------------------------------------------------------------------------------------------------ */
/* execute_impl / execute_scan / execute_hash_join -- src/execute.cpp:266-314                        */
/* ------------------------------------------------------------------------------------------------ */
static void execute_node(const rj_plan_t* plan, uint64_t node_idx, int n_threads, rel_t* out, err_t* e) {
    memset(out, 0, sizeof *out);
    if (node_idx >= plan->n_nodes) {
        fail(e, "node index out of range");
        return;
    }
    const rj_node_t* node = &plan->nodes[node_idx];
    if (!node->is_join) {
        /* execute_scan :284-300 -- from_columnar decodes EVERY column of the input, one column per
         * pool thread (filter_tp has 12, include/inner_column.h:105), then projects output_attrs */
        if (node->base_table_id >= plan->n_inputs) {
            fail(e, "base table out of range");
            return;
        }
        const rj_table_t* in  = &plan->inputs[node->base_table_id];
        col_t*            all = xcalloc(in->n_columns, sizeof(col_t));
        err_t*            errs = xcalloc(in->n_columns, sizeof(err_t));
        int pool = in->n_columns < 12 ? (int)in->n_columns : 12;
#pragma omp parallel for schedule(static, 1) num_threads(pool > 0 ? pool : 1)
        for (uint32_t c = 0; c < in->n_columns; ++c) {
            decode_column(&in->columns[c], in->num_rows, &all[c], &errs[c]);
        }
        for (uint32_t c = 0; c < in->n_columns; ++c) {
            if (errs[c].failed) fail(e, "%s", errs[c].msg);
        }
        out->rows  = in->num_rows;
        out->ncols = node->n_output_attrs;
        out->cols  = xcalloc(out->ncols, sizeof(col_t));
        for (uint32_t a = 0; a < node->n_output_attrs && !e->failed; ++a) {
            uint64_t ci = node->output_attrs[a].index;
            if (ci >= in->n_columns) {
                fail(e, "scan attribute out of range");
                break;
            }
            col_gather(&all[ci], NULL, in->num_rows, &out->cols[a]);
        }
        for (uint32_t c = 0; c < in->n_columns; ++c) col_free(&all[c]);
        free(all);
        free(errs);
        return;
    }
    /* execute_hash_join :266-282: key type = the BUILD side's declared attribute type */
    if (node->left >= plan->n_nodes || node->right >= plan->n_nodes) {
        fail(e, "join child out of range");
        return;
    }
    const rj_node_t* bnode = &plan->nodes[node->build_left ? node->left : node->right];
    uint64_t         battr = node->build_left ? node->left_attr : node->right_attr;
    if (battr >= bnode->n_output_attrs) {
        fail(e, "join attribute out of range");
        return;
    }
    int32_t key_type = bnode->output_attrs[battr].type;
    if (key_type < RJ_INT32 || key_type > RJ_VARCHAR) {
        fail(e, "Unsupported join type"); /* :280 */
        return;
    }
    rel_t left, right;
    execute_node(plan, node->left, n_threads, &left, e);   /* :48 */
    execute_node(plan, node->right, n_threads, &right, e); /* :49 */
    out->ncols = node->n_output_attrs;
    out->cols  = xcalloc(out->ncols, sizeof(col_t));
    if (!e->failed && left.rows != 0 && right.rows != 0) { /* :50 */
        pairs_t pairs;
        hash_join(&left, &right, node->build_left, node->left_attr, node->right_attr, key_type,
            n_threads, &pairs, e);
        if (!e->failed) {
            uint64_t left_w = left.ncols; /* :57 */
            out->rows       = pairs.n;
            for (uint32_t a = 0; a < node->n_output_attrs; ++a) {
                uint64_t ci = node->output_attrs[a].index; /* :238-241 */
                if (ci < left_w) {
                    col_gather(&left.cols[ci], pairs.l, pairs.n, &out->cols[a]);
                } else if (ci - left_w < right.ncols) {
                    col_gather(&right.cols[ci - left_w], pairs.r, pairs.n, &out->cols[a]);
                } else {
                    fail(e, "join output attribute out of range");
                    break;
                }
            }
        }
        free(pairs.l);
        free(pairs.r);
    } else {
        /* `return {}`: zero rows; give the columns their declared types so the root stays typed */
        for (uint32_t a = 0; a < node->n_output_attrs; ++a) {
            out->cols[a].type = node->output_attrs[a].type;
        }
    }
    rel_free(&left);
    rel_free(&right);
}

/* ------------------------------------------------------------------------------------------------ */
/* results                                                                                          */
/* ------------------------------------------------------------------------------------------------ */
struct orc_result {
    uint64_t num_rows;
    uint32_t ncols;
    int32_t* types;
    pages_t* pages;
};

static void copy_err(const err_t* e, char* err, size_t errlen) {
    if (err && errlen) snprintf(err, errlen, "%s", e->msg);
}

void orc_result_free(orc_result* r) {
    if (!r) return;
    for (uint32_t c = 0; c < r->ncols; ++c) free(r->pages[c].data);
    free(r->pages);
    free(r->types);
    free(r);
}

/* Contest::execute -- src/execute.cpp:316-324: run the tree, then Table{rows, root types}.to_columnar() */
int orc_execute(const rj_plan_t* plan, int n_threads, orc_result** out, char* err, size_t errlen) {
    err_t e;
    memset(&e, 0, sizeof e);
    rel_t root;
    execute_node(plan, plan->root, n_threads, &root, &e);
    orc_result* r = xcalloc(1, sizeof *r);
    if (!e.failed) {
        const rj_node_t* rn = &plan->nodes[plan->root];
        r->num_rows         = root.rows;
        r->ncols            = rn->n_output_attrs;
        r->types            = xcalloc(r->ncols, sizeof(int32_t));
        r->pages            = xcalloc(r->ncols, sizeof(pages_t));
        err_t* errs         = xcalloc(r->ncols, sizeof(err_t));
#pragma omp parallel for schedule(dynamic, 1)
        for (uint32_t c = 0; c < r->ncols; ++c) {
            r->types[c] = rn->output_attrs[c].type;
            encode_column(&root.cols[c], r->types[c], &r->pages[c], &errs[c]);
        }
        for (uint32_t c = 0; c < r->ncols; ++c)
            if (errs[c].failed) fail(&e, "%s", errs[c].msg);
        free(errs);
    }
    rel_free(&root);
    if (e.failed) {
        copy_err(&e, err, errlen);
        orc_result_free(r);
        return 1;
    }
    *out = r;
    return 0;
}

static void cells_to_col(const orc_cells_t* in, uint64_t n_rows, col_t* c) {
    /* non-owning view; never freed through col_free */
    c->type   = in->type;
    c->rows   = n_rows;
    c->valid  = (uint8_t*)in->valid;
    c->values = (uint8_t*)in->values;
    c->soff   = (uint64_t*)in->str_off;
    c->chars  = (char*)in->chars;
}

int orc_encode(const orc_cells_t* cols, uint32_t n_cols, uint64_t n_rows, orc_result** out,
    char* err, size_t errlen) {
    err_t e;
    memset(&e, 0, sizeof e);
    orc_result* r = xcalloc(1, sizeof *r);
    r->num_rows   = n_rows;
    r->ncols      = n_cols;
    r->types      = xcalloc(n_cols, sizeof(int32_t));
    r->pages      = xcalloc(n_cols, sizeof(pages_t));
    for (uint32_t c = 0; c < n_cols; ++c) {
        col_t view;
        cells_to_col(&cols[c], n_rows, &view);
        r->types[c] = cols[c].type;
        encode_column(&view, cols[c].type, &r->pages[c], &e);
    }
    if (e.failed) {
        copy_err(&e, err, errlen);
        orc_result_free(r);
        return 1;
    }
    *out = r;
    return 0;
}

uint64_t orc_result_num_rows(const orc_result* r) { return r->num_rows; }
uint32_t orc_result_num_columns(const orc_result* r) { return r->ncols; }
int32_t  orc_result_column_type(const orc_result* r, uint32_t c) { return r->types[c]; }
uint64_t orc_result_column_pages(const orc_result* r, uint32_t c) { return r->pages[c].n_pages; }
const void* orc_result_column_data(const orc_result* r, uint32_t c) { return r->pages[c].data; }

int orc_decode_sizes(const rj_column_t* col, uint64_t num_rows, uint64_t* n_chars, char* err,
    size_t errlen) {
    err_t e;
    memset(&e, 0, sizeof e);
    col_t c;
    decode_column(col, num_rows, &c, &e);
    *n_chars = (c.soff && !e.failed) ? c.soff[num_rows] : 0;
    col_free(&c);
    if (e.failed) {
        copy_err(&e, err, errlen);
        return 1;
    }
    return 0;
}

int orc_decode_fill(const rj_column_t* col, uint64_t num_rows, uint8_t* valid, void* values,
    uint64_t* str_off, char* chars, char* err, size_t errlen) {
    err_t e;
    memset(&e, 0, sizeof e);
    col_t c;
    decode_column(col, num_rows, &c, &e);
    if (!e.failed) {
        memcpy(valid, c.valid, num_rows);
        if (col->type == RJ_VARCHAR) {
            memcpy(str_off, c.soff, (num_rows + 1) * sizeof(uint64_t));
            memcpy(chars, c.chars, c.soff[num_rows]);
        } else {
            memcpy(values, c.values, num_rows * type_width(col->type));
        }
    }
    col_free(&c);
    if (e.failed) {
        copy_err(&e, err, errlen);
        return 1;
    }
    return 0;
}
