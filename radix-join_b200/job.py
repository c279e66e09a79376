"""JOB-shaped plans on synthetic IMDB-shaped data (BASELINE.json configs 3-5).

Harness-side code: the role of the reference's `tests/read_sql.cpp` (plan construction) and
`Table::from_csv` + `ColumnInserter` (input pages), restated without hsql / DuckDB / the IMDB
dataset, none of which exist offline:

  * `build_plan` follows `load_join_pipeline` (tests/read_sql.cpp:861-1141): the Hash child is the
    build side (:943-953), a join's children are asked for the attributes required above plus their
    own join key (:981-1008), `output_attrs` index the concatenation of the children's outputs
    (:1034-1055), a scan's `output_attrs` index ALL columns of its base table (:1109-1135).  The join
    columns come from the plan's `Hash Cond` (the harness derives an equivalent pair from the SQL).
  * `make_inputs` generates one filtered table per scan: post-filter
    cardinalities from plans.json, PK `id` = a sorted sample of 1..n, FK columns Zipf(0.5) over the
    referenced PK range, NULLs where job/schema.sql allows them, VARCHAR lengths by column name,
    optional long strings (0xffff/0xfffe page chains) in root-output VARCHAR columns.
  * `write_fixed_pages` / `write_varchar_pages` are numpy page writers with the semantics of
    ColumnInserter (include/plan.h:151-335).

The workload definition (tree shapes, join columns, estimates, schema) is derived data:
`radix-join_b200/job/job_workload.json`, produced by tools/extract_job_workload.py.
"""
import json
import os
import zlib

import numpy as np

from . import pagewriter
from .plan import Column, ColumnarTable, DataType, Plan

_HERE = os.path.dirname(os.path.abspath(__file__))
PAGE = 8192
_TYPES = {"INT32": DataType.INT32, "INT64": DataType.INT64, "FP64": DataType.FP64, "VARCHAR": DataType.VARCHAR}

_workload = None


def workload():
    global _workload
    if _workload is None:
        with open(os.path.join(_HERE, "job", "job_workload.json")) as f:
            _workload = json.load(f)
    return _workload


# --------------------------------------------------------------------------------------------------
# plan construction
# --------------------------------------------------------------------------------------------------
def _aliases_under(node):
    if "scan" in node:
        return {node["scan"]}
    return _aliases_under(node["join"][0]) | _aliases_under(node["join"][1])


def scans_of(node):
    if "scan" in node:
        return [node]
    return scans_of(node["join"][0]) + scans_of(node["join"][1])


def build_plan(query_name, tables):
    """tables: alias -> ColumnarTable (all columns of the base table, filtered rows).
    Returns (Plan, [(alias, column, DataType)] of the root)."""
    w = workload()
    q = w["queries"][query_name]
    schema = w["schema"]
    plan = Plan()

    def col_type(alias, col):
        for name, t in schema[q["aliases"][alias]]:
            if name == col:
                return _TYPES[t]
        raise KeyError(f"{alias}.{col}")

    def recurse(node, required):
        if "scan" in node:
            alias = node["scan"]
            cols = schema[q["aliases"][alias]]
            names = [c for c, _ in cols]
            input_id = plan.new_input(tables[alias])
            attrs, out = [], []
            for a, c in required:
                assert a == alias, (a, alias)
                attrs.append((names.index(c), _TYPES[cols[names.index(c)][1]]))
                out.append((a, c, attrs[-1][1]))
            return plan.new_scan_node(input_id, attrs), out
        left_node, right_node = node["join"]
        left_aliases = _aliases_under(left_node)
        (a1, c1), (a2, c2) = node["cond"]
        lkey, rkey = ((a1, c1), (a2, c2)) if a1 in left_aliases else ((a2, c2), (a1, c1))
        left_req = [r for r in required if r[0] in left_aliases]
        right_req = [r for r in required if r[0] not in left_aliases]
        if tuple(lkey) not in [tuple(r) for r in left_req]:
            left_req.append(lkey)
        if tuple(rkey) not in [tuple(r) for r in right_req]:
            right_req.append(rkey)
        left, lcols = recurse(left_node, left_req)
        right, rcols = recurse(right_node, right_req)
        left_attr = [(a, c) for a, c, _ in lcols].index(tuple(lkey))
        right_attr = [(a, c) for a, c, _ in rcols].index(tuple(rkey))
        both = lcols + rcols
        keys = [(a, c) for a, c, _ in both]
        attrs, out = [], []
        for a, c in required:
            i = keys.index((a, c))
            attrs.append((i, both[i][2]))
            out.append(both[i])
        return plan.new_join_node(node["build_left"], left, right, left_attr, right_attr, attrs), out

    plan.root, root_cols = recurse(q["tree"], [tuple(o) for o in q["outputs"]])
    return plan, root_cols


def needed_columns(query_name):
    """alias -> set of columns any node of the plan reads (join keys + root outputs)"""
    q = workload()["queries"][query_name]
    need = {}
    for a, c in q["outputs"]:
        need.setdefault(a, set()).add(c)

    def walk(node):
        if "join" in node:
            for a, c in node["cond"]:
                need.setdefault(a, set()).add(c)
            walk(node["join"][0])
            walk(node["join"][1])

    walk(q["tree"])
    return need


# --------------------------------------------------------------------------------------------------
# which PK does a join column range over?  union-find over every Hash Cond of the workload
# --------------------------------------------------------------------------------------------------
_domains = None


def column_domains():
    """(table, column) -> table whose `id` range the column draws from"""
    global _domains
    if _domains is not None:
        return _domains
    w = workload()
    parent = {}

    def find(x):
        parent.setdefault(x, x)
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    def walk(node, aliases):
        if "join" in node:
            (a1, c1), (a2, c2) = node["cond"]
            parent[find((aliases[a1], c1))] = find((aliases[a2], c2))
            walk(node["join"][0], aliases)
            walk(node["join"][1], aliases)

    for q in w["queries"].values():
        walk(q["tree"], q["aliases"])
    groups = {}
    for x in list(parent):
        groups.setdefault(find(x), []).append(x)
    _domains = {}
    for members in groups.values():
        pks = [t for t, c in members if c == "id"]
        # a class without a primary key (never happens in JOB) ranges over its largest member table
        dom = max(pks, key=lambda t: w["table_rows"][t]) if pks else max((t for t, _ in members), key=lambda t: w["table_rows"][t])
        for m in members:
            _domains[m] = dom
    return _domains


# --------------------------------------------------------------------------------------------------
# page writers (ColumnInserter semantics, vectorised)
# --------------------------------------------------------------------------------------------------
def write_fixed_pages(values, valid, dtype):
    """values: int32 / int64 / float64 array, valid: bool array or None -> (n_pages, 8192) uint8.
    Fixed 1984 / 1007 rows per page (any packing that decodes is legal)."""
    n = len(values)
    w = 4 if dtype == DataType.INT32 else 8
    rpp = 1984 if w == 4 else 1007
    n_pages = (n + rpp - 1) // rpp
    pages = np.zeros((n_pages, PAGE), dtype=np.uint8)
    if n == 0:
        return pages
    if valid is None:
        valid = np.ones(n, dtype=bool)
    valid = np.asarray(valid, dtype=bool)
    page = np.arange(n) // rpp
    n_r = np.bincount(page, minlength=n_pages).astype(np.uint16)
    n_v = np.bincount(page, weights=valid, minlength=n_pages).astype(np.uint16)
    hdr = pages[:, :4].view(np.uint16)
    hdr[:, 0], hdr[:, 1] = n_r, n_v
    # rank of every non-NULL value inside its page
    cum = np.cumsum(valid) - valid
    first = np.zeros(n_pages, dtype=np.int64)
    first[1:] = np.cumsum(n_v.astype(np.int64))[:-1]
    rank = cum - first[page]
    words = pages.view(np.uint32 if w == 4 else np.uint64)
    vv = np.asarray(values)[valid]
    words[page[valid], 1 + rank[valid]] = vv.view(np.uint32 if w == 4 else np.uint64)
    # bitmaps: the last ceil(n_r / 8) bytes of each page
    padded = np.zeros(n_pages * ((rpp + 7) // 8 * 8), dtype=bool).reshape(n_pages, -1)
    padded[page, np.arange(n) - page * rpp] = valid
    bits = np.packbits(padded, axis=1, bitorder="little")
    for nb in np.unique((n_r.astype(np.int64) + 7) // 8):
        sel = (n_r.astype(np.int64) + 7) // 8 == nb
        pages[sel, PAGE - nb:] = bits[sel, :nb]
    return pages


def write_varchar_pages(lengths, chars, valid):
    """lengths[n] (bytes per row, 0 for NULL), chars (flat uint8, rows back to back), valid[n] bool.
    Greedy packing as ColumnInserter<std::string>::insert (include/plan.h:301-320); strings longer than
    8185 bytes become 0xffff/0xfffe page chains (:256-273)."""
    n = len(lengths)
    lengths = np.asarray(lengths, dtype=np.int64)
    valid = np.asarray(valid, dtype=bool)
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lengths, out=off[1:])
    is_long = valid & (lengths > PAGE - 7)
    # bits used on a page by rows [s, e): 8 * sum(2 + len) over valid rows + (e - s) for the bitmap
    wbits = np.where(valid, 8 * (2 + lengths), 0) + 1
    g = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(wbits, out=g[1:])
    cap = 8 * (PAGE - 4) - 7
    longs = np.nonzero(is_long)[0]
    out = []
    s, li = 0, 0
    while s < n:
        nxt = longs[li] if li < len(longs) else n
        if s == nxt:
            # long string: chain of <= 8188-char pages
            data = chars[off[s]:off[s + 1]]
            for k, a in enumerate(range(0, len(data), PAGE - 4)):
                page = np.zeros(PAGE, dtype=np.uint8)
                piece = data[a:a + PAGE - 4]
                page[:4].view(np.uint16)[:] = (0xFFFF if k == 0 else 0xFFFE, len(piece))
                page[4:4 + len(piece)] = piece
                out.append(page)
            s += 1
            li += 1
            continue
        e = int(np.searchsorted(g, g[s] + cap, side="right")) - 1
        e = max(s + 1, min(e, nxt, s + 65000))
        page = np.zeros(PAGE, dtype=np.uint8)
        v = valid[s:e]
        n_v = int(v.sum())
        page[:4].view(np.uint16)[:] = (e - s, n_v)
        ends = np.cumsum(lengths[s:e][v]).astype(np.uint16)
        page[4:4 + 2 * n_v] = ends.view(np.uint8)
        body = chars[off[s]:off[e]]
        page[4 + 2 * n_v:4 + 2 * n_v + len(body)] = body
        bm = np.packbits(v, bitorder="little")
        page[PAGE - len(bm):] = bm
        out.append(page)
        s = e
    return np.stack(out) if out else np.zeros((0, PAGE), dtype=np.uint8)


# --------------------------------------------------------------------------------------------------
# synthetic IMDB-shaped inputs
# --------------------------------------------------------------------------------------------------
def _zipf_ids(rng, n, domain, theta=0.5):
    """n draws from 1..domain, P(rank) ~ rank^-theta (continuous inverse-CDF approximation)"""
    u = rng.random(n)
    r = np.floor(((domain ** (1 - theta) - 1) * u + 1) ** (1 / (1 - theta))).astype(np.int64)
    r = np.clip(r, 1, domain)
    return _rank_to_id(r, domain)


def _rank_to_id(r, domain):
    """popularity rank -> id through a fixed bijection of 1..domain (2654435761 is prime), so that the
    popular ids are spread over the whole range"""
    return ((r * 2654435761) % domain + 1).astype(np.int32)


def _string_lengths(rng, n, col):
    if col == "md5sum":
        return np.full(n, 32, dtype=np.int64)
    if "pcode" in col or col in ("imdb_index", "phonetic_code", "gender"):
        return rng.integers(0, 6, n)
    if col in ("title", "name", "kind", "role", "link", "keyword", "country_code"):
        return rng.integers(5, 61, n)
    return rng.integers(0, 201, n)  # note, info, ...


_SMALL = 10000  # dimension tables (info_type, kind_type, ...) keep their size whatever the scale


def _scaled(rows, scale, min_rows=1):
    return rows if rows <= _SMALL else max(min_rows, int(round(rows * scale)))


def make_inputs(query_name, scale=1.0, seed=0, long_strings=False, min_rows=1, ctx=None):
    """alias -> ColumnarTable for every scan of the query.  Only the columns the plan reads carry
    pages; the others are typed, page-less placeholders (the engine never touches them).
    With an engine context the pages are written on the device (radix_join_b200.pagewriter) instead of by
    the numpy loops below: same rows, possibly another (equally legal) page packing."""
    w = workload()
    q = w["queries"][query_name]
    need = needed_columns(query_name)
    domains = column_domains()
    outputs = {tuple(o) for o in q["outputs"]}
    tables = {}
    for scan in scans_of(q["tree"]):
        alias, table = scan["scan"], scan["table"]
        rng = np.random.default_rng(zlib.crc32(f"{query_name}/{alias}/{seed}".encode()))
        n_base = _scaled(w["table_rows"][table], scale, min_rows)
        n = max(min_rows, min(n_base, _scaled(scan["rows"], scale, min_rows) if w["table_rows"][table] > _SMALL else scan["rows"])) \
            if scan["filtered"] else n_base
        cols = []
        for col, tname in w["schema"][table]:
            dt = _TYPES[tname]
            if col not in need.get(alias, ()):
                cols.append(Column(dt))
                continue
            nullable = col not in w["not_null"].get(table, [])
            valid = rng.random(n) >= 0.3 if nullable else None
            if dt == DataType.INT32:
                if col == "id":
                    # a filtered scan keeps the n most popular entities (the ids the Zipf-distributed
                    # foreign keys hit most), which stands in for the correlation real filters have
                    vals = np.sort(_rank_to_id(np.arange(1, n + 1, dtype=np.int64), n_base)) if n < n_base \
                        else np.arange(1, n + 1, dtype=np.int32)
                elif (table, col) in domains:
                    dom = _scaled(w["table_rows"][domains[(table, col)]], scale, min_rows)
                    vals = _zipf_ids(rng, n, dom)
                else:
                    vals = rng.integers(1880, 2020, n).astype(np.int32)
                cols.append(Column(dt, pagewriter.fixed_pages(ctx, vals, valid, dt) if ctx is not None else write_fixed_pages(vals, valid, dt)))
            else:
                lens = _string_lengths(rng, n, col)
                if long_strings and (alias, col) in outputs and n > 0:
                    pick = rng.random(n) < 1e-4
                    pick[rng.integers(0, n)] = True  # at least one chain per column
                    lens = np.where(pick, rng.integers(8186, 40001, n), lens)
                if valid is not None:
                    lens = np.where(valid, lens, 0)
                chars = rng.integers(97, 123, int(lens.sum()), dtype=np.uint8)
                if ctx is not None:
                    cols.append(Column(dt, pagewriter.varchar_pages(ctx, lens, chars, valid)))
                else:
                    cols.append(Column(dt, write_varchar_pages(lens, chars, valid if valid is not None else np.ones(n, bool))))
        tables[alias] = ColumnarTable(num_rows=n, columns=cols)
    return tables


def make_job(query_name, scale=1.0, seed=0, long_strings=False, ctx=None):
    """-> (Plan, root columns, total scan rows) for one JOB query on synthetic inputs"""
    tables = make_inputs(query_name, scale, seed, long_strings, ctx=ctx)
    plan, root_cols = build_plan(query_name, tables)
    return plan, root_cols, sum(t.num_rows for t in tables.values())
