"""Shared test helpers: the reference's golden cases (tests/unit_tests.cpp) restated as data, the gap
cases SURVEY.md section 4 lists, and seeded random plan generators.  Inputs are built with the
ORACLE's page encoder (the reference's own tests build theirs with Table::to_columnar)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import radix_join_b200 as rj  # noqa: E402
from oracle import pyoracle as orc  # noqa: E402

INT32, INT64, FP64, VARCHAR = orc.INT32, orc.INT64, orc.FP64, orc.VARCHAR


def encoder_impl():
    return "port" if orc.available("port") else "ref"


def cells_from_python(type, values):
    """values: python list with None for NULL"""
    if type == VARCHAR:
        return orc.Cells.from_strings(values)
    valid = np.array([v is not None for v in values], dtype=np.uint8)
    dt = {INT32: np.int32, INT64: np.int64, FP64: np.float64}[type]
    arr = np.array([0 if v is None else v for v in values], dtype=dt)
    return orc.Cells(type, valid, values=arr)


def table_from_python(types, rows):
    """rows: list of tuples; -> ColumnarTable"""
    cols = [cells_from_python(t, [r[i] for r in rows]) for i, t in enumerate(types)]
    return table_from_cells(cols, len(rows))


def table_from_cells(cells, n_rows=None):
    if not cells:
        return rj.ColumnarTable(num_rows=n_rows or 0)
    t = orc.encode(cells, impl=encoder_impl())
    return t


def empty_table(types):
    return rj.ColumnarTable(num_rows=0, columns=[rj.Column(t) for t in types])


def rows_of(table):
    """decode a ColumnarTable into a sorted list of python tuples (NULL = None, strings = bytes)"""
    cols = [c.to_python() for c in orc.decode_table(table, impl=encoder_impl())]
    rows = list(zip(*cols)) if cols else []
    key = lambda r: tuple((v is None, type(v).__name__, v if v is not None else 0) for v in r)
    return sorted(rows, key=key)


def sort_rows(rows):
    key = lambda r: tuple((v is None, type(v).__name__, v if v is not None else 0) for v in r)
    return sorted(rows, key=key)


# --------------------------------------------------------------------------------------------------
# golden cases: /root/reference/tests/unit_tests.cpp, restated as data
# --------------------------------------------------------------------------------------------------
def _two_scan_join(build_left, t1, t2, scan1, scan2, left_attr, right_attr, outs):
    plan = rj.Plan()
    plan.new_scan_node(0, scan1)
    plan.new_scan_node(1, scan2)
    plan.new_join_node(build_left, 0, 1, left_attr, right_attr, outs)
    plan.new_input(t1)
    plan.new_input(t2)
    plan.root = 2
    return plan


def golden_cases():
    """name -> (plan, expected num_rows, expected column types, expected sorted rows)"""
    I = INT32
    cases = {}
    s0 = [(0, I)]
    outs2 = [(0, I), (1, I)]

    # "Empty join" unit_tests.cpp:10-28 -- page-less typed inputs
    cases["empty_join"] = (_two_scan_join(True, empty_table([I]), empty_table([I]), s0, s0, 0, 0, outs2),
                           0, [I, I], [])
    # "One line join" :30-57
    t = table_from_python([I], [(1,)])
    cases["one_line_join"] = (_two_scan_join(True, t, t, s0, s0, 0, 0, outs2), 1, [I, I], [(1, 1)])
    # "Simple join" :59-91
    t = table_from_python([I], [(1,), (2,), (3,)])
    cases["simple_join"] = (_two_scan_join(True, t, t, s0, s0, 0, 0, outs2), 3, [I, I],
                            [(1, 1), (2, 2), (3, 3)])
    # "Empty Result" :93-123
    t1 = table_from_python([I], [(1,), (2,), (3,)])
    t2 = table_from_python([I], [(4,), (5,), (6,)])
    cases["empty_result"] = (_two_scan_join(True, t1, t2, s0, s0, 0, 0, outs2), 0, [I, I], [])
    # "Multiple same keys" :125-161
    t = table_from_python([I], [(1,), (1,), (2,), (3,)])
    six = [(1, 1)] * 4 + [(2, 2), (3, 3)]
    cases["multiple_same_keys"] = (_two_scan_join(True, t, t, s0, s0, 0, 0, outs2), 6, [I, I], six)
    # "NULL keys" :163-200
    t = table_from_python([I], [(1,), (1,), (None,), (2,), (3,)])
    cases["null_keys"] = (_two_scan_join(True, t, t, s0, s0, 0, 0, outs2), 6, [I, I], six)
    # "Multiple columns" :202-241 and "Build on right" :243-282
    t = table_from_python([I, VARCHAR],
                          [(1, "xxx"), (1, "yyy"), (None, "zzz"), (2, "uuu"), (3, "vvv")])
    exp = [(1, 1, b"xxx")] * 2 + [(1, 1, b"yyy")] * 2 + [(2, 2, b"uuu"), (3, 3, b"vvv")]
    outs3 = [(0, I), (2, I), (1, VARCHAR)]
    for name, bl in (("multiple_columns", True), ("build_on_right", False)):
        cases[name] = (_two_scan_join(bl, t, t, s0, [(1, VARCHAR), (0, I)], 0, 1, outs3), 6,
                       [I, I, VARCHAR], sort_rows(exp))
    return cases


# --------------------------------------------------------------------------------------------------
# seeded random inputs
# --------------------------------------------------------------------------------------------------
def random_cells(rng, type, n, null_frac=0.1, key_range=None, max_len=40, long_frac=0.0):
    valid = (rng.random(n) >= null_frac).astype(np.uint8) if null_frac > 0 else np.ones(n, np.uint8)
    if type == INT32:
        hi = key_range if key_range else 2**31 - 1
        lo = 0 if key_range else -2**31
        return orc.Cells(type, valid, values=rng.integers(lo, hi, n, dtype=np.int64).astype(np.int32))
    if type == INT64:
        if key_range:
            v = rng.integers(0, key_range, n, dtype=np.int64) * np.int64(0x100000001)
        else:
            v = rng.integers(-2**63, 2**63 - 1, n, dtype=np.int64)
        return orc.Cells(type, valid, values=v)
    if type == FP64:
        bits = rng.integers(0, 2**63 - 1, n, dtype=np.int64).view(np.uint64)
        bits ^= rng.integers(0, 2, n, dtype=np.uint64) << np.uint64(63)
        v = bits.view(np.float64).copy()
        v[~np.isfinite(v)] = -0.0
        return orc.Cells(type, valid, values=v)
    # VARCHAR
    if key_range:
        lens = np.full(n, 0, dtype=np.int64)
        ids = rng.integers(0, key_range, n)
        strs = [b"k%d" % i + b"x" * (int(i) % 7) for i in ids]
    else:
        lens = rng.integers(0, max_len + 1, n)
        strs = [bytes(rng.integers(32, 127, int(l), dtype=np.uint8)) for l in lens]
        if long_frac > 0:
            for i in np.nonzero(rng.random(n) < long_frac)[0]:
                ln = int(rng.choice([8185, 8186, 8188, 8189, 8188 * 2, 8188 * 2 + 1, 20000]))
                strs[i] = bytes(rng.integers(48, 122, ln, dtype=np.uint8))
    strs = [s if v else None for s, v in zip(strs, valid)]
    return orc.Cells.from_strings(strs)


def random_table(rng, types, n, key_cols=(), key_range=None, null_frac=0.1, key_null_frac=0.05,
                 max_len=40, long_frac=0.0):
    cells = []
    for i, t in enumerate(types):
        if i in key_cols:
            cells.append(random_cells(rng, t, n, key_null_frac, key_range))
        else:
            cells.append(random_cells(rng, t, n, null_frac, None, max_len, long_frac))
    return table_from_cells(cells, n), cells


def single_join_plan(t_left, t_right, left_types, right_types, left_key, right_key, build_left,
                     out_cols=None):
    """2 x Scan -> Join over all columns of both tables (the shape of unit_tests.cpp:12-14)"""
    plan = rj.Plan()
    plan.new_scan_node(0, [(i, t) for i, t in enumerate(left_types)])
    plan.new_scan_node(1, [(i, t) for i, t in enumerate(right_types)])
    all_types = list(left_types) + list(right_types)
    if out_cols is None:
        out_cols = list(range(len(all_types)))
    plan.new_join_node(build_left, 0, 1, left_key, right_key, [(c, all_types[c]) for c in out_cols])
    plan.new_input(t_left)
    plan.new_input(t_right)
    plan.root = 2
    return plan


# --------------------------------------------------------------------------------------------------
# numpy mirror of the engine's radix hash (radix-join_b200/csrc/rj_common.cuh: hash_key) -- used by
# the stage tests to predict which radix digit a key falls into
# --------------------------------------------------------------------------------------------------
def _fmix32(h):
    h = h.astype(np.uint32).copy()
    with np.errstate(over="ignore"):
        h ^= h >> np.uint32(16)
        h *= np.uint32(0x85EBCA6B)
        h ^= h >> np.uint32(13)
        h *= np.uint32(0xC2B2AE35)
        h ^= h >> np.uint32(16)
    return h


def hash_keys(keys):
    keys = np.asarray(keys)
    if keys.dtype.itemsize == 4:
        return _fmix32(keys.view(np.uint32))
    k = keys.view(np.uint64)
    lo = (k & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    hi = (k >> np.uint64(32)).astype(np.uint32)
    return _fmix32(lo ^ _fmix32(hi ^ np.uint32(0x9E3779B9)))
