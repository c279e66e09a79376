"""Generate tests/golden/unit_cases.json by running the UNMODIFIED reference (oracle/_ref) on the
inputs of its own unit tests (tests/unit_tests.cpp) plus the gap cases of SURVEY.md section 4.
Run in the CPU container (needs /root/reference):  python tests/golden/make_golden.py
The fixture holds the input pages (zlib + base64) and the reference's sorted output rows, so GPU parity
tests do not need the reference at run time."""
import base64
import json
import zlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
import helpers as H  # noqa: E402
from helpers import orc, rj  # noqa: E402


def dump_plan(plan):
    return {
        "root": plan.root,
        "nodes": [
            ({"join": [n.data.build_left, n.data.left, n.data.right, n.data.left_attr, n.data.right_attr]}
             if isinstance(n.data, rj.JoinNode) else {"scan": n.data.base_table_id})
            | {"out": [[i, int(t)] for i, t in n.output_attrs]}
            for n in plan.nodes
        ],
        "inputs": [
            {"num_rows": t.num_rows,
             "columns": [{"type": int(c.type), "pages": base64.b64encode(zlib.compress(c.pages.tobytes(), 9)).decode()}
                         for c in t.columns]}
            for t in plan.inputs
        ],
    }


def gap_cases():
    """cases the reference's tests do not pin (SURVEY section 4 'gaps'); expectations come from
    running the reference itself"""
    rng = np.random.default_rng(2025)
    I, L, F, V = H.INT32, H.INT64, H.FP64, H.VARCHAR
    cases = {}
    # INT64 key beyond 2^32, FP64 (-0.0, NULL) and VARCHAR ("", NULL) payloads
    tl = H.table_from_python([L, F], [(2**40, -0.0), (2**40, None), (-5, 1.5), (None, 2.5), (7, 3.5)])
    tr = H.table_from_python([L, V], [(2**40, ""), (-5, None), (-5, "neg"), (None, "null"), (8, "x")])
    cases["int64_key_fp64_varchar_payload"] = H.single_join_plan(tl, tr, [L, F], [L, V], 0, 0, True)
    # VARCHAR key
    tl = H.table_from_python([V, I], [("a", 1), ("b", 2), ("", 3), (None, 4), ("a", 5)])
    tr = H.table_from_python([V], [("a",), ("",), (None,), ("c",)])
    cases["varchar_key"] = H.single_join_plan(tl, tr, [V, I], [V], 0, 0, False)
    # long strings through a join (8185 / 8186 / 20000 chars)
    tl = H.table_from_python([I, V], [(1, "s" * 8185), (2, "t" * 8186), (3, "u" * 20000), (4, None)])
    tr = H.table_from_python([I], [(1,), (2,), (3,), (3,), (4,)])
    cases["long_strings"] = H.single_join_plan(tl, tr, [I, V], [I], 0, 0, True)
    # many NULLs, more than one page per column, duplicate output attrs
    tl, _ = H.random_table(rng, [I, L], 2500, key_cols=(0,), key_range=1500, null_frac=0.6)
    tr, _ = H.random_table(rng, [I, F], 3000, key_cols=(0,), key_range=1500, null_frac=0.6)
    cases["multi_page_nulls_dup_attrs"] = H.single_join_plan(tl, tr, [I, L], [I, F], 0, 0, True,
                                                             out_cols=[1, 3, 1, 0])
    return cases


def main():
    fixture = {}
    for name, (plan, n_rows, types, rows) in H.golden_cases().items():
        res = orc.execute(plan, impl="ref")
        got = H.rows_of(res)
        assert res.num_rows == n_rows and got == rows, name
        fixture[name] = {"plan": dump_plan(plan), "num_rows": n_rows, "types": types,
                         "rows": [[v.decode() if isinstance(v, bytes) else v for v in r] for r in got]}
    for name, plan in gap_cases().items():
        res = orc.execute(plan, impl="ref")
        got = H.rows_of(res)
        fixture[name] = {"plan": dump_plan(plan), "num_rows": res.num_rows,
                         "types": [int(c.type) for c in res.columns],
                         "rows": [[v.decode() if isinstance(v, bytes) else
                                   (v.hex() if isinstance(v, float) else v) for v in r] for r in got]}
    with open(os.path.join(HERE, "unit_cases.json"), "w") as f:
        json.dump(fixture, f)
    print("wrote", len(fixture), "cases,", os.path.getsize(os.path.join(HERE, "unit_cases.json")), "bytes")


if __name__ == "__main__":
    main()
