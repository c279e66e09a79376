"""JOB-shaped plans (configs 3-5): the plan builder restating load_join_pipeline, the synthetic
IMDB-shaped generator and its numpy page writers (CPU), and engine-vs-oracle parity (GPU)."""
import numpy as np
import pytest

import helpers as H
from helpers import INT32, INT64, FP64, VARCHAR, orc, rj
from radix_join_b200 import job


def test_workload_shape_matches_the_survey():
    w = job.workload()
    assert len(w["queries"]) == 113 and len(w["schema"]) == 21
    n_join = n_scan = 0
    for name in w["queries"]:
        tree = w["queries"][name]["tree"]
        n_scan += len(job.scans_of(tree))

        def joins(n):
            return 0 if "scan" in n else 1 + joins(n["join"][0]) + joins(n["join"][1])
        n_join += joins(tree)
    assert (n_join, n_scan) == (864, 977)
    assert w["table_rows"]["cast_info"] == 36224371 and w["table_rows"]["title"] == 2528244
    dom = job.column_domains()
    assert dom[("movie_companies", "movie_id")] == "title" and dom[("cast_info", "person_id")] == "name"


def test_every_plan_builds_and_is_well_formed():
    w = job.workload()
    for name in w["queries"]:
        tables = {s["scan"]: rj.ColumnarTable(num_rows=0, columns=[rj.Column(job._TYPES[t]) for _, t in w["schema"][s["table"]]])
                  for s in job.scans_of(w["queries"][name]["tree"])}
        plan, root_cols = job.build_plan(name, tables)
        assert len(root_cols) == len(w["queries"][name]["outputs"])
        for n in plan.nodes:
            if isinstance(n.data, rj.JoinNode):
                lw = len(plan.nodes[n.data.left].output_attrs)
                rw = len(plan.nodes[n.data.right].output_attrs)
                assert n.data.left_attr < lw and n.data.right_attr < rw
                assert all(i < lw + rw for i, _ in n.output_attrs)
                # all JOB join keys are INT32 (ANNOUNCEMENTS.md:11)
                assert plan.nodes[n.data.left].output_attrs[n.data.left_attr][1] == rj.DataType.INT32
    # 1a: 4 joins, build side is always the right (Hash) child, root = (mc.note, t.title, t.production_year)
    plan, root_cols = job.build_plan("1a", {s["scan"]: rj.ColumnarTable(0, [rj.Column(job._TYPES[t]) for _, t in w["schema"][s["table"]]])
                                            for s in job.scans_of(w["queries"]["1a"]["tree"])})
    joins = [n for n in plan.nodes if isinstance(n.data, rj.JoinNode)]
    assert len(joins) == 4 and not any(j.data.build_left for j in joins)
    assert [(a, c) for a, c, _ in root_cols] == [("mc", "note"), ("t", "title"), ("t", "production_year")]


@pytest.mark.parametrize("dtype,n,null_frac", [(INT32, 5000, 0.0), (INT32, 5000, 0.3), (INT64, 3000, 0.5), (INT32, 1, 0.0),
                                               (INT32, 1984 * 2, 1.0)])
def test_fixed_page_writer_decodes(dtype, n, null_frac):
    rng = np.random.default_rng(1)
    cells = H.random_cells(rng, dtype, n, null_frac)
    pages = job.write_fixed_pages(cells.values, cells.valid.astype(bool), rj.DataType(dtype))
    back = orc.decode(rj.Column(dtype, pages), n, impl="port")
    assert np.array_equal(back.valid, cells.valid) and np.array_equal(back.bits(), cells.bits())


@pytest.mark.parametrize("n,null_frac,long_frac,max_len", [(3000, 0.2, 0.0, 80), (500, 0.1, 0.03, 300), (4000, 1.0, 0.0, 5),
                                                           (2000, 0.0, 0.0, 0), (50, 0.0, 0.5, 10)])
def test_varchar_page_writer_decodes(n, null_frac, long_frac, max_len):
    rng = np.random.default_rng(2)
    cells = H.random_cells(rng, VARCHAR, n, null_frac, max_len=max_len, long_frac=long_frac)
    lens = (cells.str_off[1:] - cells.str_off[:-1]).astype(np.int64)
    pages = job.write_varchar_pages(lens, cells.chars, cells.valid.astype(bool))
    back = orc.decode(rj.Column(VARCHAR, pages), n, impl="port")
    assert back.to_python() == cells.to_python()
    # the greedy packing is as tight as the reference's own writer
    assert len(pages) <= orc.encode([cells], impl="port").columns[0].n_pages + 1


def test_job_1a_on_the_oracles():
    plan, root_cols, scan_rows = job.make_job("1a", scale=0.004, seed=1)
    a = orc.execute(plan, impl="port")
    assert [int(c.type) for c in a.columns] == [VARCHAR, VARCHAR, INT32]
    if orc.available("ref"):
        assert orc.result_equal(a, orc.execute(plan, impl="ref"))


@pytest.mark.gpu
@pytest.mark.parametrize("name,scale,long_strings", [("1a", 0.05, False), ("13a", 0.05, True), ("33a", 0.1, True),
                                                     ("24a", 0.01, False), ("31c", 0.01, True), ("17e", 0.01, False),
                                                     ("16b", 0.004, False),
                                                     # BASELINE.json configs 3 and 4 "at IMDB row counts"
                                                     ("1a", 1.0, False), ("13a", 1.0, True), ("33a", 1.0, True)])
def test_job_plans_match_the_oracle_on_gpu(name, scale, long_strings):
    """configs 3 and 4: JOB 1a (5-way), 13a (9 scans, depth 6), 33a (14 scans), 24a / 31c (12 / 11 scans,
    non-empty) with NULL bitmaps and long-string page chains in the VARCHAR columns; bit-exact multiset
    equality with the CPU oracle.  The scale-1.0 cases run at the IMDB row counts of plans.json (largest
    scan: movie_info, 14.8 M rows)."""
    plan, root_cols, scan_rows = job.make_job(name, scale=scale, seed=3, long_strings=long_strings)
    ctx = rj.build_context(0)
    try:
        got = rj.execute(plan, ctx)
    finally:
        rj.destroy_context(ctx)
    want = orc.execute(plan, impl="port")
    assert got.num_rows == want.num_rows
    assert [int(c.type) for c in got.columns] == [int(t) for _, _, t in root_cols]
    assert orc.result_equal(got, want)


@pytest.mark.gpu
def test_whole_job_suite_matches_the_oracle_on_gpu():
    """config 5 (shape): all 113 plans of plans.json (864 joins, 977 scans) on synthetic IMDB-shaped data at
    1 % of the IMDB cardinalities, every result compared with the CPU oracle"""
    ctx = rj.build_context(0)
    non_empty = 0
    try:
        for name in job.workload()["queries"]:
            plan, root_cols, _ = job.make_job(name, scale=0.01, seed=1)
            got = rj.execute(plan, ctx)
            want = orc.execute(plan, impl="port")
            assert got.num_rows == want.num_rows, name
            assert orc.result_equal(got, want), name
            non_empty += got.num_rows > 0
    finally:
        rj.destroy_context(ctx)
    assert non_empty >= 60
