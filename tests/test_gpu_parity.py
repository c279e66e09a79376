"""GPU tests (-m gpu): the CUDA path, called through the C-ABI, against the oracle.

Bit-exact multiset equality of result rows (INT32/INT64/FP64 by bit pattern, VARCHAR by bytes, NULL
distinct from every value) -- the check of the reference harness (tests/read_sql.cpp:1159-1222).
"""
import base64
import ctypes as C
import json
import os
import subprocess
import zlib

import numpy as np
import pytest

import helpers as H
from helpers import FP64, INT32, INT64, VARCHAR, orc, rj

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def ctx():
    c = rj.build_context(0)
    yield c
    rj.destroy_context(c)


def oracle_impl():
    return "ref" if orc.available("ref") else "port"


def check_plan(plan, ctx, impl=None, expect_rows=None):
    got = rj.execute(plan, ctx)
    want = orc.execute(plan, impl=impl or oracle_impl())
    assert got.num_rows == want.num_rows
    assert [int(c.type) for c in got.columns] == [int(c.type) for c in want.columns]
    if expect_rows is not None:
        assert got.num_rows == expect_rows
    assert orc.result_equal(got, want)
    return got


# ---- golden vectors ----------------------------------------------------------------------------------
def load_fixture():
    with open(os.path.join(H.ROOT, "tests", "golden", "unit_cases.json")) as f:
        return json.load(f)


def plan_from_fixture(d):
    plan = rj.Plan()
    for t in d["inputs"]:
        cols = []
        for c in t["columns"]:
            raw = zlib.decompress(base64.b64decode(c["pages"]))
            cols.append(rj.Column(c["type"], np.frombuffer(raw, dtype=np.uint8).reshape(-1, 8192).copy()))
        plan.new_input(rj.ColumnarTable(num_rows=t["num_rows"], columns=cols))
    for n in d["nodes"]:
        out = [(i, t) for i, t in n["out"]]
        if "join" in n:
            bl, l, r, la, ra = n["join"]
            plan.new_join_node(bl, l, r, la, ra, out)
        else:
            plan.new_scan_node(n["scan"], out)
    plan.root = d["root"]
    return plan


FIXTURE_NAMES = ["empty_join", "one_line_join", "simple_join", "empty_result", "multiple_same_keys",
                 "null_keys", "multiple_columns", "build_on_right", "int64_key_fp64_varchar_payload",
                 "varchar_key", "long_strings", "multi_page_nulls_dup_attrs"]


@pytest.mark.parametrize("name", FIXTURE_NAMES)
def test_golden_fixture(ctx, name):
    """committed fixtures: inputs + the rows the UNMODIFIED reference produced for them"""
    case = load_fixture()[name]
    got = rj.execute(plan_from_fixture(case["plan"]), ctx)
    assert got.num_rows == case["num_rows"]
    assert [int(c.type) for c in got.columns] == case["types"]
    if case["num_rows"] == 0:
        assert all(c.n_pages == 0 for c in got.columns)  # typed, page-less (unit_tests.cpp:24-27)
    want = []
    for r in case["rows"]:
        row = []
        for v, t in zip(r, case["types"]):
            if v is None:
                row.append(None)
            elif t == VARCHAR:
                row.append(v.encode())
            elif t == FP64:
                row.append(float.fromhex(v) if isinstance(v, str) else float(v))
            else:
                row.append(v)
        want.append(tuple(row))
    got_rows = H.rows_of(got)
    assert len(got_rows) == len(want)
    # compare FP64 by bit pattern (-0.0 != 0.0 here)
    canon = lambda rows: sorted(repr(tuple(np.float64(v).tobytes().hex() if isinstance(v, float) else v
                                           for v in r)) for r in rows)
    assert canon(got_rows) == canon(want)


def test_reference_unit_tests_on_the_engine():
    """the reference's tests/unit_tests.cpp, compiled UNMODIFIED and linked against
    Contest::execute of this repo (radix-join_b200/csrc/contest_execute.cpp)"""
    exe = os.path.join(H.ROOT, "oracle", "_ref", "unit_tests_b200")
    if not os.path.exists(exe):
        pytest.skip("unit_tests_b200 not built (needs the reference headers at build time)")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "0 failures / 8 cases" in out.stdout


# ---- random plans vs the oracle ------------------------------------------------------------------------
@pytest.mark.parametrize("key_type", [INT32, INT64, VARCHAR])
@pytest.mark.parametrize("build_left", [True, False])
def test_single_join_all_types(ctx, key_type, build_left):
    rng = np.random.default_rng(100 + key_type)
    lt = [key_type, INT64, VARCHAR]
    rt = [FP64, key_type, INT32]
    tl, _ = H.random_table(rng, lt, 1500, key_cols=(0,), key_range=600)
    tr, _ = H.random_table(rng, rt, 6000, key_cols=(1,), key_range=600, long_frac=0.002)
    plan = H.single_join_plan(tl, tr, lt, rt, 0, 1, build_left, out_cols=[5, 0, 2, 1, 3, 3, 4])
    check_plan(plan, ctx)


def test_multi_join_tree(ctx):
    import test_oracle
    plan = test_oracle._three_way_plan(np.random.default_rng(3))
    got = check_plan(plan, ctx)
    assert got.num_rows > 0


def test_scan_as_root_and_empty_side(ctx):
    t = H.table_from_python([INT32, VARCHAR], [(1, "a"), (None, ""), (3, None)])
    plan = rj.Plan()
    plan.new_input(t)
    plan.root = plan.new_scan_node(0, [(1, VARCHAR), (0, INT32), (1, VARCHAR)])
    check_plan(plan, ctx, expect_rows=3)
    plan = H.single_join_plan(t, H.empty_table([INT32]), [INT32, VARCHAR], [INT32], 0, 0, True)
    got = check_plan(plan, ctx, expect_rows=0)
    assert all(c.n_pages == 0 for c in got.columns)


def test_key_type_mismatch_matches_nothing(ctx):
    t = H.table_from_python([INT32], [(1,), (2,)])
    plan = rj.Plan()
    plan.new_input(t)
    plan.new_input(t)
    plan.new_scan_node(0, [(0, INT64)])
    plan.new_scan_node(1, [(0, INT32)])
    plan.root = plan.new_join_node(True, 0, 1, 0, 0, [(1, INT32)])
    check_plan(plan, ctx, expect_rows=0)


def test_errors_surface_as_engine_errors(ctx):
    t = H.table_from_python([INT32], [(1,)])
    plan = rj.Plan()
    plan.new_input(t)
    plan.new_scan_node(0, [(0, INT32)])
    plan.root = 7  # out of range
    with pytest.raises(rj.EngineError):
        rj.execute(plan, ctx)
    # FP64 join key: the reference never terminates on it (src/execute.cpp:28-31); the engine refuses
    tf = H.table_from_python([FP64], [(1.0,)])
    plan = H.single_join_plan(tf, tf, [FP64], [FP64], 0, 0, True)
    with pytest.raises(rj.EngineError):
        rj.execute(plan, ctx)
    # and the context is still usable afterwards
    check_plan(H.single_join_plan(t, t, [INT32], [INT32], 0, 0, True), ctx, expect_rows=1)


def _mistyped_plan(declared):
    tl = H.table_from_python([INT32, INT64], [(1, 10), (2, None), (3, 30), (4, None)])
    tr = H.table_from_python([INT32], [(1,), (2,), (3,), (4,), (2,)])
    plan = rj.Plan()
    plan.new_scan_node(0, [(0, INT32), (1, declared)])  # column 1 physically holds INT64
    plan.new_scan_node(1, [(0, INT32)])
    plan.new_join_node(True, 0, 1, 0, 0, [(0, INT32), (1, declared)])
    plan.new_input(tl)
    plan.new_input(tr)
    plan.root = 2
    return plan


def test_mistyped_root_attribute_keeps_only_the_null_rows(ctx):
    """Table::to_columnar silently skips cells whose alternative is not the column type and keeps the
    NULLs (src/build_table.cpp:484-501): the column comes out shorter than num_rows.  Same here."""
    plan = _mistyped_plan(INT32)
    got = rj.execute(plan, ctx)
    for impl in (["ref"] if orc.available("ref") else []) + ["port"]:
        want = orc.execute(plan, impl=impl)
        assert got.num_rows == want.num_rows == 5
        assert [int(c.type) for c in got.columns] == [int(c.type) for c in want.columns] == [INT32, INT32]
        for gc, wc in zip(got.columns, want.columns):
            n = sum(int(pg[0]) | int(pg[1]) << 8 for pg in wc.pages)
            assert n == sum(int(pg[0]) | int(pg[1]) << 8 for pg in gc.pages)
            assert orc.decode(gc, n).to_python() == orc.decode(wc, n).to_python()
    assert orc.decode(got.columns[1], 3).to_python() == [None, None, None]
    # no NULL among the selected rows: the column has no page at all
    plan.inputs[1] = H.table_from_python([INT32], [(1,), (3,)])
    got = rj.execute(plan, ctx)
    want = orc.execute(plan, impl="port")
    assert got.num_rows == want.num_rows == 2 and got.columns[1].n_pages == want.columns[1].n_pages == 0


def test_mistyped_varchar_root_attribute_throws_like_the_reference(ctx):
    """src/build_table.cpp:667-669: 'not string or null' at the first non-NULL cell; all-NULL passes"""
    plan = _mistyped_plan(VARCHAR)
    with pytest.raises(orc.OracleError, match="not string or null"):
        orc.execute(plan, impl="port")
    with pytest.raises(rj.EngineError, match="not string or null"):
        rj.execute(plan, ctx)
    plan.inputs[1] = H.table_from_python([INT32], [(2,), (4,), (4,)])  # only rows whose INT64 cell is NULL
    check_plan(plan, ctx, expect_rows=3)


def _with_orphan_page(rows):
    t = H.table_from_python([INT32, VARCHAR], rows)
    orphan = np.zeros((1, 8192), np.uint8)
    orphan[0, 0], orphan[0, 1], orphan[0, 2] = 0xfe, 0xff, 3
    orphan[0, 4:7] = [120, 121, 122]
    t.columns[1] = rj.Column(VARCHAR, np.concatenate([t.columns[1].pages, orphan]))
    plan = rj.Plan()
    plan.new_scan_node(0, [(0, INT32), (1, VARCHAR)])
    plan.new_input(t)
    plan.root = 0
    return plan


def test_continuation_page_after_a_regular_string_is_appended(ctx):
    """a 0xfffe page extends the row before it, whatever page that row came from
    (src/build_table.cpp:392-405)"""
    plan = _with_orphan_page([(1, None), (2, "bb")])
    got = check_plan(plan, ctx, expect_rows=2)
    assert H.rows_of(got) == [(1, None), (2, b"bbxyz")]
    # ... and through a join, where the string is re-encoded from its descriptor
    t2 = H.table_from_python([INT32], [(2,), (2,), (1,)])
    p2 = rj.Plan()
    p2.new_scan_node(0, [(0, INT32), (1, VARCHAR)])
    p2.new_scan_node(1, [(0, INT32)])
    p2.new_join_node(False, 0, 1, 0, 0, [(1, VARCHAR), (2, INT32)])
    p2.new_input(plan.inputs[0])
    p2.new_input(t2)
    p2.root = 2
    check_plan(p2, ctx, expect_rows=3)


def test_continuation_page_after_a_null_row_is_an_error(ctx):
    """the reference throws 'long string page 0xfffe must follows a string' (build_table.cpp:401-402;
    from a pool thread, so its process terminates); the engine raises the same message"""
    plan = _with_orphan_page([(1, "aa"), (2, None)])
    with pytest.raises(orc.OracleError, match="0xfffe must follows a string"):
        orc.execute(plan, impl="port")
    with pytest.raises(rj.EngineError, match="0xfffe must follows a string"):
        rj.execute(plan, ctx)
    check_plan(_with_orphan_page([(1, None), (2, "bb")]), ctx, expect_rows=2)  # the context survives


def test_duplicates_on_both_sides_overflow_the_first_guess(ctx):
    """M >> max(|build|, |probe|): the speculative output buffer overflows and the exact count sizes
    the second run ('unknown, possibly exploding output cardinality')"""
    rng = np.random.default_rng(9)
    a = H.table_from_cells([orc.Cells.from_values(INT32, rng.integers(0, 20, 3000).astype(np.int32))])
    b = H.table_from_cells([orc.Cells.from_values(INT32, rng.integers(0, 20, 4000).astype(np.int32))])
    got = check_plan(H.single_join_plan(a, b, [INT32], [INT32], 0, 0, True), ctx, impl="port")
    assert got.num_rows > 400_000


def test_one_heavy_build_key_uses_the_overflow_path(ctx):
    """one key repeated 20000 times on the build side: its partition exceeds a shared-memory table
    and is processed as several build chunks"""
    rng = np.random.default_rng(10)
    bk = np.concatenate([np.full(20000, 7, np.int32), rng.permutation(30000).astype(np.int32) + 100])
    pk = np.concatenate([np.full(5, 7, np.int32), rng.integers(100, 30100, 50000).astype(np.int32)])
    a = H.table_from_cells([orc.Cells.from_values(INT32, bk)])
    b = H.table_from_cells([orc.Cells.from_values(INT32, pk)])
    check_plan(H.single_join_plan(a, b, [INT32], [INT32], 0, 0, True), ctx, impl="port",
               expect_rows=5 * 20000 + 50000)


@pytest.mark.parametrize("n_build,n_probe", [(200_000, 1_000_000),     # one scatter pass (6 bits)
                                             (3_000_000, 4_000_000),   # two passes (10 bits)
                                             # 4 partitions of ~500 K probe tuples each = 8 work units per partition:
                                             # more units than the planner's unit -> partition table holds
                                             # (2 x partitions), so the fused join looks the others up by search
                                             (6_000, 2_000_000)])
def test_partitioned_join_with_payloads(ctx, n_build, n_probe):
    rng = np.random.default_rng(n_build)
    bk = rng.permutation(n_build).astype(np.int32)
    ba = orc.Cells(INT64, (rng.random(n_build) > 0.01).astype(np.uint8),
                   values=rng.integers(-2**62, 2**62, n_build))
    pk = orc.Cells(INT32, (rng.random(n_probe) > 0.02).astype(np.uint8),
                   values=rng.integers(0, int(n_build * 1.1), n_probe).astype(np.int32))
    pb = H.random_cells(rng, FP64, n_probe, null_frac=0.01)
    tl = H.table_from_cells([orc.Cells.from_values(INT32, bk), ba])
    tr = H.table_from_cells([pk, pb])
    plan = H.single_join_plan(tl, tr, [INT32, INT64], [INT32, FP64], 0, 0, True, out_cols=[0, 1, 3])
    check_plan(plan, ctx, impl="port")


def _fused_launches(ctx, plan, impl="port", expect_fused=True):
    """run the plan with stage profiling on; returns the result after checking it against the oracle and that
    the root join did (or did not) take the fused join + page-output kernel"""
    ctx.profile_enable(True)
    ctx.profile_reset()
    try:
        got = check_plan(plan, ctx, impl=impl)
        prof = ctx.profile_read()
    finally:
        ctx.profile_enable(False)
    assert (prof["join_emit"]["launches"] > 0) == expect_fused, prof
    return got, prof


@pytest.mark.parametrize("n_build,n_probe", [(40_000, 300_000), (2_500_000, 3_000_000)])  # one / two scatter passes
def test_fused_root_join_two_payloads_per_side(ctx, n_build, n_probe):
    """root join of two scans, INT32 key, 4- and 8-byte payloads with and without NULLs on both sides, the key
    taken from either side, one column twice: the join kernel writes the result pages itself"""
    rng = np.random.default_rng(n_build + 1)
    bk = orc.Cells(INT32, (rng.random(n_build) > 0.01).astype(np.uint8), values=rng.permutation(n_build).astype(np.int32))
    b1 = orc.Cells(INT64, (rng.random(n_build) > 0.05).astype(np.uint8), values=rng.integers(-2**62, 2**62, n_build))
    b2 = orc.Cells.from_values(INT32, rng.integers(-2**31, 2**31 - 1, n_build).astype(np.int32))
    pk = orc.Cells(INT32, (rng.random(n_probe) > 0.02).astype(np.uint8),
                   values=rng.integers(0, int(n_build * 1.2), n_probe).astype(np.int32))
    p1 = H.random_cells(rng, FP64, n_probe, null_frac=0.3)
    p2 = H.random_cells(rng, INT32, n_probe, null_frac=0.0)
    tl = H.table_from_cells([b1, bk, b2])
    tr = H.table_from_cells([p1, p2, pk])
    lt, rt = [INT64, INT32, INT32], [FP64, INT32, INT32]
    for build_left in (True, False):
        plan = H.single_join_plan(tl, tr, lt, rt, 1, 2, build_left, out_cols=[3, 1, 0, 4])
        _fused_launches(ctx, plan)
        plan = H.single_join_plan(tl, tr, lt, rt, 1, 2, build_left, out_cols=[5, 2, 3, 3])
        _fused_launches(ctx, plan)
    # the same plan on the general path gives the same multiset
    os.environ["RJ_NO_FUSED_ROOT"] = "1"
    try:
        _fused_launches(ctx, plan, expect_fused=False)
    finally:
        del os.environ["RJ_NO_FUSED_ROOT"]


def test_fused_root_join_gives_way_to_duplicate_build_keys(ctx):
    """the fused kernel needs distinct keys per table; a duplicate raises its flag and the general path
    (duplicate chains) produces the result"""
    rng = np.random.default_rng(9)
    n_b, n_p = 50_000, 200_000
    bkeys = rng.integers(0, 30_000, n_b).astype(np.int32)  # many duplicates
    tl = H.table_from_cells([orc.Cells.from_values(INT32, bkeys), H.random_cells(rng, INT64, n_b, null_frac=0.1)])
    tr = H.table_from_cells([orc.Cells.from_values(INT32, rng.integers(0, 30_000, n_p).astype(np.int32))])
    plan = H.single_join_plan(tl, tr, [INT32, INT64], [INT32], 0, 0, True)
    got, prof = _fused_launches(ctx, plan, expect_fused=True)  # it is tried first ...
    assert prof["join"]["launches"] > 0 and prof["encode"]["launches"] > 0  # ... then the general path ran


def test_fused_root_join_not_taken_for_varchar_or_wide_outputs(ctx):
    rng = np.random.default_rng(10)
    n_b, n_p = 20_000, 60_000
    tl = H.table_from_cells([orc.Cells.from_values(INT32, rng.permutation(n_b).astype(np.int32)),
                             H.random_cells(rng, VARCHAR, n_b, null_frac=0.1, max_len=12)])
    tr = H.table_from_cells([orc.Cells.from_values(INT32, rng.integers(0, n_b, n_p).astype(np.int32)),
                             H.random_cells(rng, INT64, n_p), H.random_cells(rng, INT64, n_p), H.random_cells(rng, FP64, n_p)])
    _fused_launches(ctx, H.single_join_plan(tl, tr, [INT32, VARCHAR], [INT32, INT64, INT64, FP64], 0, 0, True, out_cols=[0, 1, 3]),
                    expect_fused=False)
    # three carried columns on one side: more than the kernel stages
    _fused_launches(ctx, H.single_join_plan(tl, tr, [INT32, VARCHAR], [INT32, INT64, INT64, FP64], 0, 0, True, out_cols=[3, 4, 5]),
                    expect_fused=False)
    _fused_launches(ctx, H.single_join_plan(tl, tr, [INT32, VARCHAR], [INT32, INT64, INT64, FP64], 0, 0, True, out_cols=[0, 3, 5]),
                    expect_fused=True)


def test_config1_shape_1m_x_10m(ctx):
    """BASELINE config 1: single INT32 equi-join, 1 M build x 10 M probe, unique FK keys"""
    rng = np.random.default_rng(42)
    n_b, n_p = 1_000_000, 10_000_000
    tb = H.table_from_cells([orc.Cells.from_values(INT32, rng.permutation(n_b).astype(np.int32))])
    tp = H.table_from_cells([orc.Cells.from_values(INT32, rng.integers(0, n_b, n_p).astype(np.int32))])
    assert tb.columns[0].n_pages == 505 and tp.columns[0].n_pages == 5041
    plan = H.single_join_plan(tb, tp, [INT32], [INT32], 0, 0, True)
    got = rj.execute(plan, ctx)
    assert got.num_rows == n_p
    # size-independent property: both output columns equal the probe keys as a multiset
    cells = orc.decode_table(got)
    assert np.array_equal(cells[0].values, cells[1].values)
    assert np.array_equal(np.sort(cells[0].values), np.sort(orc.decode(tp.columns[0], n_p).values))
    # and the full oracle comparison on this size (the port finishes in seconds)
    assert orc.result_equal(got, orc.execute(plan, impl="port"))


def test_zipf_skew(ctx):
    rng = np.random.default_rng(44)
    n_b, n_p = 500_000, 3_000_000
    perm = rng.permutation(n_b).astype(np.int32)
    ranks = np.minimum(rng.zipf(1.3, n_p) - 1, n_b - 1)
    tb = H.table_from_cells([orc.Cells.from_values(INT32, perm)])
    tp = H.table_from_cells([orc.Cells.from_values(INT32, perm[ranks])])
    check_plan(H.single_join_plan(tb, tp, [INT32], [INT32], 0, 0, False), ctx, impl="port", expect_rows=n_p)


# ---- per-stage entry points ----------------------------------------------------------------------------
def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def unpack_valid(words, n):
    bits = np.unpackbits(words.view(np.uint8), bitorder="little")
    return bits[:n]


@pytest.mark.parametrize("type,n,null_frac", [(INT32, 100_000, 0.0), (INT32, 100_000, 0.3), (INT64, 50_000, 0.5),
                                              (FP64, 50_000, 0.05), (INT32, 200_000, 1.0), (INT64, 140_000, 0.999)])
def test_stage_decode_fixed(ctx, type, n, null_frac):
    """rj_page_row_offsets + rj_decode_fixed vs the oracle's from_columnar, incl. all-NULL pages of
    65504 / 65472 rows (SURVEY section 7 'pages with up to 65 504 rows')"""
    rng = np.random.default_rng(1)
    cells = H.random_cells(rng, type, n, null_frac)
    col = orc.encode([cells], impl=H.encoder_impl()).columns[0]
    lib, h = ctx.lib, ctx.handle
    d_pages = dev(col.pages)
    d_start = torch.zeros(col.n_pages + 1, dtype=torch.int64, device="cuda")
    d_tot = torch.zeros(2, dtype=torch.int64, device="cuda")
    ctx.check(lib.rj_page_row_offsets(h, d_pages.data_ptr(), col.n_pages, type, d_start.data_ptr(), d_tot.data_ptr(), None))
    w = 4 if type == INT32 else 8
    d_vals = torch.zeros(n * w, dtype=torch.uint8, device="cuda")
    d_valid = torch.zeros((n + 31) // 32 + 1, dtype=torch.int32, device="cuda")
    ctx.check(lib.rj_decode_fixed(h, d_pages.data_ptr(), col.n_pages, type, d_start.data_ptr(), d_vals.data_ptr(),
                                  d_valid.data_ptr(), None))
    torch.cuda.synchronize()
    assert d_tot.cpu().tolist() == [n, int(cells.valid.sum())]
    assert int(d_start.cpu()[-1]) == n
    valid = unpack_valid(d_valid.cpu().numpy(), n)
    assert np.array_equal(valid, cells.valid)
    vals = d_vals.cpu().numpy().view(np.uint32 if w == 4 else np.uint64)
    got = vals.astype(np.uint64)
    got[valid == 0] = 0
    assert np.array_equal(got, cells.bits())


@pytest.mark.parametrize("n,null_frac,long_frac,max_len", [(20_000, 0.1, 0.0, 60), (3_000, 0.2, 0.02, 300),
                                                           (70_000, 1.0, 0.0, 10), (5_000, 0.0, 0.0, 0),
                                                           (2_000, 0.1, 0.0, 3000)])
def test_stage_varchar_roundtrip(ctx, n, null_frac, long_frac, max_len):
    """rj_decode_varchar -> rj_encode_varchar_plan/write (through a shuffled row-id list) -> oracle decode"""
    rng = np.random.default_rng(2)
    cells = H.random_cells(rng, VARCHAR, n, null_frac, max_len=max_len, long_frac=long_frac)
    col = orc.encode([cells], impl=H.encoder_impl()).columns[0]
    lib, h = ctx.lib, ctx.handle
    d_pages = dev(col.pages)
    d_start = torch.zeros(col.n_pages + 1, dtype=torch.int64, device="cuda")
    ctx.check(lib.rj_page_row_offsets(h, d_pages.data_ptr(), col.n_pages, VARCHAR, d_start.data_ptr(), None, None))
    d_desc = torch.zeros(n, dtype=torch.int64, device="cuda")
    d_valid = torch.zeros((n + 31) // 32 + 1, dtype=torch.int32, device="cuda")
    ctx.check(lib.rj_decode_varchar(h, d_pages.data_ptr(), col.n_pages, d_start.data_ptr(), d_desc.data_ptr(),
                                    d_valid.data_ptr(), None))
    torch.cuda.synchronize()
    assert np.array_equal(unpack_valid(d_valid.cpu().numpy(), n), cells.valid)
    desc = d_desc.cpu().numpy().view(np.uint64)
    lens = ((desc >> np.uint64(40)) & np.uint64(0x7FFFFF)).astype(np.int64)
    want_lens = (cells.str_off[1:] - cells.str_off[:-1]).astype(np.int64)
    assert np.array_equal(lens[cells.valid == 1], want_lens[cells.valid == 1])
    # re-encode through a permutation with repeats
    m = n + n // 3
    idx = rng.integers(0, n, m).astype(np.uint32)
    d_idx = dev(idx)
    layout, n_pages = C.c_void_p(), C.c_uint64()
    ctx.check(lib.rj_encode_varchar_plan(h, d_pages.data_ptr(), d_desc.data_ptr(), d_valid.data_ptr(), d_idx.data_ptr(), m,
                                         C.byref(layout), C.byref(n_pages), None))
    out = torch.zeros(max(n_pages.value, 1) * 8192, dtype=torch.uint8, device="cuda")
    ctx.check(lib.rj_encode_varchar_write(h, layout, out.data_ptr(), None))
    torch.cuda.synchronize()
    lib.rj_encode_varchar_free(h, layout)
    pages = out.cpu().numpy()[: n_pages.value * 8192].reshape(-1, 8192)
    back = orc.decode(rj.Column(VARCHAR, pages), m, impl=H.encoder_impl())
    want = cells.to_python()
    assert back.to_python() == [want[i] for i in idx]
    # page budget: the parallel layout may under-fill pages, but not by much for short strings
    if max_len and max_len <= 300 and long_frac == 0 and null_frac < 1:
        greedy = orc.encode([orc.Cells.from_strings([want[i] for i in idx])], impl=H.encoder_impl()).columns[0].n_pages
        assert n_pages.value <= greedy * 1.15 + 2


@pytest.mark.parametrize("key_bytes", [4, 8])
def test_stage_histogram_scatter(ctx, key_bytes):
    """rj_radix_histogram / rj_radix_scatter: counts match numpy; the scatter is a permutation that
    groups tuples by radix digit and drops NULL keys"""
    hash_keys = H.hash_keys
    rng = np.random.default_rng(3)
    n, bits, shift = 300_000, 7, 3
    keys = (rng.integers(0, 50_000, n).astype(np.uint32) if key_bytes == 4
            else rng.integers(0, 2**63, n).astype(np.uint64))
    validb = (rng.random(n) > 0.1)
    words = np.packbits(validb, bitorder="little")
    words = np.concatenate([words, np.zeros((-len(words)) % 4 + 4, np.uint8)]).view(np.uint32)
    lib, h = ctx.lib, ctx.handle
    d_keys, d_valid = dev(keys), dev(words.view(np.int32))
    d_hist = torch.zeros(1 << bits, dtype=torch.int32, device="cuda")
    ctx.check(lib.rj_radix_histogram(h, d_keys.data_ptr(), d_valid.data_ptr(), n, key_bytes, shift, bits, d_hist.data_ptr(), None))
    digit = (hash_keys(keys) >> np.uint32(shift)) & np.uint32((1 << bits) - 1)
    want_hist = np.bincount(digit[validb], minlength=1 << bits)
    torch.cuda.synchronize()
    assert np.array_equal(d_hist.cpu().numpy(), want_hist)
    start = np.concatenate([[0], np.cumsum(want_hist)]).astype(np.uint32)
    d_cur = dev(start[:-1].view(np.int32).copy())
    d_ko = torch.zeros(n * key_bytes, dtype=torch.uint8, device="cuda")
    d_io = torch.zeros(n, dtype=torch.int32, device="cuda")
    ctx.check(lib.rj_radix_scatter(h, d_keys.data_ptr(), d_valid.data_ptr(), None, n, key_bytes, shift, bits,
                                   d_cur.data_ptr(), d_ko.data_ptr(), d_io.data_ptr(), None))
    torch.cuda.synchronize()
    n_valid = int(validb.sum())
    ko = d_ko.cpu().numpy().view(keys.dtype)[:n_valid]
    io = d_io.cpu().numpy().view(np.uint32)[:n_valid]
    assert np.array_equal(np.sort(io), np.nonzero(validb)[0].astype(np.uint32))  # a permutation of the valid rows
    assert np.array_equal(keys[io], ko)                                          # key travels with its row id
    dig_out = (hash_keys(ko) >> np.uint32(shift)) & np.uint32((1 << bits) - 1)
    assert np.all(np.diff(dig_out.astype(np.int64)) >= 0)                        # grouped by digit
    assert np.array_equal(d_cur.cpu().numpy().view(np.uint32), start[1:])        # cursors advanced to the ends


def test_stage_join_keys_and_gather(ctx):
    rng = np.random.default_rng(4)
    nb, np_ = 40_000, 150_000
    bk = rng.integers(0, 30_000, nb).astype(np.uint32)
    pk = rng.integers(0, 60_000, np_).astype(np.uint32)
    lib, h = ctx.lib, ctx.handle
    d_bk, d_pk = dev(bk), dev(pk)
    cap = 1 << 20
    d_ob = torch.zeros(cap, dtype=torch.int32, device="cuda")
    d_op = torch.zeros(cap, dtype=torch.int32, device="cuda")
    m = C.c_uint64()
    ctx.check(lib.rj_join_keys(h, d_bk.data_ptr(), None, nb, d_pk.data_ptr(), None, np_, 4, cap, d_ob.data_ptr(),
                               d_op.data_ptr(), C.byref(m), None))
    cnt = np.bincount(bk, minlength=60_000)
    assert m.value == int(cnt[pk].sum())
    ob = d_ob.cpu().numpy().view(np.uint32)[: m.value]
    op = d_op.cpu().numpy().view(np.uint32)[: m.value]
    assert np.array_equal(bk[ob], pk[op])
    pairs = np.unique(ob.astype(np.uint64) << np.uint64(32) | op.astype(np.uint64))
    assert len(pairs) == m.value  # no pair twice
    # gather
    d_out = torch.zeros(m.value, dtype=torch.int32, device="cuda")
    ctx.check(lib.rj_gather(h, d_bk.data_ptr(), None, d_ob.data_ptr(), m.value, 4, d_out.data_ptr(), None, None))
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy().view(np.uint32), bk[ob])


def test_resident_inputs_and_profile(ctx):
    """upload once, execute twice from HBM; per-stage timers are populated"""
    rng = np.random.default_rng(6)
    tl, _ = H.random_table(rng, [INT32, INT64], 300_000, key_cols=(0,), key_range=10**6, null_frac=0.0, key_null_frac=0.0)
    tr, _ = H.random_table(rng, [INT32, FP64], 900_000, key_cols=(0,), key_range=10**6, null_frac=0.0, key_null_frac=0.0)
    plan = H.single_join_plan(tl, tr, [INT32, INT64], [INT32, FP64], 0, 0, True, out_cols=[0, 1, 3])
    inputs = rj.upload(plan, ctx)
    ctx.profile_enable(True)
    ctx.profile_reset()
    want = orc.execute(plan, impl="port")
    for _ in range(2):
        res = rj.execute_resident(plan, inputs, ctx)
        assert orc.result_equal(res.to_columnar(), want)
        res.free()
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    inputs.free()
    for st in ("decode", "histogram", "scatter", "join", "encode"):
        assert prof[st]["ms"] > 0 and prof[st]["launches"] > 0 and prof[st]["bytes"] > 0, (st, prof)


# ---- streamed execution (rj_execute_streamed): row windows of the largest table --------------------------
def check_streamed(plan, ctx, chunk_bytes, impl=None, min_chunks=2):
    want = orc.execute(plan, impl=impl or oracle_impl())
    calls = []

    def alloc(column, dtype, n_pages):
        calls.append((column, int(dtype), n_pages))
        return np.empty((n_pages, 8192), dtype=np.uint8)

    n, chunks = rj.execute_streamed(plan, ctx, chunk_bytes, alloc=alloc)
    assert n == want.num_rows
    got = rj.execute_streamed_columnar(plan, ctx, chunk_bytes)
    assert got.num_rows == want.num_rows
    assert [int(c.type) for c in got.columns] == [int(c.type) for c in want.columns]
    assert orc.result_equal(got, want)
    if min_chunks and want.num_rows:
        per_col = max(sum(1 for c in calls if c[0] == col) for col in range(len(want.columns)))
        assert per_col >= min_chunks, calls
    return got


@pytest.mark.parametrize("build_left", [True, False])
def test_streamed_probe_windows_match_the_oracle(ctx, build_left):
    # the right table is ~64 pages per column; 6-page windows cut every column in the middle of pages,
    # INT32 and 8-byte columns at different rows, with NULLs in keys and payloads
    rng = np.random.default_rng(77)
    lt, rt = [INT32, INT64], [FP64, INT32, INT64]
    tl, _ = H.random_table(rng, lt, 20_000, key_cols=(0,), key_range=15_000)
    tr, _ = H.random_table(rng, rt, 60_000, key_cols=(1,), key_range=15_000)
    plan = H.single_join_plan(tl, tr, lt, rt, 0, 1, build_left, out_cols=[1, 3, 2, 4, 0])
    check_streamed(plan, ctx, chunk_bytes=6 * 8192 * 3)


def test_streamed_build_side_and_multi_join(ctx):
    # the streamed table sits below two joins, on the build side of the first
    rng = np.random.default_rng(78)
    big, _ = H.random_table(rng, [INT32, INT64, INT32], 50_000, key_cols=(0, 2), key_range=3_000)
    a, _ = H.random_table(rng, [INT32, FP64], 3_000, key_cols=(0,), key_range=3_000)
    b, _ = H.random_table(rng, [INT32, VARCHAR], 2_000, key_cols=(0,), key_range=3_000)
    plan = rj.Plan()
    s_big = plan.new_scan_node(0, [(0, INT32), (1, INT64), (2, INT32)])
    s_a = plan.new_scan_node(1, [(0, INT32), (1, FP64)])
    s_b = plan.new_scan_node(2, [(0, INT32), (1, VARCHAR)])
    j1 = plan.new_join_node(True, s_big, s_a, 0, 0, [(0, INT32), (1, INT64), (2, INT32), (4, FP64)])
    plan.root = plan.new_join_node(False, j1, s_b, 2, 0, [(5, VARCHAR), (1, INT64), (3, FP64), (0, INT32)])
    for t in (big, a, b):
        plan.new_input(t)
    check_streamed(plan, ctx, chunk_bytes=5 * 8192 * 3)


def test_streamed_falls_back_when_the_table_cannot_be_windowed(ctx):
    rng = np.random.default_rng(79)
    # (a) the big table carries a VARCHAR column the plan reads, (b) it is scanned twice (self join)
    lt, rt = [INT32, INT64], [INT32, VARCHAR]
    tl, _ = H.random_table(rng, lt, 2_000, key_cols=(0,), key_range=1_500)
    tr, _ = H.random_table(rng, rt, 30_000, key_cols=(0,), key_range=1_500)
    check_streamed(H.single_join_plan(tl, tr, lt, rt, 0, 0, True), ctx, chunk_bytes=4 * 8192, min_chunks=0)
    t, _ = H.random_table(rng, [INT32, INT64], 30_000, key_cols=(0,), key_range=20_000)
    plan = rj.Plan()
    plan.new_scan_node(0, [(0, INT32), (1, INT64)])
    plan.new_scan_node(0, [(0, INT32), (1, INT64)])
    plan.root = plan.new_join_node(True, 0, 1, 0, 0, [(0, INT32), (1, INT64), (3, INT64)])
    plan.new_input(t)
    check_streamed(plan, ctx, chunk_bytes=4 * 8192, min_chunks=0)


def test_streamed_scan_root_and_larger_join(ctx):
    rng = np.random.default_rng(80)
    t, _ = H.random_table(rng, [INT64, INT32], 40_000, null_frac=0.2)
    plan = rj.Plan()
    plan.new_input(t)
    plan.root = plan.new_scan_node(0, [(1, INT32), (0, INT64)])
    check_streamed(plan, ctx, chunk_bytes=7 * 8192)
    # two scatter passes inside every window
    n_build, n_probe = 3_000_000, 6_000_000
    bk = rng.permutation(n_build).astype(np.int32)
    ba = orc.Cells(INT64, (rng.random(n_build) > 0.01).astype(np.uint8), values=rng.integers(-2**62, 2**62, n_build))
    pk = orc.Cells(INT32, (rng.random(n_probe) > 0.02).astype(np.uint8),
                   values=rng.integers(0, int(n_build * 1.1), n_probe).astype(np.int32))
    pb = H.random_cells(rng, FP64, n_probe, null_frac=0.01)
    tl = H.table_from_cells([orc.Cells.from_values(INT32, bk), ba])
    tr = H.table_from_cells([pk, pb])
    plan = H.single_join_plan(tl, tr, [INT32, INT64], [INT32, FP64], 0, 0, True, out_cols=[0, 1, 3])
    check_streamed(plan, ctx, chunk_bytes=24 << 20, impl="port")


# ---- BASELINE.json config 2 at FULL size: parity through a size-independent property ---------------------
def test_config2_full_size_multiset_checksum(ctx):
    """64 Mi x 512 Mi rows, Zipf(0.75), INT64 + FP64 payloads with NULLs.  The oracle cannot hold this, so
    the result is pinned by a multiset checksum (two wrapping sums over a per-row hash) whose expected value
    is derived from the generator by direct addressing -- no join involved.  Both the resident path and
    the streamed host-to-host path must reproduce it."""
    from radix_join_b200 import synthetic as syn
    free, _ = torch.cuda.mem_get_info()
    scale = 1 if free > 120 << 30 else 8
    nb, np_ = (64 << 20) // scale, (512 << 20) // scale
    dt = syn.make_c2_device(ctx, nb, np_, checksum=True)
    assert dt.expected_checksum[2] == np_
    inputs = rj.adopt_device(dt.plan, dt.device_pages, ctx, keep=dt.keep)
    res = rj.execute_resident(dt.plan, inputs, ctx)
    assert res.num_rows == np_
    assert syn.result_checksum(ctx, res) == dt.expected_checksum
    res.free()
    # streamed: host pages in, host pages out, 13 windows
    host_plan, keep = syn.to_host_plan(dt, pinned=False)
    inputs.free()
    rows, chunks = rj.execute_streamed(host_plan, ctx)
    assert rows == np_
    types = [t for _, t in host_plan.nodes[host_plan.root].output_attrs]
    assert max(len(v) for v in chunks.values()) >= 2
    assert syn.host_chunks_checksum(ctx, types, chunks) == dt.expected_checksum


def _carry_desc(keys, valid_words, n, shift, bits, cursor, vals, flags):
    from radix_join_b200 import _cabi
    d = _cabi.rj_carry_scatter_t()
    d.d_keys, d.d_valid, d.n = keys.data_ptr(), (valid_words.data_ptr() if valid_words is not None else None), n
    d.shift, d.bits, d.d_cursor = shift, bits, cursor.data_ptr()
    d.n_val, d.n_flag = len(vals), len(flags)
    return d


@pytest.mark.parametrize("widths", [(8,), (4, 8), (8, 4), (4, 4), ()])
def test_stage_scatter_carry_two_passes_and_scattered_sources(ctx, widths):
    """rj_scatter_carry: flat pass (bitmaps in, bytes out, NULL keys dropped), region pass over its output, and the
    same region pass reading its regions through a table of per-region source addresses (the multi-GPU pull) --
    every tuple must arrive in its final partition with its values and validity bytes; 4- and 8-byte columns in
    both orders (the kernel takes the wider one first)"""
    from radix_join_b200 import _cabi
    rng = np.random.default_rng(sum(widths) + len(widths))
    n, b1, b2 = 500_003, 5, 4
    keys = rng.integers(0, 1 << 30, n).astype(np.uint32)
    kvalid = rng.random(n) > 0.05
    vals = [rng.integers(0, 2**31 if w == 4 else 2**62, n).astype(np.uint32 if w == 4 else np.uint64) for w in widths]
    vvalid = [rng.random(n) > 0.3 for _ in widths]

    def words(mask):
        by = np.packbits(mask, bitorder="little")
        return dev(np.concatenate([by, np.zeros((-len(by)) % 4 + 8, np.uint8)]).view(np.int32))
    h = H.hash_keys(keys)
    full = h & np.uint32((1 << (b1 + b2)) - 1)
    d1 = full >> np.uint32(b2)
    n_valid = int(kvalid.sum())
    lib, hd = ctx.lib, ctx.handle
    d_keys, d_kvalid = dev(keys), words(kvalid)
    d_vals, d_vbits = [dev(v) for v in vals], [words(m) for m in vvalid]
    # ---- flat pass by the top b1 of the b1 + b2 bits
    cnt1 = np.bincount(d1[kvalid], minlength=1 << b1)
    start1 = np.concatenate([[0], np.cumsum(cnt1)]).astype(np.uint32)
    cur1 = dev(start1[:-1].view(np.int32).copy())
    k1 = torch.zeros(n + 16, dtype=torch.int32, device="cuda")
    v1 = [torch.zeros((n + 16) * (w // 4), dtype=torch.int32, device="cuda") for w in widths]
    f1 = [torch.zeros(n + 16, dtype=torch.uint8, device="cuda") for _ in widths]
    d = _carry_desc(d_keys, d_kvalid, n, b2, b1, cur1, vals, vvalid)
    d.d_keys_out = k1.data_ptr()
    for i, w in enumerate(widths):
        d.val_src[i], d.val_dst[i], d.val_width[i] = d_vals[i].data_ptr(), v1[i].data_ptr(), w
        d.flag_src[i], d.flag_dst[i] = d_vbits[i].data_ptr(), f1[i].data_ptr()
    ctx.check(lib.rj_scatter_carry(hd, C.byref(d), None))
    torch.cuda.synchronize()
    ko = k1.cpu().numpy().view(np.uint32)[:n_valid]
    assert np.all(np.diff(((H.hash_keys(ko) & np.uint32((1 << (b1 + b2)) - 1)) >> np.uint32(b2)).astype(np.int64)) >= 0)

    def check_final(kf, vf, ff):
        """kf / vf / ff: final arrays; compare the multiset of (key, values, flags) tuples with the input's"""
        got = [kf.astype(np.uint64)] + [v.astype(np.uint64) for v in vf] + [f.astype(np.uint64) for f in ff]
        want = [keys[kvalid].astype(np.uint64)] + [v[kvalid].astype(np.uint64) for v in vals] + [m[kvalid].astype(np.uint64) for m in vvalid]
        go = np.lexsort(got[::-1])
        wo = np.lexsort(want[::-1])
        for a, b in zip(got, want):
            assert np.array_equal(a[go], b[wo])
        assert np.all(np.diff((H.hash_keys(kf) & np.uint32((1 << (b1 + b2)) - 1)).astype(np.int64)) >= 0)  # final partition order

    # ---- region pass (second pass) over the flat pass's output
    cnt_full = np.bincount(full[kvalid], minlength=1 << (b1 + b2))
    off = np.concatenate([[0], np.cumsum(cnt_full)]).astype(np.uint32)
    tiles = (cnt1 + 4095) // 4096
    tile_start = np.concatenate([[0], np.cumsum(tiles)]).astype(np.uint32)

    def region_pass(extra):
        cur2 = dev(off[:-1].view(np.int32).copy())
        k2 = torch.zeros(n + 16, dtype=torch.int32, device="cuda")
        v2 = [torch.zeros((n + 16) * (w // 4), dtype=torch.int32, device="cuda") for w in widths]
        f2 = [torch.zeros(n + 16, dtype=torch.uint8, device="cuda") for _ in widths]
        d = _carry_desc(k1, None, n_valid, 0, b2, cur2, vals, vvalid)
        d.d_keys_out = k2.data_ptr()
        for i, w in enumerate(widths):
            d.val_src[i], d.val_dst[i], d.val_width[i] = v1[i].data_ptr(), v2[i].data_ptr(), w
            d.flag_src[i], d.flag_dst[i] = f1[i].data_ptr(), f2[i].data_ptr()
        keep = extra(d)
        ctx.check(lib.rj_scatter_carry(hd, C.byref(d), None))
        torch.cuda.synchronize()
        del keep
        kf = k2.cpu().numpy().view(np.uint32)[:n_valid]
        vf = [v.cpu().numpy().view(np.uint32 if w == 4 else np.uint64)[:n_valid] for v, w in zip(v2, widths)]
        ff = [f.cpu().numpy()[:n_valid] for f in f2]
        check_final(kf, vf, ff)

    def plain(d):
        rs, ts = dev(start1.view(np.int32).copy()), dev(tile_start.view(np.int32).copy())
        d.d_region_start, d.d_tile_start, d.n_regions = rs.data_ptr(), ts.data_ptr(), 1 << b1
        return rs, ts
    region_pass(plain)

    def scattered(d):
        # every region split in two sub-regions read through the address table, in swapped order inside the table's
        # address space: sub-region x = (region r, half q) with biased base addresses
        n_sub = 2 << b1
        sub_cnt = np.zeros(n_sub, dtype=np.int64)
        src_start = np.zeros(n_sub, dtype=np.int64)
        for r in range(1 << b1):
            a = int(cnt1[r]) // 3
            sub_cnt[2 * r], sub_cnt[2 * r + 1] = int(cnt1[r]) - a, a            # the second part of the run first
            src_start[2 * r], src_start[2 * r + 1] = int(start1[r]) + a, int(start1[r])
        vstart = np.concatenate([[0], np.cumsum(sub_cnt)])
        vtile = np.concatenate([[0], np.cumsum((sub_cnt + 4095) // 4096)])
        table = np.zeros((n_sub, 5), dtype=np.int64)
        bases = [k1.data_ptr()] + [v.data_ptr() for v in v1] + [0] * (2 - len(widths)) + [f.data_ptr() for f in f1] + [0] * (2 - len(widths))
        ws = [4] + list(widths) + [0] * (2 - len(widths)) + [1, 1]
        for a in range(5):
            table[:, a] = bases[a] + (src_start - vstart[:-1]) * ws[a]
        t_tab, t_start, t_tile = dev(table), dev(vstart.astype(np.int32)), dev(vtile.astype(np.int32))
        t_group = dev(np.repeat(np.arange(1 << b1, dtype=np.int32), 2))
        d.d_keys = None
        d.d_src_table, d.d_region_group = t_tab.data_ptr(), t_group.data_ptr()
        d.d_region_start, d.d_tile_start, d.n_regions = t_start.data_ptr(), t_tile.data_ptr(), n_sub
        for i in range(len(widths)):
            d.val_src[i] = None
            d.flag_src[i] = None
        return t_tab, t_start, t_tile, t_group
    region_pass(scattered)


@pytest.mark.parametrize("bits,p1,build_nulls", [(6, 0, False), (15, 7, True), (15, 7, False)])
def test_stage_join_partitioned(ctx, bits, p1, build_nulls):
    """rj_join_partitioned on inputs that a flat rj_scatter_carry pass partitioned completely (local_pass1_bits = 0:
    the hash-table kernel), and on pass-1 partitioned inputs with 15 radix bits in all: the engine runs the second
    pass itself and the join kernel uses its rank-structure table (the 17 hash bits the partitioning leaves identify
    a 32-bit key: k_join_emit.cuh, DIRECT)"""
    from radix_join_b200 import _cabi
    rng = np.random.default_rng(12)
    nb, np_ = 100_000, 700_000
    bk = (rng.permutation(4 * nb)[:nb] * 977 + 12345).astype(np.uint32)
    ba = rng.integers(0, 2**62, nb).astype(np.uint64)
    bvalid = (rng.random(nb) > 0.3) if build_nulls else None
    pk = np.where(rng.random(np_) < 0.9, bk[rng.integers(0, nb, np_)], rng.integers(0, 2**32, np_).astype(np.uint32)).astype(np.uint32)
    pb = rng.integers(0, 2**31, np_).astype(np.uint32)
    pvalid = rng.random(np_) > 0.2
    lib, hd = ctx.lib, ctx.handle

    def part(keys, vals, width, vmask):
        n = len(keys)
        part_id = H.hash_keys(keys) & np.uint32((1 << bits) - 1)
        cnt = np.bincount(part_id, minlength=1 << bits)
        pass_bits, shift = (p1, bits - p1) if p1 else (bits, 0)
        dig_cnt = np.bincount(part_id >> shift, minlength=1 << pass_bits)
        cur = dev(np.concatenate([[0], np.cumsum(dig_cnt)])[:-1].astype(np.int32))
        ko = torch.zeros(n + 16, dtype=torch.int32, device="cuda")
        vo = torch.zeros((n + 16) * (width // 4), dtype=torch.int32, device="cuda")
        fo = torch.zeros(n + 16, dtype=torch.uint8, device="cuda") if vmask is not None else None
        d_k, d_v = dev(keys), dev(vals)
        d = _cabi.rj_carry_scatter_t()
        d.d_keys, d.n, d.shift, d.bits, d.d_cursor, d.d_keys_out = d_k.data_ptr(), n, shift, pass_bits, cur.data_ptr(), ko.data_ptr()
        d.n_val, d.val_src[0], d.val_dst[0], d.val_width[0] = 1, d_v.data_ptr(), vo.data_ptr(), width
        if vmask is not None:
            by = np.packbits(vmask, bitorder="little")
            d_w = dev(np.concatenate([by, np.zeros((-len(by)) % 4 + 8, np.uint8)]).view(np.int32))
            d.n_flag, d.flag_src[0], d.flag_dst[0] = 1, d_w.data_ptr(), fo.data_ptr()
        ctx.check(lib.rj_scatter_carry(hd, C.byref(d), None))
        torch.cuda.synchronize()
        return ko, vo, fo, dev(cnt.astype(np.int32))
    bko, bvo, bfo, bh = part(bk, ba, 8, bvalid)
    pko, pvo, pfo, ph = part(pk, pb, 4, pvalid)
    sides = []
    for ko, vo, fo, n, t in ((bko, bvo, bfo, nb, INT64), (pko, pvo, pfo, np_, INT32)):
        sd = _cabi.rj_part_side_t()
        sd.d_keys, sd.n, sd.n_cols = ko.data_ptr(), n, 1
        sd.d_vals[0], sd.types[0], sd.d_valid_bytes[0] = vo.data_ptr(), int(t), (fo.data_ptr() if fo is not None else None)
        sides.append(sd)
    outs = (_cabi.rj_part_out_t * 3)()
    outs[0].side, outs[0].col = 1, -1
    outs[1].side, outs[1].col = 0, 0
    outs[2].side, outs[2].col = 1, 0
    hres = C.c_void_p()
    ctx.check(lib.rj_join_partitioned(hd, C.byref(sides[0]), C.byref(sides[1]), bh.data_ptr(), ph.data_ptr(), bits, p1, bits, outs, 3, C.byref(hres)))
    assert hres.value
    from radix_join_b200.engine import Result
    res = Result(ctx, hres)
    got = res.to_columnar()
    res.free()
    inv = {int(k): i for i, k in enumerate(bk)}
    hit = np.array([int(k) in inv for k in pk])
    assert got.num_rows == int(hit.sum())
    def build_val(k):
        i = inv[int(k)]
        return int(ba[i]) if bvalid is None or bvalid[i] else None
    want_rows = H.sort_rows([(int(k), build_val(k), (int(v) if ok else None)) for k, v, ok in zip(pk[hit], pb[hit], pvalid[hit])])
    assert H.rows_of(got) == want_rows
