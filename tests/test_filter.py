"""Pre-filter evaluation and filter + emit: the reference's Statement::eval over
InnerColumns (src/statement.cpp:46-133,186-200; include/inner_column.h:170-325,386-562) and the
row-emitting tail of Table::from_csv (src/build_table.cpp:94-119,247-303).

CPU tests pin the numpy restatement (oracle/filter_oracle.py) to the UNMODIFIED reference
(oracle/_ref/libref_oracle.so: ref_filter); GPU tests compare the CUDA kernels (csrc/k_filter.cu, through
rj_filter_table of the C-ABI) with the restatement -- and with the reference itself where its library
travelled -- on the same seeded inputs: same rows, same order, NULLs and strings bit-exact."""
import numpy as np
import pytest

import helpers as H
from helpers import FP64, INT32, INT64, VARCHAR, orc, rj
from oracle import filter_oracle as fo
from radix_join_b200 import Comparison as Cmp
from radix_join_b200 import LogicalOperation as Lop

WORDS = ["", "a", "ab", "abc", "(co-production)", "x(co-production)y", "50%", "a_b", "a.b", "a\nb", "Zürich", "日本語", "[1]",
         "^start$", "back\\slash", "%", "_", "abcabc", "abd"]


def make_table(rng, n, long_frac=0.0, ascii_only=False):
    k = H.random_cells(rng, INT32, n, null_frac=0.1, key_range=50)
    a = orc.Cells(INT64, (rng.random(n) > 0.2).astype(np.uint8),
                  values=rng.choice(np.array([-2**40, -1, 0, 1, 7, 2**31, 2**32 + 5, 2**53 + 1, 2**62], dtype=np.int64), n))
    f = orc.Cells(FP64, (rng.random(n) > 0.1).astype(np.uint8),
                  values=rng.choice(np.array([-1.5, -0.0, 0.0, 1.0, 2.5, np.nan, np.inf, 1e300]), n))
    words = [w for w in WORDS if w.isascii()] if ascii_only else WORDS
    picks = rng.integers(0, len(words), n)
    strings = [None if rng.random() < 0.15 else words[i] for i in picks]
    if long_frac > 0:
        for i in range(n):
            if strings[i] is not None and rng.random() < long_frac:
                strings[i] = ("L" * int(rng.integers(8186, 20000))) + strings[i]
    s = orc.Cells.from_strings(strings)
    cells = [k, a, f, s]
    return H.table_from_cells(cells), cells


STATEMENTS = [
    Cmp(0, Cmp.EQ, 7), Cmp(0, Cmp.NEQ, 7), Cmp(0, Cmp.LT, 25), Cmp(0, Cmp.GEQ, 25),
    Cmp(0, Cmp.EQ, 2**32 + 7),            # narrowed to int32 7 like the reference does
    Cmp(1, Cmp.GT, 2**31), Cmp(1, Cmp.LEQ, -1), Cmp(1, Cmp.EQ, 2**53 + 1), Cmp(1, Cmp.NEQ, 0),
    Cmp(2, Cmp.LT, 1.0), Cmp(2, Cmp.GEQ, 0.0), Cmp(2, Cmp.EQ, 0.0), Cmp(2, Cmp.NEQ, float("nan")), Cmp(2, Cmp.LEQ, float("inf")),
    Cmp(0, Cmp.IS_NULL), Cmp(1, Cmp.IS_NOT_NULL), Cmp(3, Cmp.IS_NULL),
    Cmp(3, Cmp.EQ, "abc"), Cmp(3, Cmp.NEQ, "abc"), Cmp(3, Cmp.LT, "abc"), Cmp(3, Cmp.GT, "ab"), Cmp(3, Cmp.LEQ, "Zürich"), Cmp(3, Cmp.GEQ, ""),
    Cmp(3, Cmp.LIKE, "%(co-production)%"), Cmp(3, Cmp.NOT_LIKE, "%(co-production)%"), Cmp(3, Cmp.LIKE, "a_c"), Cmp(3, Cmp.LIKE, "a%c"),
    Cmp(3, Cmp.LIKE, "%"), Cmp(3, Cmp.LIKE, "_"), Cmp(3, Cmp.LIKE, ""), Cmp(3, Cmp.LIKE, "a.b"), Cmp(3, Cmp.LIKE, "a%b"), Cmp(3, Cmp.LIKE, "_ü%"),
    Cmp(3, Cmp.LIKE, "___"), Cmp(3, Cmp.LIKE, "%b%c"), Cmp(3, Cmp.LIKE, "[1]"), Cmp(3, Cmp.LIKE, "^start$"), Cmp(3, Cmp.LIKE, "50\\%"), Cmp(3, Cmp.LIKE, "%abc"),
    Lop.makeAnd(Cmp(0, Cmp.GT, 10), Cmp(3, Cmp.LIKE, "a%")),
    Lop.makeOr(Cmp(1, Cmp.LT, 0), Cmp(2, Cmp.GT, 2.0)),
    Lop.makeNot(Cmp(0, Cmp.LT, 25)),      # NULL keys become true (bitmap_not)
    Lop.makeNot(Lop.makeOr(Cmp(3, Cmp.IS_NULL), Lop.makeAnd(Cmp(0, Cmp.GEQ, 5), Cmp(2, Cmp.LEQ, 1.0)))),
]


@pytest.mark.skipif(not orc.available("ref"), reason="oracle/_ref/libref_oracle.so not built")
def test_restatement_matches_the_reference_eval():
    """ASCII strings only: offline, the reference library is built with a std::regex stand-in for RE2
    (oracle/shim/re2/re2.h) whose '.' is one BYTE, while RE2 proper -- and the restatement, and the CUDA
    matcher -- take '_' as one UTF-8 code point (RE2's default)."""
    rng = np.random.default_rng(5)
    table, cells = make_table(rng, 3000, ascii_only=True)
    for i, st in enumerate(STATEMENTS):
        want = fo.selected_ref(table, st)
        got = fo.selected(cells, st)
        assert np.array_equal(got, want), f"statement {i}: {np.flatnonzero(got != want)[:5]}"


def test_wrong_literal_kind_raises_like_std_get():
    rng = np.random.default_rng(6)
    _table, cells = make_table(rng, 50)
    with pytest.raises(TypeError):
        fo.selected(cells, Cmp(0, Cmp.EQ, 1.5))
    with pytest.raises(TypeError):
        fo.selected(cells, Cmp(2, Cmp.EQ, 1))
    with pytest.raises(TypeError):
        fo.selected(cells, Cmp(3, Cmp.EQ, 1))


def decoded(table):
    return [c.to_python() for c in orc.decode_table(table, impl=H.encoder_impl())]


def same_cells(a, b):
    """lists of python values per column; NaN == NaN, -0.0 != 0.0 (bit-exact doubles)"""
    if len(a) != len(b):
        return False
    for ca, cb in zip(a, b):
        if len(ca) != len(cb):
            return False
        for x, y in zip(ca, cb):
            if isinstance(x, float) and isinstance(y, float):
                if np.float64(x).view(np.uint64) != np.float64(y).view(np.uint64):
                    return False
            elif x != y:
                return False
    return True


@pytest.fixture(scope="module")
def ctx():
    c = rj.build_context(0)
    yield c
    rj.destroy_context(c)


@pytest.mark.gpu
@pytest.mark.parametrize("n,long_frac", [(5000, 0.0), (70000, 0.0), (3000, 0.02)])
def test_filter_table_matches_the_oracle(ctx, n, long_frac):
    rng = np.random.default_rng(n)
    table, cells = make_table(rng, n, long_frac)
    base = decoded(table)
    for i, st in enumerate(STATEMENTS):
        mask = fo.selected(cells, st)

        got = rj.filter_table(table, st, ctx)
        assert got.num_rows == int(mask.sum()), f"statement {i}"
        assert [int(c.type) for c in got.columns] == [INT32, INT64, FP64, VARCHAR]
        idx = np.flatnonzero(mask)
        want = [[col[j] for j in idx] for col in base]
        assert same_cells(decoded(got), want), f"statement {i}"


@pytest.mark.gpu
def test_filter_without_statement_keeps_every_row_and_empty_selection(ctx):
    rng = np.random.default_rng(1)
    table, cells = make_table(rng, 4000)
    got = rj.filter_table(table, None, ctx)
    assert got.num_rows == 4000 and same_cells(decoded(got), decoded(table))
    none = rj.filter_table(table, Cmp(0, Cmp.GT, 10**6), ctx)
    assert none.num_rows == 0 and [int(c.type) for c in none.columns] == [INT32, INT64, FP64, VARCHAR]
    assert all(c.n_pages == 0 for c in none.columns)


@pytest.mark.gpu
def test_filter_errors_surface(ctx):
    rng = np.random.default_rng(2)
    table, _ = make_table(rng, 100)
    with pytest.raises(rj.EngineError):
        rj.filter_table(table, Cmp(9, Cmp.EQ, 1), ctx)          # column out of range
    with pytest.raises(rj.EngineError):
        rj.filter_table(table, Cmp(0, Cmp.EQ, 1.5), ctx)        # std::get<int64_t> on a double literal
    with pytest.raises(rj.EngineError):
        rj.filter_table(table, Cmp(0, Cmp.LIKE, "a%"), ctx)     # LIKE on a fixed-width column


@pytest.mark.gpu
def test_filtered_scan_feeds_the_join(ctx):
    """the harness's pipeline: filter + emit the base tables, then Contest::execute on the result"""
    rng = np.random.default_rng(3)
    table, cells = make_table(rng, 20000)
    st = Lop.makeAnd(Cmp(0, Cmp.LT, 30), Cmp(3, Cmp.IS_NOT_NULL))
    filtered = rj.filter_table(table, st, ctx)
    types = [INT32, INT64, FP64, VARCHAR]
    plan = H.single_join_plan(filtered, table, types, types, 0, 0, True, out_cols=[0, 3, 5])
    got = rj.execute(plan, ctx)
    want = orc.execute(plan, impl="port")
    assert got.num_rows == want.num_rows and orc.result_equal(got, want)
