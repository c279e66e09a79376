"""Result validation on the device (the harness's sorted-multiset `compare`,
tests/read_sql.cpp:1159-1222): rj_tables_equal must agree with the CPU comparer (oracle.result_equal) -- equal
for any row permutation and any page packing, unequal as soon as ONE cell differs (a NULL flipped, a value bit,
a string byte, a duplicate count)."""
import numpy as np
import pytest

import helpers as H
from helpers import FP64, INT32, INT64, VARCHAR, orc, rj

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = rj.build_context(0)
    yield c
    rj.destroy_context(c)


def make_cells(rng, n, long_frac=0.0):
    return [H.random_cells(rng, INT32, n, null_frac=0.1, key_range=40),          # many duplicate rows
            H.random_cells(rng, INT64, n, null_frac=0.2),
            H.random_cells(rng, FP64, n, null_frac=0.1),
            H.random_cells(rng, VARCHAR, n, null_frac=0.15, max_len=30, long_frac=long_frac)]


def permuted(cells, perm):
    out = []
    for c in cells:
        if c.type == VARCHAR:
            strs = c.to_python()
            out.append(orc.Cells.from_strings([strs[i] for i in perm]))
        else:
            out.append(orc.Cells(c.type, c.valid[perm], values=c.values[perm]))
    return out


@pytest.mark.parametrize("n,long_frac", [(1, 0.0), (5000, 0.0), (150_000, 0.0), (3000, 0.01)])
def test_permutations_are_equal_and_single_cell_changes_are_not(ctx, n, long_frac):
    rng = np.random.default_rng(n)
    cells = make_cells(rng, n, long_frac)
    a = H.table_from_cells(cells)
    b = H.table_from_cells(permuted(cells, rng.permutation(n)))
    eq, bad = rj.tables_equal(a, b, ctx)
    assert eq and bad == 0 and orc.result_equal(a, b)
    if n < 2:
        return
    # one cell changed at a time
    for col in range(4):
        mod = permuted(cells, rng.permutation(n))
        c = mod[col]
        i = int(rng.integers(0, n))
        if c.type == VARCHAR:
            strs = c.to_python()
            strs[i] = (strs[i] or b"") + b"x"
            mod[col] = orc.Cells.from_strings(strs)
        elif c.valid[i]:
            c.values[i] = c.values[i] + 1 if c.type != FP64 else -c.values[i] - 1.0
        else:
            c.valid[i] = 1
        t = H.table_from_cells(mod)
        eq, bad = rj.tables_equal(a, t, ctx)
        assert not eq and bad > 0, col
        assert not orc.result_equal(a, t)


def test_shape_mismatches_and_empty_tables(ctx):
    rng = np.random.default_rng(3)
    cells = make_cells(rng, 100)
    a = H.table_from_cells(cells)
    assert rj.tables_equal(a, H.table_from_cells(cells[:3]), ctx)[0] is False          # column count
    assert rj.tables_equal(a, H.table_from_cells(make_cells(rng, 99)), ctx)[0] is False  # row count
    e = H.empty_table([INT32, VARCHAR])
    assert rj.tables_equal(e, H.empty_table([INT32, VARCHAR]), ctx)[0] is True
    assert rj.tables_equal(e, H.empty_table([INT32, INT64]), ctx)[0] is False           # column type


def test_join_result_validates_against_the_oracle_on_device(ctx):
    """what the harness does with it: the engine's result vs the reference's, compared on the GPU"""
    rng = np.random.default_rng(4)
    tl, _ = H.random_table(rng, [INT32, VARCHAR], 4000, key_cols=(0,), key_range=1500)
    tr, _ = H.random_table(rng, [INT64, INT32], 30000, key_cols=(1,), key_range=1500)
    plan = H.single_join_plan(tl, tr, [INT32, VARCHAR], [INT64, INT32], 0, 1, True, out_cols=[1, 0, 2])
    got = rj.execute(plan, ctx)
    want = orc.execute(plan, impl="port")
    eq, bad = rj.tables_equal(got, want, ctx)
    assert eq and bad == 0
