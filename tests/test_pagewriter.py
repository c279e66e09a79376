"""Harness-side page writer on the device (ColumnInserter, include/plan.h:151-335): pages
written by rj_gen_fixed_pages and by rj_varchar_descriptors + rj_encode_varchar_* from dense values must
decode -- with the UNMODIFIED reference's Table::from_columnar where its library travelled, and with the C
port -- to exactly the values that went in: NULLs, empty strings, -0.0 / NaN bit patterns, strings of
8185 / 8186 / 20000 bytes (the long-string boundary, src/build_table.cpp:603-619,644-648)."""
import numpy as np
import pytest

import helpers as H
from helpers import FP64, INT32, INT64, VARCHAR, orc, rj
from radix_join_b200 import pagewriter as pw

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = rj.build_context(0)
    yield c
    rj.destroy_context(c)


def impls():
    return [i for i in ("ref", "port") if orc.available(i)]


@pytest.mark.parametrize("dtype,n,null_frac", [(INT32, 1, 0.0), (INT32, 1984, 0.0), (INT32, 1985, 0.3), (INT32, 100_003, 0.5),
                                               (INT64, 1007, 0.0), (INT64, 1008, 0.2), (INT64, 70_001, 1.0),
                                               (FP64, 5000, 0.1), (FP64, 300_000, 0.01)])
def test_fixed_pages_decode_to_the_input(ctx, dtype, n, null_frac):
    rng = np.random.default_rng(n)
    cells = H.random_cells(rng, dtype, n, null_frac=null_frac)
    if dtype == FP64:
        cells.values[: min(n, 4)] = [-0.0, np.nan, np.inf, 5e-324][: min(n, 4)]
    pages = pw.fixed_pages(ctx, cells.values, cells.valid.astype(bool), dtype)
    for impl in impls():
        back = orc.decode(rj.Column(dtype, pages), n, impl=impl)
        assert np.array_equal(back.valid, cells.valid), impl
        assert np.array_equal(back.bits(), cells.bits()), impl  # bit patterns, NULL rows forced to 0


def test_varchar_pages_decode_to_the_input(ctx):
    rng = np.random.default_rng(7)
    strings = []
    for i in range(30_000):
        r = rng.random()
        if r < 0.2:
            strings.append(None)
        elif r < 0.25:
            strings.append(b"")
        else:
            strings.append(bytes(rng.integers(32, 127, int(rng.integers(1, 120)), dtype=np.uint8)))
    for at, ln in ((5, 8185), (6, 8186), (7, 20000), (100, 1021), (101, 1020), (29_999, 40000)):
        strings[at] = bytes(rng.integers(97, 123, ln, dtype=np.uint8))
    cells = orc.Cells.from_strings(strings)
    lens = (cells.str_off[1:] - cells.str_off[:-1]).astype(np.int64)
    pages = pw.varchar_pages(ctx, lens, cells.chars, cells.valid.astype(bool))
    for impl in impls():
        back = orc.decode(rj.Column(VARCHAR, pages), len(strings), impl=impl)
        assert back.to_python() == strings, impl


def test_all_null_and_empty_columns(ctx):
    assert pw.fixed_pages(ctx, np.zeros(0, np.int32), None, INT32).shape == (0, 8192)
    assert pw.varchar_pages(ctx, np.zeros(0, np.int64), np.zeros(0, np.uint8), None).shape == (0, 8192)
    n = 70_000  # more all-NULL rows than one page's bitmap holds
    pages = pw.varchar_pages(ctx, np.zeros(n, np.int64), np.zeros(0, np.uint8), np.zeros(n, bool))
    for impl in impls():
        back = orc.decode(rj.Column(VARCHAR, pages), n, impl=impl)
        assert not back.valid.any()


def test_job_generator_on_device_matches_the_host_generator(ctx):
    """radix_join_b200.job writes its inputs with this page writer when a context is passed: same rows"""
    from radix_join_b200 import job
    host = job.make_inputs("13a", scale=0.01, seed=3, long_strings=True)
    dev = job.make_inputs("13a", scale=0.01, seed=3, long_strings=True, ctx=ctx)
    assert host.keys() == dev.keys()
    for alias in host:
        a, b = host[alias], dev[alias]
        assert a.num_rows == b.num_rows
        for ca, cb in zip(a.columns, b.columns):
            assert ca.type == cb.type and (ca.n_pages == 0) == (cb.n_pages == 0)
            if ca.n_pages:
                da = orc.decode(ca, a.num_rows, impl="port")
                db = orc.decode(cb, b.num_rows, impl="port")
                assert da.to_python() == db.to_python(), alias
