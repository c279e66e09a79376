"""GPU tests (-m gpu) of the drop-in itself: `Contest::execute` of radix-join_b200/libcontest_b200.so,
called by a C++ harness (tests/contest_harness.cpp, built against the reference's headers) on Plans
whose inputs are individually `new`-ed pages written by the reference's ColumnInserter -- the data
path a contest harness drives (tests/read_sql.cpp:1224-1249) -- side by side with the UNMODIFIED
reference's `Contest::execute` on the same Plan object.  Results are decoded with the reference's
Table::from_columnar and compared as sorted multisets (tests/read_sql.cpp:1206-1221); the generator's
own statement of the result (a multiset checksum derived without joining) must agree too.

The harness binary is built where /root/reference exists and travels in oracle/_ref/.
"""
import json
import os
import subprocess

import pytest

import helpers as H

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

HARNESS = os.path.join(H.ROOT, "oracle", "_ref", "contest_harness")
needs_harness = pytest.mark.skipif(not os.path.exists(HARNESS), reason="oracle/_ref/contest_harness not built")


def run_harness(*args, env=None, timeout=1500, want_stderr=False):
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run([HARNESS, *args], capture_output=True, text=True, timeout=timeout, env=e)
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert lines, (out.returncode, out.stdout[-2000:], out.stderr[-4000:])
    res = json.loads(lines[-1])
    assert out.returncode == 0 and res["ok"], (res, out.stderr[-2000:])
    return (res, out.stderr) if want_stderr else res


@needs_harness
def test_config1_full_size_matches_the_reference():
    """BASELINE.json configs[0]: 1 M x 10 M INT32 keys, full size on both implementations.  The probe
    column is 5 041 individually allocated pages; the result 2 x 5 041 `new Page`s."""
    res = run_harness("parity", "c1")
    assert res["build_rows"] == 1_000_000 and res["probe_rows"] == 10_000_000
    assert res["rows"] == res["ref_rows"] == 10_000_000


@needs_harness
def test_config1_windowed_with_a_wrapping_staging_ring():
    """same join cut into ~4 MiB row windows, staging rings capped at 2 buffers (1 024 pages): every
    ring wraps many times and upload / kernels / download of different windows overlap"""
    res = run_harness("parity", "c1", "--div", "2", env={"RJ_WINDOW_BYTES": str(4 << 20), "RJ_PIPE_BUFS": "2"})
    assert res["rows"] == res["ref_rows"] == 5_000_000


@needs_harness
def test_config2_sample_matches_the_reference():
    """config 2 at 1/64 scale (1 Mi x 8 Mi, Zipf(0.75), INT64 + FP64 payloads, 1 % NULLs): the S.b
    column alone is > 8 192 pages"""
    res = run_harness("parity", "c2", "--div", "64")
    assert res["rows"] == res["ref_rows"] == (1 << 29) // 64
    assert res["input_pages"] > 8192


@needs_harness
def test_config2_sample_windowed():
    res = run_harness("parity", "c2", "--div", "128", env={"RJ_WINDOW_BYTES": str(8 << 20), "RJ_PIPE_BUFS": "3"})
    assert res["rows"] == res["ref_rows"] == (1 << 29) // 128


@needs_harness
def test_config2_sample_on_a_device_group_matches_the_reference():
    """RJ_GPUS=2: Contest::build_context opens a device group (rj_ctx_create_multi) and the join runs on both
    GPUs -- row slices per device, scatter pass 1 local, pass 2 pulling its regions over NVLink, result page
    lists appended in device order -- against the unmodified reference on the same Plan (1/16 scale: 4 Mi x 32 Mi,
    a two-pass join)"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    res, err = run_harness("parity", "c2", "--div", "16", env={"RJ_GPUS": "2", "RJ_TRACE": "1"}, want_stderr=True)
    assert res["rows"] == res["ref_rows"] == (1 << 29) // 16
    assert "[rj multi] 2 devices" in err, err[-2000:]  # the join did run on the group, not on device 0 alone
