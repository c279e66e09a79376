"""Host-side logic that needs no GPU: the launch sizing of the fused root join (grid, emitting warps, page
budget -- radix-join_b200/csrc/rj_internal.h), checked by a small C++ program compiled against the engine's own
header."""
import os
import shutil
import subprocess

import pytest

import helpers as H

CUDA_INC = "/usr/local/cuda/include"


@pytest.mark.skipif(shutil.which("g++") is None or not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")),
                    reason="needs g++ and the CUDA headers")
def test_fused_join_launch_sizing(tmp_path):
    exe = str(tmp_path / "launch_sizing")
    src = os.path.join(H.ROOT, "tests", "host", "launch_sizing.cpp")
    out = subprocess.run(["g++", "-std=c++17", "-O1", "-I", CUDA_INC, src, "-o", exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-3000:]
    run = subprocess.run([exe], capture_output=True, text=True)
    assert run.returncode == 0 and run.stdout.strip().endswith("ok"), run.stdout[-3000:]
    # the tuning knob is honoured (and cannot break the page budget)
    run = subprocess.run([exe], capture_output=True, text=True, env=dict(os.environ, RJ_EMIT_MIN_CHUNKS="1"))
    assert run.returncode == 0, run.stdout[-3000:]
