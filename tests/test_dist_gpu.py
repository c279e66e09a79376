"""GPU test (-m gpu, needs >= 2 GPUs): the multi-GPU join under torchrun -- every rank joins its shard
of a scaled config 2 through radix_join_b200.dist_join (fused partition + exchange over NVLink, then the
NCCL collective exchange, then the broadcast path); rank 0 compares the appended result pages with the
single-GPU engine AND the CPU oracle on the unsharded tables (tests/dist_gpu_check.py)."""
import os
import subprocess
import sys

import pytest

import helpers as H

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_two_gpu_join_matches_single_gpu_and_oracle(exchange, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, RJ_DIST_EXCHANGE=exchange, RJ_CHECK_DIR=str(tmp_path), MASTER_ADDR="127.0.0.1")
    port = "29541" if exchange == "p2p" else "29542"
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", port,
                          os.path.join(H.ROOT, "tests", "dist_gpu_check.py")],
                         capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, (out.stdout[-3000:], out.stderr[-3000:])
    assert out.stdout.count("-> OK") >= 2 and "MISMATCH" not in out.stdout
