"""Spawns tests/dp_check.py under torchrun on all visible GPUs (needs >= 2; skipped otherwise):
shard parity with the full-batch oracle, the fused wgrad + peer-memory all-reduce, cross-rank batchnorm."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

import mygrad_b200 as mg
from mygrad_b200._lib import lib
import ctypes as C

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _gpu_count() -> int:
    n = C.c_int(0)
    return n.value if lib.mgb_device_count(C.byref(n)) == 0 else 0


@pytest.mark.parametrize("peer", ["1", "0"])
def test_data_parallel_on_all_gpus(peer):
    n = min(_gpu_count(), 8)
    if n < 2:
        pytest.skip("needs at least two GPUs")
    env = dict(os.environ, MGB_DP_PEER=peer)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29611" if peer == "1" else "29612",
           str(ROOT / "tests" / "dp_check.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DP_CHECK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
