"""Multi-GPU parity on hardware (needs >= 2 GPUs; skipped otherwise): the read-sharded run over NCCL equals the single-GPU run."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_nccl_allreduce_matches_single_gpu():
    from dbgphmm_b200 import hmmv2 as H
    if H.device_count() < 2:
        pytest.skip("needs two GPUs (run under `gpurun --gpus 2`)")
    env = dict(os.environ); env.pop("DBGPHMM_STRATEGY", None)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(ROOT, "tests", "nccl_worker.py")], capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(" ok") == 2
