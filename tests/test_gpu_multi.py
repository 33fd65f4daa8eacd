"""Two ranks over NCCL on a box with at least two GPUs (skipped elsewhere; `gpurun --gpus 2`):
the sharded front end of bench.py -- faces dealt by centroid z (sb_shard_*), one all_reduce per
step -- must hand back, after the gather, byte for byte what the single-GPU front end computes
(sorted pair lists, segments, per-face flags).  The comparison itself runs
inside bench.py (rank 0 runs the single-GPU front end in the same process); this test launches
it the way the driver does and reads the verdict from the JSON line."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.slow
@pytest.mark.parametrize("config", ["c2", "c3"])
def test_two_ranks_nccl_equal_single_gpu(config):
    if _gpus() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29613", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "3", "--warmup", "3",
           "--config", config, "--no-cpu-baseline"]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["n_gpus"] == 2
    p = d["multi_gpu"]["parity_vs_single_gpu"]
    assert p["candidate_pairs_equal"] and p["hit_pairs_identical"] and p["segments_bit_identical"] and p["flags_identical"], p
    assert sum(s["fallbacks"] for s in d["multi_gpu"]["shards"]) >= 0
