"""Frame sharding over 2 GPUs against the single-GPU run: needs two visible GPUs (skipped on a 1-GPU box).
Bounds: the only arithmetic difference is the fp32 re-association of the all-reduced GroupNorm sums, amplified by the sampler like any
fp16 perturbation (same style of bound as tests/test_gpu_pipeline.py)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="frame sharding needs >= 2 GPUs")
@pytest.mark.parametrize("index", ["gather", "const"])
def test_frame_sharded_matches_single_gpu(index, report):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517" if index == "gather" else "29518", os.path.join(ROOT, "tools", "shard_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, FZ_SHARD_INDEX=index))
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert line, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.loads(line[-1])
    report["frame_shard_2gpu"] = res
    assert r.returncode == 0 and res["ok"], res
