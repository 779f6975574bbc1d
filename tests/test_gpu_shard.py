"""Frames of ONE clip sharded over the GPUs of this box, non-identity temporal layers, peer-memory exchange
(fatezero_b200/csrc/fz_p2p.cu): tools/shard_check.py under torchrun — sharded forward vs single GPU, the reference's golden cases on the
ranks' frames, CUDA-graph replay of the sharded loops.  Needs >= 2 visible GPUs (skipped on a 1-GPU box)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="frame sharding needs >= 2 GPUs")
@pytest.mark.parametrize("world", [2, 4])
def test_frame_sharded(world, report):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29517 + world), os.path.join(ROOT, "tools", "shard_check.py")] + (["--big"] if world == 2 else [])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert line, r.stdout[-2000:] + r.stderr[-3000:]
    res = json.loads(line[-1])
    report[f"frame_shard_{world}gpu"] = res
    assert r.returncode == 0 and res["ok"], res
