"""Multi-GPU data path on hardware (needs >= 2 visible GPUs; skipped on a 1-GPU box): two ranks render their interleaved
row blocks, NCCL all-gathers the image, and bench.py's verification re-renders row blocks of the OTHER rank as ordinary
tiles on rank 0 and compares them bit for bit (N-GPU image == 1-GPU image)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("workload,extra", [("cornell", ["--width", "256", "--height", "256", "--spp", "4"]),
                                            ("soup1m", ["--width", "512", "--height", "512", "--spp", "2"])])
def test_two_ranks_allgather_is_the_single_gpu_image(workload, extra):
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "2", "--warmup", "1",
           "--workload", workload, "--no-cpu-baseline"] + extra
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["n_gpus"] == 2 and out["allgather_verified"] is True, out.get("allgather_check")
    assert out["allgather_check"]["rows_checked"] >= 8 and out["allgather_check"]["foreign_rows_bit_identical"] is True
    assert len(out["ranks"]["frames_ms"]) == 2 and out["e2e"]["value"] > 0
