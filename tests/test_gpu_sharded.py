"""One proof over several GPUs (BASELINE.json configs[4] shape): the sharded prover (spartan_b200/csrc/comm.cu + the sharded
paths of prover.cpp / snark.cpp) must return, on every rank, the very bytes the single-GPU prover returns — and those are diffed against the
oracle.  Needs >= 2 GPUs on the box (skipped on the single-GPU box); launched as `torchrun --nproc-per-node N tools/run_sharded.py`.
The 8-GPU 2^22 run of configs[4] is the same script with `--logn 22 --golden tests/golden/snark_proof_sha256.json` (profiles/r02_sharded.md)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _run(nproc, extra, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tools", "run_sharded.py")] + extra
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    return [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]


@pytest.mark.skipif(_gpus() < 2, reason="needs at least 2 GPUs")
def test_sharded_snark_bytes_equal_single_gpu_and_oracle():
    n = 8 if _gpus() >= 8 else (4 if _gpus() >= 4 else 2)
    lines = _run(n, ["--logn", "16", "18", "--oracle", "--golden", os.path.join(ROOT, "tests", "golden", "snark_proof_sha256.json"), "--reps", "1"], 29541)
    assert len(lines) == 2
    for l in lines:
        assert l["bytes_identical_to_single_gpu_on_every_rank"] and l["bytes_identical_to_oracle"] and l["matches_golden_fixture"], l


@pytest.mark.skipif(_gpus() < 2, reason="needs at least 2 GPUs")
def test_sharded_nizk_bytes_equal_single_gpu_and_oracle():
    lines = _run(2, ["--logn", "16", "--oracle", "--nizk", "--reps", "1"], 29542)
    assert lines and lines[0]["bytes_identical_to_single_gpu_on_every_rank"] and lines[0]["bytes_identical_to_oracle"], lines
