"""The N > 1 data plane under pytest: spawns tests/multi_gpu_check.py with torchrun on 2 (and, if present, all) GPUs of
the box — every rank a shard of the window, H | b | cost summed inside the fused pass kernel over NVLink peer memory —
checks the all-reduced system and the redundant solves against the full-window oracle and that the ranks end
bit-identical.  Skipped on a 1-GPU box (the round-end GPU tier); bench.py's N > 1 lines carry the same parity keys."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def n_gpus():
    import torch
    return torch.cuda.device_count()


def run_check(n, env_extra=None, port=29517):
    env = dict(os.environ)
    env.update(env_extra or {})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0 and "MULTI_GPU_CHECK PASS" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
    return out.stdout


@pytest.mark.skipif("n_gpus() < 2")
def test_two_ranks_fused_exchange():
    out = run_check(2)
    assert "fused=True" in out and "exchange=kernel" in out, out


@pytest.mark.skipif("n_gpus() < 2")
def test_two_ranks_per_stream_kernels_and_nccl_fallbacks():
    run_check(2, {"CLIC_FUSED": "0"}, port=29518)          # round-1 path: per-stream kernels + one-shot all-reduce kernel
    run_check(2, {"CLIC_NO_P2P": "1"}, port=29519)          # fused pass + ncclAllReduce of the dense system


@pytest.mark.skipif("n_gpus() < 4")
def test_all_ranks_fused_exchange():
    run_check(n_gpus(), port=29520)
