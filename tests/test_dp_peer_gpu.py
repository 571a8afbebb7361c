"""Peer-memory data-parallel optimizer step vs the single-GPU fused step (needs >= 2 GPUs; skipped otherwise)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("precision,batch", [("fp32", 32), ("bf16", 32), ("bf16", 128)])
def test_dp_peer_matches_single_gpu(precision, batch):
    """batch 128 in bf16 mode exercises the fused push (wgrad epilogue -> owners' staging slots); batch 32 the fallback."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tests", "dp_peer_check.py"), precision]
    env = dict(os.environ, FND_DP_CHECK_BATCH=str(batch))
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600, env=env)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0


def test_dp_trainer_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29519", os.path.join(ROOT, "tests", "dp_trainer_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0


def test_dp_sequence_trainer_two_gpus():
    """Sequence front-end (Tier B, self-oracle scope): 2-rank SequenceTrainer vs a single-GPU replica on the whole batch."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29523", os.path.join(ROOT, "tests", "dp_seq_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0
