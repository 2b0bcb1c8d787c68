"""Multi-process GPU tests of the row-sharded training path: two ranks, one GPU each, NCCL.

Launched with torchrun from inside the test (the pytest process itself never joins a process group); skipped when the
box has fewer than two GPUs.  The worker checks, on the CUDA path:
  * row-sharded training through the LIBRARY-owned communicator (vqb_comm_init_rank, VQB_TRAIN_USE_COMM) and through
    the host callback (torch.distributed) give the same codebooks on every rank, bit for bit;
  * those codebooks agree with single-GPU training and with the CPU oracle within 1e-4 relative per subspace
    (row-sharded sums cannot be bit-exact with a sequential f32 sum);
  * iteration counts agree, empty-cluster re-seeding across ranks included (a row owned by the other rank)."""
import os
import signal
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_gpu_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("world", [2])
def test_row_sharded_training_two_ranks(world):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "check_row_shard.py")]
    # own process group + hard deadline: a rank that fails leaves its peer waiting inside a collective
    proc = subprocess.Popen(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, start_new_session=True)
    try:
        out, err = proc.communicate(timeout=420)
    except subprocess.TimeoutExpired:
        os.killpg(proc.pid, signal.SIGKILL)
        out, err = proc.communicate()
        sys.stdout.write(out[-3000:]); sys.stderr.write(err[-3000:])
        pytest.fail("row-shard check did not finish in 420 s (a rank failed or hung)")
    sys.stdout.write(out[-3000:])
    sys.stderr.write(err[-3000:])
    assert proc.returncode == 0
    assert "ROW-SHARD OK" in out
