"""N-rank reduced image == 1-GPU image of the same frame indices, on real GPUs over NCCL (tests/multi_gpu_check.py under torchrun).
Needs at least two GPUs on the box; the single-GPU driver run skips it, `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`
runs it.  The same reduce identity is also asserted inside every `bench.py --gpus N` run (reduce_check)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_partition_equals_single_gpu():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs two GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "MULTI_GPU_CHECK_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
