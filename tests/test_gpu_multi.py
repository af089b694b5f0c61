"""GPU suite (-m gpu), N > 1: spawns torchrun over min(device_count, 8) GPUs of this box and runs tools/check_multi_gpu.py --
read-sharded pools with the NCCL accumulator all-reduce against the reference's golden accumulators, and the sample-sharded
VarStats reduce (gtb_allreduce_varstats) against the host merge.  Skipped on a single-GPU box."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_sharded_parity_under_torchrun():
    import torch
    n = min(torch.cuda.device_count(), 8)
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tools", "check_multi_gpu.py")], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "multi-GPU sharded parity: PASS" in r.stdout, r.stdout[-3000:]
