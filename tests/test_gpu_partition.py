"""Partitioned-space parity under pytest: world 1 in-process (1x1x1 grid, the same code path minus NCCL) and, when the box has at
least two GPUs, tests/mgpu_check.py under torchrun on 2 (and 4) GPUs with the NCCL ghost exchange (ellip, elast, stokes MatMults
with their two-space halos, Q1 rows summed over the ranks) against the single-domain oracle."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_parity_world1():
    import torch
    import partition_parity
    from dohp_b200 import lib
    ctx = lib.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    res = partition_parity.run_all(ctx, None, 0, 1, (1, 1, 1))
    ctx.close()
    assert all(v <= 1e-12 for v in res.values()), res


@pytest.mark.parametrize("world", [2, 4])
def test_partition_parity_nccl(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs on the box (bench.py --gpus N prints the same parity block for the driver's scaling runs)" % world)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_check.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "MGPU PARITY OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
