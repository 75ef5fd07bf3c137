"""Multi-GPU parity check, run under torchrun on N GPUs of one box (one process per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py
Block-partitioned ellip / elast / stokes MatMults and embedded-Q1 rows through the C ABI, NCCL ghost exchange inside, against the
single-domain oracle (tests/partition_parity.py).  tests/test_gpu_partition.py launches this under pytest when the box has >= 2 GPUs;
bench.py --gpus N prints the same numbers in its "parity" block."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch
import torch.distributed as dist

import partition_parity
from dohp_b200 import lib

rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lrank)
dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    idt = torch.frombuffer(bytearray(lib.Context.nccl_unique_id()), dtype=torch.uint8).cuda()
dist.broadcast(idt, 0)
ctx = lib.Context(lrank, rank, world, bytes(idt.cpu().numpy().tobytes()), stream=torch.cuda.current_stream().cuda_stream)
procs = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[world]
res = partition_parity.run_all(ctx, dist, rank, world, procs)
dist.barrier()
if rank == 0:
    print("partition parity (world %d): %s" % (world, res), flush=True)
    assert all(v <= 1e-12 for v in res.values()), res
    print("MGPU PARITY OK (world %d): MatMult ellip %.2e elast %.2e stokes %.2e, Q1 rows %.2e" % (world, res["ellip"], res["elast"], res["stokes"], res["q1"]))
dist.destroy_process_group()
