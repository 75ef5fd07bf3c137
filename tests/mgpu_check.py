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
# configs[3] as a solve: the same global Stokes problem on the partition and (rank 0, CPU oracle) on a single domain: identical
# Newton / FGMRES iteration counts (the preconditioners are partition independent), same residual history to 1e-6
import dist_krylov
import oracle_lib as ol
ne1 = (2, 2, 2)
r = dist_krylov.run_stokes_gpu(ctx, lib, ol, dist, rank, world, procs, ne_per_gpu=ne1, p=3)
dist.barrier()
if rank == 0:
    ol.oracle().or_set_num_threads(8)
    ro = dist_krylov.run_stokes_oracle(lib, ol, ne=tuple(a * b for a, b in zip(ne1, procs)), p=3)
    print("stokes Newton-Krylov world %d: newton %d outer %s inner %s |F| %.3e err %.2e ; single-domain oracle: newton %d outer %s inner %s" %
          (world, r["newton_its"], r["outer"], r["inner_total"], r["fnorms"][-1], r["velocity_err_rel"], ro["newton_its"], ro["outer"], ro["inner_total"]), flush=True)
    # identical counts while the nonlinear residual is well above the noise of the inner solves (they stop on a tolerance and the GPU sums
    # in another order); in the last steps a Krylov solve may take an iteration more or less (as in tests/test_gpu_partition.py)
    assert abs(r["newton_its"] - ro["newton_its"]) <= 1, (r, ro)
    early = [k for k, f in enumerate(ro["fnorms"][:-1]) if f > 1e-3 * ro["fnorms"][0] and k < len(r["outer"])]
    assert len(early) >= 3 and [r["outer"][k] for k in early] == [ro["outer"][k] for k in early], (r, ro)
    assert all(abs(a - b) <= 2 for a, b in zip(r["outer"], ro["outer"])), (r, ro)
    # the inner CG solves stop on a tolerance: one iteration more or less in one of them moves the Newton iterates by O(1e-5)
    assert all(abs(a - b) <= 1e-3 * ro["fnorms"][0] for a, b in zip(r["fnorms"], ro["fnorms"])), (r["fnorms"], ro["fnorms"])
dist.barrier()
if rank == 0:
    print("partition parity (world %d): %s" % (world, res), flush=True)
    assert all(v <= 1e-12 for v in res.values()), res
    print("MGPU PARITY OK (world %d): MatMult ellip %.2e elast %.2e stokes %.2e, Q1 rows %.2e" % (world, res["ellip"], res["elast"], res["stokes"], res["q1"]))
dist.destroy_process_group()
