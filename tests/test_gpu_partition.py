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


def test_stokes_newton_krylov_counts_match_oracle_world1():
    """BASELINE.json configs[3] in small on one GPU: Newton + FGMRES + Schur preconditioner (tests/dist_krylov.py) driven by the library
    and by the CPU oracle: identical Newton and FGMRES iteration counts, same residual history, converges to the manufactured solution."""
    import torch
    import dist_krylov
    import oracle_lib as ol
    from dohp_b200 import lib
    ol.build_oracle(ref=False)
    ol.oracle().or_set_num_threads(8)
    ctx = lib.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    r = dist_krylov.run_stokes_gpu(ctx, lib, ol, None, 0, 1, (1, 1, 1), ne_per_gpu=(3, 3, 3), p=3)
    ro = dist_krylov.run_stokes_oracle(lib, ol, ne=(3, 3, 3), p=3)
    ctx.close()
    # identical counts while the nonlinear residual is well above the noise of the inner solves; in the last steps (relative residual
    # below 1e-3) a Krylov solve that stops on a tolerance may take an iteration more or less on the GPU (different summation order)
    assert abs(r["newton_its"] - ro["newton_its"]) <= 1, (r, ro)
    early = [k for k, f in enumerate(ro["fnorms"][:-1]) if f > 1e-3 * ro["fnorms"][0] and k < len(r["outer"])]
    assert len(early) >= 3 and [r["outer"][k] for k in early] == [ro["outer"][k] for k in early], (r, ro)
    assert all(abs(a - b) <= 2 for a, b in zip(r["outer"], ro["outer"])), (r, ro)
    # the inner CG solves stop on a tolerance: one iteration more or less in one of them moves the Newton iterates by O(1e-5)
    assert all(abs(a - b) <= 1e-3 * ro["fnorms"][0] for a, b in zip(r["fnorms"], ro["fnorms"])), (r["fnorms"], ro["fnorms"])
    assert r["fnorms"][-1] <= 1e-7 * r["fnorms"][0] and r["velocity_err_rel"] < 1e-2


def test_stokes_fgmres_closing_lines_at_the_gpu_solution():
    """stokes-q3q2-fgmres.refout, last four lines (stokes.c:1152-1170): the configs[3] problem at the refout's own size (genblock 4x4x4,
    velocity degree 3, pressure degree 2, power-law rheology) solved through the C ABI; the twelve printed error numbers belong to the
    discrete solution and must come out of the GPU solve (the oracle evaluates the error integrals: tests/test_stokes_closing_lines.py)."""
    import numpy as np
    import torch
    import dist_krylov
    import oracle_lib as ol
    from dohp_b200 import lib
    from projection_common import printed_equal
    from test_stokes_closing_lines import closing_lines
    ol.build_oracle(ref=False)
    ol.oracle().or_set_num_threads(8)
    ctx = lib.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    r = dist_krylov.run_stokes_gpu(ctx, lib, ol, None, 0, 1, (1, 1, 1), ne_per_gpu=(4, 4, 4), p=3, newton_rtol=1e-8, return_solution=True)
    ctx.close()
    o, xu, xp = r["oracle"], r["xu"], r["xp"]
    xp = xp - xp.mean()
    su, spn = o.nodal_solution()
    spn = spn - spn.mean()
    en, gn, pn = o.error_norms(xu, xp)
    mine = {"Interpolation residual": ol.norms_algebraic_scaled(np.concatenate([xu - su, xp - spn])), "Pointwise velocity error": en,
            "Pointwise gradient error": gn, "Pointwise pressure error": pn}
    gold = closing_lines("stokes-q3q2-fgmres")
    for label, vals in mine.items():
        for v, pr in zip(vals, gold[label]):
            assert printed_equal(v, pr, "%8.2e"), (label, "%8.2e" % v, pr)


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
