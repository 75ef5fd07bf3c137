#!/usr/bin/env python
"""Benchmark of the matrix-free MatMult (BASELINE.json: "matrix-free MatMult GDoF/s (FP64, p=2-10)").

Workload (configs[1]): linear Poisson (ellip, p=2, lambda=0) on a 64^3-element genblock box per GPU, polynomial degree
sweep p = 2..10.  One "step" = one MatMult (dB200OpMult = EllipShellMatMult, ellip.c:442-472) at every degree of the
sweep; value = (global dofs summed over the sweep and over all GPUs) / (time of the step), device-resident vectors.
e2e = the same step through the C ABI with HOST (pinned) vectors: H2D of x and D2H of y inside the timed region.
N > 1: block partition of a (64M x 64N x 64P)-element box over an M x N x P process grid as genblock writes it
(weak scaling, 64^3 elements per GPU), ghost exchange over NCCL inside every MatMult.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--ne 64] [--degrees 2,3,...]

--impl reference times the CPU restatement of the reference's path (oracle/, all host threads) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np


def flops_per_element(n):
    """Algorithmic flops of one ellip (bs=1, lambda=0) element apply, SURVEY.md 8(d)."""
    return 36 * n ** 4 + 131 * n ** 3 + 72 * n ** 2 + 144 * n


def bytes_per_element(p):
    """Algorithmic HBM bytes: read x + write y once per dof, 8 vertices, one index per expanded node."""
    return 16 * p ** 3 + 192 + 4 * (p + 1) ** 3


def proc_grid(n):
    return {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[n]


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        try:
            self.proc.terminate()
        except Exception:
            pass
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def physical_cores():
    try:
        import psutil
        return psutil.cpu_count(logical=False) or os.cpu_count() or 1
    except Exception:
        return os.cpu_count() or 1


def cpu_reference_run(degrees, threads, budget_s):
    """The reference's CPU path on a bounded sample of the same workload: for every degree of the sweep a small box, repeated
    EllipShellMatMult (ellip.c:442-472).  When oracle/_ref ships (the reference's own dJacobi layer compiled unmodified,
    oracle/Makefile) the per-element work is the REFERENCE's dEFSApply / dRuleComputePhysical driven by the restated iterator
    loop (oracle/ref_driver.c RefEllipMatMult, kind "reference"); otherwise the plain-C port (oracle/oracle.c, kind "port").
    The mesh comes from the numpy space builder (dohp_b200/box.py): this arm never maps the CUDA library.
    Returns dict(dofs, seconds, sample, kind, single_thread=(dofs, seconds))."""
    import oracle_lib as ol
    from dohp_b200 import box
    ol.build_oracle(ref=False)
    kind = "reference" if ol.have_ref() else "port"
    total_dofs, total_s, sizes, dofs1, secs1 = 0, 0.0, [], 0, 0.0
    per_deg = budget_s / len(degrees)
    for p in degrees:
        ne = 16 if p <= 4 else 12 if p <= 7 else 8
        space = box.build_box_space((ne, ne, ne), p)
        x = np.sin(0.37 * np.arange(space.n))
        xloc = np.concatenate([x, np.zeros(space.nloc - space.n)])        # homogeneous Dirichlet part, fs.c:562-571
        if kind == "reference":
            run = ol.RefEllipMult(space)
            mult = lambda nt: run.mult_local(xloc, nt)
        else:
            o = ol.OracleEllip(space, p=2.0, lam=0.0)
            o.stash[:, :, 0] = 1.0                # linear Poisson store: eta = 1, s = 0
            def mult(nt, o=o):
                ol.oracle().or_set_num_threads(nt)
                return o.mult_local(xloc)
        mult(threads)                           # warm-up
        reps, t0 = 0, time.perf_counter()
        while True:
            mult(threads)
            reps += 1
            dt = time.perf_counter() - t0
            if dt > per_deg or reps >= 500:
                break
        total_dofs += space.n
        total_s += dt / reps
        # the faithful single-process number (the reference is single-threaded per MPI rank): one product on a 6^3 (p <= 7) / 4^3 box
        ne1 = 6 if p <= 7 else 4
        sp1 = box.build_box_space((ne1, ne1, ne1), p) if ne1 != ne else space
        xl1 = np.concatenate([np.sin(0.37 * np.arange(sp1.n)), np.zeros(sp1.nloc - sp1.n)])
        if kind == "reference":
            r1 = ol.RefEllipMult(sp1)
            m1 = lambda: r1.mult_local(xl1, 1)
        else:
            o1 = ol.OracleEllip(sp1, p=2.0, lam=0.0)
            o1.stash[:, :, 0] = 1.0
            def m1(o1=o1):
                ol.oracle().or_set_num_threads(1)
                return o1.mult_local(xl1)
        m1()
        t0 = time.perf_counter()
        m1()
        secs1 += time.perf_counter() - t0
        dofs1 += sp1.n
        sizes.append("%d^3@p%d" % (ne, p))
    sample = ("one EllipShellMatMult per degree on boxes " + ",".join(sizes) + " (DoF/s is size-independent for this path); per-element kernels = "
              + ("the reference's own dEFSApply/dRuleComputePhysical (oracle/_ref, compiled unmodified)" if kind == "reference" else "the plain-C port oracle/oracle.c"))
    return dict(dofs=total_dofs, seconds=total_s, sample=sample, kind=kind, single_thread=(dofs1, secs1))


def cpu_baseline_json(r, threads, unit):
    return {"value": r["dofs"] / r["seconds"] / 1e9, "unit": unit, "cores": threads, "physical_cores": physical_cores(), "kind": r["kind"], "sample": r["sample"],
            "single_thread_value": r["single_thread"][0] / r["single_thread"][1] / 1e9,
            "single_thread_note": "1 process on 1 core, the reference's own execution model (single-threaded per MPI rank); 'value' runs the same loop on `cores` threads over disjoint element ranges"}


def time_gpu(fn, stream, reps, warm=2):
    import torch
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def vec_flops_per_element(n, bs, c_pw, pressure=False):
    """36 bs n^4 (GRAD + GRAD_T per component) + 30 bs n^3 (the two J^-1 mappings) + 45 n^3 (geometry) +
    18(2n^3+4n^2+8n) (coordinate derivatives) + c_pw n^3 (pointwise) [+ 12 n^4: INTERP + INTERP_T of the pressure field]."""
    return 36 * bs * n ** 4 + 30 * bs * n ** 3 + 45 * n ** 3 + 18 * (2 * n ** 3 + 4 * n ** 2 + 8 * n) + c_pw * n ** 3 + (12 * n ** 4 if pressure else 0)


def assemble_q1_timed(op, lib, x, mesh, stream, bs, reps):
    import torch
    rowptr, colidx = op.q1_pattern(mesh.l2g)
    vals = torch.empty(len(colidx) * bs * bs, dtype=torch.float64, device="cuda")
    l2g = torch.from_numpy(np.ascontiguousarray(mesh.l2g)).cuda()
    ms = time_gpu(lambda: lib.chk(lib.load().dB200OpAssembleQ1(op.h, (lib.C.c_void_p * 1)(x.data_ptr()), lib.ptr(l2g), None, None, lib.ptr(vals))), stream, reps, warm=2)
    del vals
    return ms, int(len(colidx))


def extras(ctx, stream, fp64_peak, hbm_peak):
    """The other configurations BASELINE.json lists, device-resident, single GPU.  Reported beside the headline."""
    import torch
    from dohp_b200 import lib
    out = {}
    # --- ellip p-Laplacian (ellip_p = 1.6) with the 40 B/point stash, 64^3, p = 4 and 8 (configs[1] variant)
    for p in (4, 8):
        mesh = lib.BoxMesh((64, 64, 64), p, node_xyz=False, ghost_l2g=False)
        rule = lib.Rule.preferred(ctx, p); efs = lib.EFS(ctx, rule, p); space = lib.Space.from_mesh(ctx, mesh)
        op = lib.Op(ctx, lib.OP_ELLIP, lib.EllipParam(1.6, 1.0, 0.0), rule, [space], [efs], mesh.coords)
        x = torch.sin(0.37 * torch.arange(mesh.n, dtype=torch.float64, device="cuda")); y = torch.empty_like(x)
        op.function(x, y)
        ms = time_gpu(lambda: op.mult(x, y), stream, 10)
        n = p + 1
        fl = flops_per_element(n) * mesh.nelem
        by = (bytes_per_element(p) + 40 * n ** 3) * mesh.nelem
        out["ellip_plaplace_p%d" % p] = {"gdofs": round(mesh.n / ms / 1e6, 3), "ms": round(ms, 4), "fp64_frac": round(fl / ms / 1e9 / fp64_peak, 4), "hbm_frac": round(by / ms / 1e6 / hbm_peak, 4)}
        if p == 4:   # embedded-Q1 assembly, ellip 64^3 p=4 (>= 0.8 ms at the HBM roofline)
            ms, nnz = assemble_q1_timed(op, lib, x, mesh, stream, 1, 10)
            # the kernel's own FP64 work: about 2 700 FP64 instructions per Q1 patch (8 points x [36 symmetric entries x 3 FMA + 72 for T d_j + state,
            # geometry, exp/log]) at one FMA per lane and clock = fp64_peak / 2
            fl_ms = 2700.0 * mesh.nelem * p ** 3 / (fp64_peak * 1e12 / 2) * 1e3
            out["q1_assembly_ellip_64^3_p4"] = {"ms": round(ms, 3), "reps": 10, "nnz": nnz, "hbm_gbs": round(12 * nnz / ms / 1e6, 1), "hbm_frac": round(12 * nnz / ms / 1e6 / hbm_peak, 4),
                                                "fp64_floor_ms": round(fl_ms, 2), "frac_of_fp64_floor": round(fl_ms / ms, 3), "bound": "fp64 (HBM floor %.2f ms)" % (12 * nnz / hbm_peak / 1e6)}
        op.close()
        # the mass operator of the projection test (src/fs/tests/ex1.c) on the same space: INTERP + INTERP_T only (12 n^4 + geometry), HBM-side of the ridge
        op = lib.Op(ctx, lib.OP_PROJECTION, None, rule, [space], [efs], mesh.coords)
        ms = time_gpu(lambda: op.mult(x, y), stream, 10)
        fl = (12 * n ** 4 + 45 * n ** 3 + 18 * (2 * n ** 3 + 4 * n ** 2 + 8 * n) + 2 * n ** 3) * mesh.nelem
        by = bytes_per_element(p) * mesh.nelem
        out["projection_mass_p%d" % p] = {"gdofs": round(mesh.n / ms / 1e6, 3), "ms": round(ms, 4), "fp64_frac": round(fl / ms / 1e9 / fp64_peak, 4), "hbm_frac": round(by / ms / 1e6 / hbm_peak, 4)}
        op.close(); del mesh, x, y
    # --- elast 48^3 p=6 (configs[2]): matrix-free MatMult + Q1 assembly (BAIJ-3 block CSR)
    p = 6
    mesh = lib.BoxMesh((48, 48, 48), p, node_xyz=False, ghost_l2g=False)
    rule = lib.Rule.preferred(ctx, p); efs = lib.EFS(ctx, rule, p); space = lib.Space.from_mesh(ctx, mesh, bs=3)
    E0, nu = 1.0, 0.3
    op = lib.Op(ctx, lib.OP_ELAST, lib.ElastParam(E0 * nu / ((1 + nu) * (1 - 2 * nu)), E0 / (2 * (1 + nu))), rule, [space], [efs], mesh.coords)
    x = 0.01 * torch.sin(0.37 * torch.arange(mesh.n * 3, dtype=torch.float64, device="cuda")); y = torch.empty_like(x)
    op.function(x, y)
    ms = time_gpu(lambda: op.mult(x, y), stream, 5)
    n = p + 1
    fl = vec_flops_per_element(n, 3, 150) * mesh.nelem
    by = (16 * 3 * p ** 3 + 192 + 4 * n ** 3 + 72 * n ** 3) * mesh.nelem
    out["elast_48^3_p6_matmult"] = {"gdofs": round(mesh.n * 3 / ms / 1e6, 3), "ms": round(ms, 4), "dofs": mesh.n * 3, "fp64_frac": round(fl / ms / 1e9 / fp64_peak, 4), "hbm_frac": round(by / ms / 1e6 / hbm_peak, 4)}
    try:
        ms, nnzb = assemble_q1_timed(op, lib, x, mesh, stream, 3, 5)
        by = nnzb * (72 + 4)
        fl_ms = 14300.0 * mesh.nelem * p ** 3 / (fp64_peak * 1e12 / 2) * 1e3      # 3 symmetric + 6 half units per patch, 8 points each
        out["q1_assembly_elast_48^3_p6_baij3"] = {"ms": round(ms, 2), "reps": 5, "nnz_blocks": nnzb, "hbm_gbs": round(by / ms / 1e6, 1), "hbm_frac": round(by / ms / 1e6 / hbm_peak, 4),
                                                  "fp64_floor_ms": round(fl_ms, 2), "frac_of_fp64_floor": round(fl_ms / ms, 3), "bound": "fp64 (HBM floor %.2f ms)" % (by / hbm_peak / 1e6)}
    except Exception as ex:          # noqa: BLE001 -- report, do not hide
        out["q1_assembly_elast_48^3_p6_baij3"] = {"error": str(ex)[:200]}
    op.close(); del mesh, x, y
    return out


def extras_partitioned(ctx, stream, rank, world, procs, fp64_peak, hbm_peak, dist, cfg4):
    """Configs of BASELINE.json that are quoted on more than one GPU, run on the job's partition (weak scaling, max over ranks):
    configs[3] stokes velocity p=4 / pressure p=3 (32^3 elements per GPU, coupled Jacobian MatMult = MatMult_Nest_StokesCoupled with
    the two-space ghost exchange) and configs[4] ellip p-Laplacian 128^3 elements per GPU, p=8 (Jacobian MatMult reading the stash)."""
    import torch
    from dohp_b200 import lib
    out = {}

    def timed(fn, reps):
        for _ in range(2):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    ne = 32
    gne = tuple(ne * m for m in procs)
    mu = lib.BoxMesh(gne, 4, procs=procs, rank=rank, node_xyz=False, ghost_l2g=False)
    mp = lib.BoxMesh(gne, 3, procs=procs, rank=rank, dirichlet=0, node_xyz=False, ghost_l2g=False)
    rule = lib.Rule.preferred(ctx, 4); efsu, efsp = lib.EFS(ctx, rule, 4), lib.EFS(ctx, rule, 3)
    su, sp = lib.Space.from_mesh(ctx, mu, bs=3), lib.Space.from_mesh(ctx, mp, bs=1)
    op = lib.Op(ctx, lib.OP_STOKES, lib.StokesParam(1.0, 0.1, 1.2, 0.0), rule, [su, sp], [efsu, efsp], mu.coords)
    xu = 0.01 * torch.sin(0.37 * (3 * mu.rstart + torch.arange(mu.n * 3, dtype=torch.float64, device="cuda")))
    xp = torch.cos(0.11 * (mp.rstart + torch.arange(mp.n, dtype=torch.float64, device="cuda")))
    yu, yp = torch.empty_like(xu), torch.empty_like(xp)
    op.function([xu, xp], [yu, yp])
    ms = timed(lambda: op.mult([xu, xp], [yu, yp]), 20)
    nd = mu.nglobal * 3 + mp.nglobal
    n = 5
    fl = vec_flops_per_element(n, 3, 90, pressure=True) * ne ** 3 * world
    by = (16 * (3 * 4 ** 3 + 3 ** 3) + 192 + 4 * (n ** 3 + (n - 1) ** 3) + 64 * n ** 3) * ne ** 3 * world
    out["stokes_p4p3_32^3_per_gpu_matmult"] = {"gdofs": round(nd / ms / 1e6, 3), "ms": round(ms, 4), "dofs": nd, "n_gpus": world, "fp64_frac": round(fl / ms / 1e9 / (fp64_peak * world), 4),
                                               "hbm_frac": round(by / ms / 1e6 / (hbm_peak * world), 4), "config": "BASELINE.json configs[3] (rheology B=1, eps=0.1, p=1.2; coupled Jacobian, two ghost exchanges each way)"}
    op.close(); del mu, mp, xu, xp, yu, yp, su, sp
    torch.cuda.empty_cache()
    # configs[3] as a SOLVE: Newton (cubic backtracking) + FGMRES with the Schur-complement preconditioner (inner CG on the velocity block,
    # embedded-Q1 block diagonal and pressure mass diagonal assembled by the library), 8^3 elements per GPU; tests/dist_krylov.py
    try:
        import dist_krylov
        import oracle_lib as ol
        ol.build_oracle(ref=False)
        t0 = time.perf_counter()
        r = dist_krylov.run_stokes_gpu(ctx, lib, ol, dist, rank, world, procs, ne_per_gpu=(8, 8, 8), p=4)
        torch.cuda.synchronize()
        out["stokes_p4p3_newton_krylov_8^3_per_gpu"] = {"newton_its": r["newton_its"], "fgmres_its": r["outer"], "inner_cg_its": r["inner_total"], "dofs": r["dofs"], "n_gpus": world,
                                                        "fnorm_first_last": [float("%.6g" % r["fnorms"][0]), float("%.3g" % r["fnorms"][-1])], "velocity_err_rel": float("%.3g" % r["velocity_err_rel"]),
                                                        "seconds_incl_setup": round(time.perf_counter() - t0, 2),
                                                        "what": "BASELINE.json configs[3]: stokes Exact1, B=1 eps=0.1 p=1.2, SNES (bt line search, rtol 1e-7) + FGMRES(30) rtol 1e-2 + upper Schur PC; iteration counts are partition independent"}
    except Exception as ex:      # noqa: BLE001
        out["stokes_p4p3_newton_krylov_8^3_per_gpu"] = {"error": str(ex)[:300]}
    if world > 1:
        # embedded-Q1 preconditioner on the partition: every rank assembles its elements (owned rows + one row per ghost node), then the
        # owner-side reduction over NCCL (MatAssemblyBegin/End, ellip.c:528-529) leaves the owned rows of the global matrix on each rank
        try:
            ne, p = 32, 4
            gne = tuple(ne * m for m in procs)
            mesh = lib.BoxMesh(gne, p, procs=procs, rank=rank, node_xyz=False)
            rule = lib.Rule.preferred(ctx, p); efs = lib.EFS(ctx, rule, p); space = lib.Space.from_mesh(ctx, mesh)
            op = lib.Op(ctx, lib.OP_ELLIP, lib.EllipParam(1.6, 1.0, 0.0), rule, [space], [efs], mesh.coords)
            x = torch.sin(0.37 * (mesh.rstart + torch.arange(mesh.n, dtype=torch.float64, device="cuda")))
            rowptr, colidx = op.q1_pattern(mesh.l2g)
            orowptr, ocolidx = op.q1_owner_pattern(rowptr, colidx)
            vals = torch.empty(len(colidx), dtype=torch.float64, device="cuda")
            ovals = torch.empty(len(ocolidx), dtype=torch.float64, device="cuda")
            l2g = torch.from_numpy(np.ascontiguousarray(mesh.l2g)).cuda()
            ms_a = timed(lambda: op.assemble_q1(x, l2g, None, None, vals), 5)
            ms_r = timed(lambda: op.q1_owner_reduce(vals, ovals), 5)
            nnz = torch.tensor([float(len(ocolidx))], dtype=torch.float64, device="cuda")
            dist.all_reduce(nnz)
            ngh = torch.tensor([float(mesh.ngh)], dtype=torch.float64, device="cuda")
            dist.all_reduce(ngh, op=dist.ReduceOp.MAX)
            out["q1_assembly_ellip_32^3_per_gpu_p4_partitioned"] = {"assemble_ms": round(ms_a, 3), "owner_reduce_ms": round(ms_r, 3), "global_nnz": int(nnz.item()), "n_gpus": world,
                                                                    "max_ghost_rows_per_rank": int(ngh.item()), "checksum": float("%.10g" % ovals.sum().item())}
            op.close(); del mesh, x, vals, ovals, l2g, space
            torch.cuda.empty_cache()
        except Exception as ex:      # noqa: BLE001
            out["q1_assembly_ellip_32^3_per_gpu_p4_partitioned"] = {"error": str(ex)[:300]}
    if cfg4:
        try:
            ne, p = 128, 8
            gne = tuple(ne * m for m in procs)
            mesh = lib.BoxMesh(gne, p, procs=procs, rank=rank, node_xyz=False, ghost_l2g=False)
            rule = lib.Rule.preferred(ctx, p); efs = lib.EFS(ctx, rule, p); space = lib.Space.from_mesh(ctx, mesh)
            op = lib.Op(ctx, lib.OP_ELLIP, lib.EllipParam(1.6, 1.0, 0.0), rule, [space], [efs], mesh.coords)
            x = torch.sin(0.37 * (mesh.rstart + torch.arange(mesh.n, dtype=torch.float64, device="cuda"))); y = torch.empty_like(x)
            op.function(x, y)
            ms = timed(lambda: op.mult(x, y), 5)
            n = p + 1
            fl = flops_per_element(n) * ne ** 3 * world
            by = (bytes_per_element(p) + 40 * n ** 3) * ne ** 3 * world
            out["ellip_plaplace_128^3_per_gpu_p8_matmult"] = {"gdofs": round(mesh.nglobal / ms / 1e6, 3), "ms": round(ms, 3), "dofs": mesh.nglobal, "n_gpus": world,
                                                              "fp64_frac": round(fl / ms / 1e9 / (fp64_peak * world), 4), "hbm_frac": round(by / ms / 1e6 / (hbm_peak * world), 4),
                                                              "config": "BASELINE.json configs[4] (weak scaling, 1.07e9 dofs and a 61 GB stash per GPU)"}
            op.close(); del mesh, x, y, space
            torch.cuda.empty_cache()
        except Exception as ex:      # noqa: BLE001 -- report, do not hide
            out["ellip_plaplace_128^3_per_gpu_p8_matmult"] = {"error": str(ex)[:300]}
    return out


def bind_to_gpu_numa_node(local_rank):
    """Pin this rank's threads to the CPUs of the NUMA node its GPU hangs off, before any pinned host buffer is allocated (first-touch
    places the pages there), so that N ranks do not funnel their PCIe traffic through one socket's memory.  Best effort."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, dev)
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = []
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.extend(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def reference_arm(args, rank):
    if rank != 0:
        return
    degrees = [int(d) for d in args.degrees.split(",")]
    metric, unit = METRIC, UNIT
    workload = workload_name(args)
    cores = os.cpu_count() or 1
    r = cpu_reference_run(degrees, cores, max(args.cpu_budget, 5.0) * 1.0)
    cb = cpu_baseline_json(r, cores, unit)
    v = cb["value"]
    print(json.dumps({"impl": "reference", "metric": metric, "value": v, "unit": unit, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": r["seconds"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                      "config": {"workload": workload, "sample": r["sample"]},
                      "cpu_baseline": cb,
                      "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))


METRIC, UNIT = "matrix-free MatMult GDoF/s (FP64, p=2-10)", "GDoF/s"


def workload_name(args):
    return "ellip %s, %d^3 hex elements per GPU, p=%s sweep, one MatMult per degree per step" % (
        "linear Poisson" if args.ellip_p == 2.0 else "p-Laplacian (ellip_p=%g, Jacobian MatMult reading the stash)" % args.ellip_p, args.ne, args.degrees)


def headline(env, morph, steps, with_e2e, with_cpu):
    """The sweep of BASELINE.json configs[1] on the job's partition: returns the JSON line's dict (meaningful on rank 0)."""
    args, rank, world, local_rank = env["args"], env["rank"], env["world"], env["local_rank"]
    torch, dist, lib, stream, ctx, procs, peaks, numa_node = env["torch"], env["dist"], env["lib"], env["stream"], env["ctx"], env["procs"], env["peaks"], env["numa_node"]
    degrees = [int(d) for d in args.degrees.split(",")]
    metric, unit, workload = METRIC, UNIT, workload_name(args)
    gne = tuple(args.ne * m for m in procs)

    # ---- setup: one space + operator per degree ---------------------------------------------------------------------------
    ops = []
    for p in degrees:
        mesh = lib.BoxMesh(gne, p, procs=procs, rank=rank, node_xyz=False, ghost_l2g=False, morph=morph, twist=0.4 * morph)
        rule = lib.Rule.preferred(ctx, p)
        efs = lib.EFS(ctx, rule, p)
        space = lib.Space.from_mesh(ctx, mesh)
        op = lib.Op(ctx, lib.OP_ELLIP, lib.EllipParam(args.ellip_p, 1.0, 0.0), rule, [space], [efs], mesh.coords)
        x = torch.sin(0.37 * (mesh.rstart + torch.arange(mesh.n, dtype=torch.float64, device="cuda")))      # x_i = sin(0.37 i)
        y = torch.empty_like(x)
        if args.ellip_p != 2.0:
            op.function(x, y)            # EllipFunction at the linearisation state: writes the stash the MatMults read (ellip.c:360)
        ops.append(dict(p=p, op=op, x=x, y=y, n=mesh.n, nelem=mesh.nelem, nglobal=mesh.nglobal, keep=(rule, efs, space)))
        del mesh
    if "fp64_peak" not in env:
        env["fp64_peak"] = ctx.fp64_peak_tflops()
    fp64_peak = env["fp64_peak"]
    hbm_peak, hbm_src = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")
    env["hbm_peak"] = hbm_peak

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device(events=None):
        for i, o in enumerate(ops):
            if events is not None:
                events[i][0].record(stream)
            o["op"].mult(o["x"], o["y"])
            if events is not None:
                events[i][1].record(stream)

    # ---- device-resident timing ----------------------------------------------------------------------------------------------
    # nvidia-smi needs a few 100 ms to deliver its first row: start it now and keep every GPU busy for ~0.6 s before the warm-up.
    # Every rank runs the SAME number of steps (the MatMults contain the NCCL ghost exchange), the count comes from rank 0.
    sampler = ClockSampler(local_rank)
    if rank == 0 and with_e2e:
        sampler.start()
    step_device()
    torch.cuda.synchronize()
    t_probe = time.perf_counter()
    step_device()
    torch.cuda.synchronize()
    n_pre = torch.tensor([min(200, max(1, int(0.6 / max(time.perf_counter() - t_probe, 1e-4)) + 1)) if with_e2e else 1], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.broadcast(n_pre, 0)
    for _ in range(int(n_pre.item())):
        step_device()
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in ops] for _ in range(steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = ctx.num_launches()
    barrier()
    e0.record(stream)
    for k in range(steps):
        step_device(ev[k])
    e1.record(stream)
    barrier()
    launches = ctx.num_launches() - n0
    total_ms = e0.elapsed_time(e1)
    per_p_ms = [sum(ev[k][i][0].elapsed_time(ev[k][i][1]) for k in range(steps)) / steps for i in range(len(ops))]
    if rank == 0 and len(sampler.rows) > 1:
        sampler.rows = sampler.rows[1:]               # the first row may predate the load
    clocks = sampler.finish() if rank == 0 and with_e2e else None
    t = torch.tensor([total_ms] + per_p_ms, dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, per_p_ms = t[0].item(), t[1:].tolist()
    ms_per_step = total_ms / steps
    dofs_step = sum(o["nglobal"] for o in ops)
    value = dofs_step / (ms_per_step * 1e-3) / 1e9

    e2e_value = e2e_blocking_ms = chk = None
    h2d = 0
    if with_e2e:
        # ---- end to end: host vectors through the C ABI (H2D + kernel + D2H inside the timed region) -------------------------------
        hx = [torch.sin(0.37 * (torch.arange(o["n"], dtype=torch.float64))).pin_memory() for o in ops]
        hy = [torch.empty(o["n"], dtype=torch.float64).pin_memory() for o in ops]
        e2e_steps = max(1, steps)                        # the K timed steps are enqueued back to back: the copy streams stay busy across step boundaries
        blocking_steps = max(1, min(steps, 3))
        # the user-facing call for host vectors is dB200OpMultAsync: every product uploads x, runs the traversal and downloads y; the
        # products of the sweep are independent, so their uploads, kernels and downloads overlap (both PCIe directions busy).  One
        # step = all degrees enqueued, then a device-side join of the copy streams; every byte of x and y crosses PCIe inside the timed region.
        def step_e2e():
            for i, o in enumerate(ops):
                o["op"].mult_async(hx[i], hy[i])
            ctx.stream_join()

        step_e2e()                                       # warm-up (allocates the per-operator staging buffers)
        ctx.synchronize()
        barrier()
        e0.record(stream)
        for _ in range(e2e_steps):
            step_e2e()
        e1.record(stream)
        ctx.synchronize()
        barrier()
        e2e_ms = torch.tensor([e0.elapsed_time(e1) / e2e_steps], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
        # the same with the blocking call (one product at a time: upload, kernel, download back to back)
        for i, o in enumerate(ops):
            o["op"].mult(hx[i], hy[i])                 # warm-up (grows the context's staging buffers)
        barrier()
        e0.record(stream)
        for _ in range(blocking_steps):
            for i, o in enumerate(ops):
                o["op"].mult(hx[i], hy[i])
        e1.record(stream)
        barrier()
        e2e_blocking_ms = torch.tensor([e0.elapsed_time(e1) / blocking_steps], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(e2e_blocking_ms, op=dist.ReduceOp.MAX)
        e2e_value = dofs_step / (e2e_ms.item() * 1e-3) / 1e9
        h2d = sum(o["n"] for o in ops) * 8
        chk = float(sum(h.double().abs().sum().item() for h in hy))

    # ---- roofline of the element kernel (per degree and aggregated over the sweep) ---------------------------------------------
    per_p, flops_step, bytes_step = [], 0.0, 0.0
    for o, ms in zip(ops, per_p_ms):
        n = o["p"] + 1
        fl, by = flops_per_element(n) * o["nelem"], (bytes_per_element(o["p"]) + (0 if args.ellip_p == 2.0 else 40 * n ** 3)) * o["nelem"]
        flops_step += fl
        bytes_step += by
        tf, gbs = fl / (ms * 1e-3) / 1e12, by / (ms * 1e-3) / 1e9
        per_p.append({"p": o["p"], "dofs": o["nglobal"], "ms": round(ms, 4), "gdofs": round(o["nglobal"] / (ms * 1e-3) / 1e9, 3),
                      "tflops": round(tf, 3), "fp64_frac": round(tf / fp64_peak, 4), "hbm_gbs": round(gbs, 1), "hbm_frac": round(gbs / hbm_peak, 4),
                      "bound": "fp64" if flops_per_element(n) / bytes_per_element(o["p"]) > fp64_peak * 1e3 / hbm_peak else "hbm"})
    kern_s = sum(per_p_ms) * 1e-3
    ach = flops_step / kern_s / 1e12
    # DRAM bytes (and the executed FP64-pipe utilisation) of the same launches from the newest committed ncu --set full capture
    traffic, traffic_src = None, None
    try:
        import glob
        capf = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")))[-1]
        cap = json.load(open(capf))["per_degree"]
        if args.ne == 64 and world == 1 and morph == 0 and args.ellip_p == 2.0 and all(str(o["p"]) in cap for o in ops):
            traffic = round(sum(cap[str(o["p"])]["dram_read_gb"] + cap[str(o["p"])]["dram_write_gb"] for o in ops) * 1e9)
            traffic_src = "dram__bytes_read.sum + dram__bytes_write.sum summed over the %d launches of one step, ncu capture profiles/%s (algorithmic bytes of the same launches: %d)" % (len(ops), os.path.basename(capf), bytes_step)
            for e in per_p:
                e["fp64_pipe_pct_ncu"] = cap[str(e["p"])]["fp64_pipe_pct"]
    except Exception:
        pass
    roofline = {"bound": "fp64", "achieved": round(ach, 3), "peak": round(fp64_peak, 3), "unit": "TFLOP/s", "frac": round(ach / fp64_peak, 4), "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": "measured live: register-resident DFMA loop (dB200MeasureFP64Peak); MEASURED_PEAKS.json has no FP64 entry",
                "kernel": "ellip_kernel<N,N,%s> (one launch per degree; algorithmic flops 36n^4+131n^3+72n^2+144n per element)" % ("LINEAR" if args.ellip_p == 2.0 else "JACOBIAN"),
                "hbm": {"achieved": round(bytes_step / kern_s / 1e9, 1), "peak": hbm_peak, "unit": "GB/s", "frac": round(bytes_step / kern_s / 1e9 / hbm_peak, 4), "peak_source": hbm_src},
                "per_p": per_p,
                "frac_note": "algorithmic flops (the reference's count, SURVEY.md 8d) over the live-measured DFMA peak; the kernel EXECUTES fewer: even-odd 1-D contractions, collocation derivatives, and on affine elements one symmetric 3x3 product per point, so frac can exceed 1; per_p[].fp64_pipe_pct_ncu is the executed FP64-pipe utilisation (sm__pipe_fp64_cycles_active) of the same launches under ncu"}

    cores = os.cpu_count() or 1
    if args.cpu_budget > 0 and with_cpu and rank == 0:
        cb = cpu_baseline_json(cpu_reference_run(degrees, cores, args.cpu_budget), cores, unit)
    else:
        cb = {"value": None, "unit": unit, "cores": cores, "kind": "skipped", "sample": "skipped (--cpu-budget 0, profiling run)"}
    out = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": workload, "elements_per_gpu": args.ne ** 3, "proc_grid": list(procs), "geometry": "affine box" if morph == 0 else "morphed (non-affine)",
                      "ghost_exchange": sorted({o["keep"][2].halo_mode() for o in ops}),
                      "l2": "a step cycles through %d problem sizes totalling %.1f GB of vectors+indices per GPU (>> 126 MB L2) before any is revisited"
                            % (len(ops), sum(16 * o["n"] + 4 * o["nelem"] * (o["p"] + 1) ** 3 for o in ops) / 1e9)},
           "clocks": clocks, "gpu_launches": launches,
           "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": h2d, "checksum": chk,
                   "call": "dB200OpMultAsync per degree + dB200StreamJoin per step on every rank (pinned host vectors, rank threads bound to the GPU's NUMA node %s; uploads, kernels incl. the NCCL ghost exchange and downloads of the independent degrees overlap)" % str(numa_node),
                   "blocking_call_value": dofs_step / (e2e_blocking_ms.item() * 1e-3) / 1e9} if with_e2e else None,
           "roofline": roofline,
           "cpu_baseline": cb}
    out["roofline"]["frac_of_nominal_37.2TF"] = round(ach / 37.2, 4)
    for o in ops:
        o["op"].close()
    ops.clear()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--ne", type=int, default=64)
    ap.add_argument("--degrees", default="2,3,4,5,6,7,8,9,10")
    ap.add_argument("--morph", type=float, default=0.0, help="non-affine variant: Morph(morph, twist=0.4*morph), ellip.c:173-181")
    ap.add_argument("--ellip-p", type=float, default=2.0, help="p of the p-Laplacian (ellip.c:156); != 2 runs the Jacobian MatMult that reads the 40 B/point stash (BASELINE.json configs[4])")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--extras", type=int, default=1, help="also time the other configs of BASELINE.json (elast 48^3 p=6 MatMult + Q1, stokes p4/p3, ellip Q1 assembly, p-Laplace with stash, morphed-mesh sweep)")
    ap.add_argument("--cfg4", type=int, default=1, help="include BASELINE.json configs[4] (128^3 elements per GPU, p=8, 61 GB stash per GPU) in other_configs")
    ap.add_argument("--parity", type=int, default=1, help="untimed parity block: partitioned ellip/elast/stokes MatMult and Q1 rows vs the single-domain oracle")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        reference_arm(args, rank)
        return
    # rank 0 must print exactly one JSON line: NCCL writes its version banner to stdout at NCCL_DEBUG=WARN/VERSION
    if os.environ.get("NCCL_DEBUG", "").upper() in ("WARN", "VERSION"):
        os.environ.pop("NCCL_DEBUG")
    import torch
    import torch.distributed as dist
    from dohp_b200 import lib
    torch.cuda.set_device(local_rank)
    numa_node = bind_to_gpu_numa_node(local_rank)
    nccl_id = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.frombuffer(bytearray(lib.Context.nccl_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(idt, 0)
        nccl_id = bytes(idt.cpu().numpy().tobytes())
    stream = torch.cuda.current_stream()
    ctx = lib.Context(local_rank, rank, world, nccl_id, stream=stream.cuda_stream)
    procs = proc_grid(world)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    env = dict(args=args, rank=rank, world=world, local_rank=local_rank, torch=torch, dist=dist, lib=lib, stream=stream, ctx=ctx, procs=procs, peaks=peaks, numa_node=numa_node)
    out = headline(env, args.morph, args.steps, with_e2e=True, with_cpu=True)
    fp64_peak, hbm_peak = env["fp64_peak"], env["hbm_peak"]
    torch.cuda.empty_cache()
    other = {}
    if args.extras:
        try:
            other.update(extras_partitioned(ctx, stream, rank, world, procs, fp64_peak, hbm_peak, dist, args.cfg4))
        except Exception as ex:      # noqa: BLE001 -- the headline line must still be printed
            other["partitioned_error"] = str(ex)[:300]
        if world == 1:
            try:
                other.update(extras(ctx, stream, fp64_peak, hbm_peak))
                if args.morph == 0:          # the general (non-affine) geometry path, driver-timed: the same sweep on Morph(0.3, twist 0.12)
                    m = headline(env, 0.3, max(2, min(args.steps, 5)), with_e2e=False, with_cpu=False)
                    other["headline_sweep_morphed_mesh"] = {"gdofs": round(m["value"], 3), "ms_per_step": round(m["ms_per_step"], 3), "fp64_frac": m["roofline"]["frac"],
                                                            "per_p_gdofs": [q["gdofs"] for q in m["roofline"]["per_p"]], "geometry": "Morph(morph=0.3, twist=0.12), ellip.c:173-181: every element non-affine"}
            except Exception as ex:  # noqa: BLE001
                other["single_gpu_error"] = str(ex)[:300]
    parity = None
    if args.parity:
        import partition_parity
        try:
            parity = partition_parity.run_all(ctx, dist, rank, world, procs, quick=True)
            if rank == 0:
                parity = {k: float("%.3g" % v) for k, v in parity.items()}
                parity["ok"] = all(v <= 1e-12 for v in parity.values())
                parity["what"] = ("max relative error vs the SINGLE-DOMAIN oracle of a small global problem block-partitioned over the %d ranks: ellip / elast / stokes "
                                  "(two spaces, A, B, Bt and coupled J) MatMult with the ghost exchange inside (NVLink peer memory, NCCL fallback), and embedded-Q1 rows summed by global (row, col); bar 1e-12" % world)
        except Exception as ex:      # noqa: BLE001
            parity = {"ok": False, "error": str(ex)[:300]}
    if rank == 0:
        if other:
            out["other_configs"] = other
        out["parity"] = parity
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    if rank == 0 and parity is not None and not parity.get("ok", False):
        sys.exit(3)


if __name__ == "__main__":
    main()
