"""Copies the reference's golden solver outputs that tests/test_refout_replay.py and tests/test_gpu_solver_parity.py replay
(src/fs/tests/refout/*.refout: plain text printed by -snes_monitor_short / -ksp_monitor_short, the known answers of SURVEY.md 8c)
into tests/golden/refout/ so that they travel to the GPU box, where /root/reference does not exist.  Run here, in the build container."""
import os
import shutil

SRC = "/root/reference/src/fs/tests/refout"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refout")
NAMES = ["ellip-e0-b1-p16", "ellip-e0-b3-p16", "ellip-e0-b3-p14-sse", "ellip-e0-b3-p17-lam1-sse", "ellip-morph", "ellip-morph-qfast",
         "elast-e0-b3-g01", "stokes-q3q2-fgmres", "stokes-q3q2-scr",
         # the projection test (src/fs/tests/ex1.c, CMakeLists.txt:37-42): replayed by tests/test_oracle_projection.py / test_gpu_projection.py
         "fs-mf-projection-v1", "fs-mf-projection-v3", "fs-mf-projection-v3-sparse", "fs-mf-op-proj-0", "fs-mf-op-proj-1",
         # the coordinate unit test (src/fs/tests/cunit.c, CMakeLists.txt:44-45): index map and nodal coordinates of the hand-built two-hex mesh
         "cunit-identity-0", "cunit-affine-0",
         # the sub-element (visualisation) mesh test (src/fs/tests/fsex4.c, CMakeLists.txt:49-51): three hexes around an axis
         "fsex4-submesh-p1", "fsex4-submesh-p1-read", "fsex4-submesh-p2-read"]
# the vector closure test (src/vec/tests/vec-ex1.c, 2 ranks): VecDohp windows + forward ghost update, replayed by tests/test_vecd_gloo.py
OTHER = [("/root/reference/src/vec/tests/refout", "vec-ex1"),
         # src/jacobi/tests/ex1.c (CMakeLists.txt:5-6): run UNMODIFIED on the b200 dJacobi type by tests/test_gpu_zz_refbinding.py
         ("/root/reference/src/jacobi/tests/refout", "jacobi-1"), ("/root/reference/src/jacobi/tests/refout", "jacobi-1-lobatto")]
os.makedirs(DST, exist_ok=True)
for n in NAMES:
    shutil.copyfile(os.path.join(SRC, n + ".refout"), os.path.join(DST, n + ".refout"))
for src, n in OTHER:
    shutil.copyfile(os.path.join(src, n + ".refout"), os.path.join(DST, n + ".refout"))
print("copied", len(NAMES) + len(OTHER), "golden outputs")
