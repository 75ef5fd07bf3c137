"""The drop-in boundary exercised FROM THE REFERENCE'S SIDE: integration/jacobi_b200.c registers a dJacobi type "b200" written against
the reference's own headers (include/dohpjacimpl.h, tensorimpl.h) whose dEFSApply forwards to dB200EFSApplyBatch.  oracle/Makefile
(`make refbind`, run by build() in the build container) links it with the reference's unmodified library sources and

* the reference's UNMODIFIED test program src/jacobi/tests/ex1.c: run with -djac_type b200 it must print the golden
  jacobi-1.refout / jacobi-1-lobatto.refout (src/jacobi/tests/CMakeLists.txt:5-6);
* integration/seam_check.c: one process, a "tensor" dJacobi (reference CPU code) and a "b200" one, every apply mode x INSERT/ADD x
  jinv / NULL on hexes (isotropic, anisotropic, P > Q, Q = 12, SPARSE rule), a quad and a line, compared to 1e-12; error cases error alike.

The CPU tests check what can be checked without a device: the type registers and is selected, refuses loudly without a GPU (no CPU
fallback behind the seam), and leaves the tensor type's golden output untouched."""
import os
import re
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REFBIN = os.path.join(os.path.dirname(HERE), "oracle", "_ref")
EX1, SEAM = os.path.join(REFBIN, "jacobi_ex1_b200"), os.path.join(REFBIN, "seam_check_b200")
GOLDEN = os.path.join(HERE, "golden", "refout")
ARGS = {"jacobi-1": "-min_rdeg 10 -max_rdeg 11 -max_bdeg 10", "jacobi-1-lobatto": "-min_rdeg 9 -max_rdeg 11 -max_bdeg 10 -dquad_tensor_gauss_family lobatto"}

needs_binding = pytest.mark.skipif(not (os.path.exists(EX1) and os.path.exists(SEAM)), reason="oracle/_ref/*_b200 not built (needs /root/reference at build time: make -C oracle refbind)")


def have_gpu():
    import torch
    return torch.cuda.is_available()


def run(exe, args, **env):
    return subprocess.run([exe] + args.split(), capture_output=True, text=True, timeout=300, env=dict(os.environ, **env))


LINE = re.compile(r"^(L2 error for element +\d+ \( *\d+, *\d+, *\d+\) with rule \( *\d+, *\d+, *\d+\)) interp +(\S+), derivative +(\S+)$")


def same_as_golden(mine, gold):
    """Text equality, except that an error printed as d.de-xx may differ by one unit of its last digit when it is below 1e-8 (there the
    L2 error of a degree >= 9 interpolant is partly accumulated rounding, which the GPU and the CPU sum in different orders), and that an
    error below 1e-14 is pure rounding noise (jacobi-1-lobatto prints 5.2e-31 where the rule points ARE the nodes): noise on both sides."""
    ml, gl = mine.splitlines(), gold.splitlines()
    assert len(ml) == len(gl), (len(ml), len(gl))
    for a, b in zip(ml, gl):
        if a == b:
            continue
        ma, mb = LINE.match(a), LINE.match(b)
        assert ma and mb and ma.group(1) == mb.group(1), (a, b)
        for x, y in zip(ma.groups()[1:], mb.groups()[1:]):
            if x != y:
                fy = float(y)
                if fy < 1e-14:
                    assert float(x) < 1e-14, (a, b)
                    continue
                assert fy < 1e-8 and abs(float(x) - fy) <= 0.11 * 10.0 ** int(y.split("e")[1]), (a, b)


@needs_binding
@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(ARGS))
def test_reference_ex1_runs_unmodified_on_the_b200_type(name):
    r = run(EX1, ARGS[name] + " -djac_type b200")
    assert r.returncode == 0, r.stderr[-2000:]
    same_as_golden(r.stdout, open(os.path.join(GOLDEN, name + ".refout")).read())


@needs_binding
@pytest.mark.gpu
def test_tensor_and_b200_types_agree_through_the_reference_api():
    r = run(SEAM, "")
    print(r.stdout)
    assert r.returncode == 0 and "0 failures: PASS" in r.stdout, (r.stdout[-3000:], r.stderr[-2000:])
    assert r.stderr.count("error") >= 4                      # the two error cases were refused by both implementations


@needs_binding
def test_b200_type_registers_and_has_no_cpu_fallback():
    if have_gpu():
        pytest.skip("a GPU is present: the GPU tests of this file cover the type")
    r = run(EX1, ARGS["jacobi-1"] + " -djac_type b200")
    assert "no CUDA device" in r.stderr and "dJacobiGetEFS_B200" in r.stderr           # selected by name, failed loudly inside the binding
    assert "L2 error" not in r.stdout                                                    # and nothing was computed on the CPU instead


@needs_binding
def test_binding_leaves_the_tensor_type_untouched():
    """The same executable with the default type prints the golden file byte for byte; the seam driver agrees with itself."""
    r = run(EX1, ARGS["jacobi-1"])
    assert r.returncode == 0 and r.stdout == open(os.path.join(GOLDEN, "jacobi-1.refout")).read()
    r = run(SEAM, "", DOHP_SEAM_OTHER="tensor")
    assert r.returncode == 0 and "0 failures: PASS" in r.stdout
