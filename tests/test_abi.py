"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol include/dohp_b200.h declares;
host-only entry points (1-D tables) work, compute entry points fail loudly without a device (no CPU path)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle_lib as ol
from dohp_b200 import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def _build():
    lib.build()
    ol.build_oracle()


def test_every_declared_symbol_is_exported():
    hdr = open(os.path.join(ROOT, "include", "dohp_b200.h")).read()
    declared = set(re.findall(r"\b(dB200[A-Za-z0-9]+)\s*\(", hdr))
    assert declared == set(lib.SYMBOLS)
    L = lib.load()
    for name in sorted(declared):
        assert hasattr(L, name), name


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(lib.DohpError, match="no CUDA device|no CPU path|CUDA"):
        lib.Context(0)


def test_product_does_not_link_or_import_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "dohp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                src = open(os.path.join(dirpath, f)).read()
                assert "liboracle" not in src and "oracle_lib" not in src and "oracle/" not in src.replace("the oracle", ""), (dirpath, f)


@pytest.mark.parametrize("method,family,order", [(lib.FAST, lib.GAUSS, o) for o in range(0, 22, 2)] + [(lib.SPARSE, lib.GAUSS, o) for o in range(2, 22, 2)] +
                         [(lib.SELF, lib.LOBATTO, o) for o in (2, 6, 12, 20)] + [(lib.FAST, lib.LOBATTO, 4), (lib.FAST, lib.LOBATTO, 10)])
def test_rule_tables_match_oracle(method, family, order):
    c, w = lib.rule_1d(method, family, order)
    oc, ow = ol.rule_1d(method, family, order)
    assert c.shape == oc.shape
    assert np.allclose(c, oc, rtol=0, atol=4e-16) and np.allclose(w, ow, rtol=0, atol=2e-15)


@pytest.mark.parametrize("P", range(2, 12))
def test_basis_tables_match_oracle(P):
    for qc in (lib.rule_1d(lib.FAST, lib.GAUSS, 2 * (P - 1))[0], lib.rule_1d(lib.FAST, lib.GAUSS, 2 * P)[0], lib.rule_1d(lib.SPARSE, lib.GAUSS, 2 * (P - 1))[0]):
        i, d, n, w = lib.basis_1d(P, qc)
        oi, od, on, ow = ol.basis_1d(P, qc)
        assert np.allclose(n, on, rtol=0, atol=4e-16) and np.allclose(w, ow, rtol=0, atol=2e-15)
        assert np.allclose(i, oi, rtol=0, atol=2e-14) and np.allclose(d, od, rtol=0, atol=2e-14 * np.abs(od).max())


def test_header_is_plain_c99_and_the_abi_is_usable_from_c(tmp_path):
    """integration/ellip_matmult_example.c (the operator seam from plain C99: box mesh, 1-D tables, rule, dEFS, space, operator, one
    product) compiles warning-free with -std=c99 -pedantic -Werror against include/dohp_b200.h, its host-side setup runs (8^3 elements,
    degree 4: 31^3 dofs), and without a GPU it stops at dB200Init with the library's message: no CPU path behind the ABI."""
    import subprocess
    exe = str(tmp_path / "ellip_example")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "integration", "ellip_matmult_example.c"), "-L" + os.path.join(ROOT, "dohp_b200"), "-ldohp_b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "dohp_b200"), "-lm", "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert "host setup: 512 elements, 29791 dofs (+ 6146 Dirichlet nodes), 5^3 quadrature points per element" in r.stdout
    import torch
    if not torch.cuda.is_available():
        assert r.returncode == 1 and "no CUDA device" in r.stderr and "|J x|_2" not in r.stdout
