"""The driver-facing contract of bench.py that can be checked without a GPU: the reference arm (`--impl reference`: the reference's own
CPU implementation of the path, oracle/_ref where it is built, else the port) prints exactly one JSON line with the keys the driver reads,
loads no product library, and the flops / bytes formulas the roofline uses are SURVEY.md 8(d)'s."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-budget", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "GDoF/s" and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["metric"].startswith("matrix-free MatMult GDoF/s") and "workload" in d["config"] and "model" not in d["config"]
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0 and d["n_gpus"] == 1
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["single_thread_value"] > 0 and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "GDoF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_does_not_touch_the_product_library():
    """the CPU arm must time the reference / the oracle only: importing bench.py's reference leg may not map libdohp_b200.so"""
    code = ("import sys, os; sys.argv=['bench.py','--impl','reference','--steps','1','--warmup','0','--cpu-budget','0.2']; "
            "import runpy; runpy.run_path(%r, run_name='__main__'); "
            "maps=open('/proc/self/maps').read(); assert 'libdohp_b200' not in maps, 'product library mapped in the reference arm'; "
            "assert 'liboracle' in maps or 'libref_driver' in maps" % os.path.join(ROOT, "bench.py"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-500:], r.stderr[-1500:])


def test_roofline_formulas_are_the_surveys():
    sys.path.insert(0, ROOT)
    import bench
    # ellip (bs=1, lambda=0) F = 36n^4+131n^3+72n^2+144n ; p=2: 7.5 kflop, p=4: 41.4, p=8: 339, p=10: 712 (rounded)
    assert [round(bench.flops_per_element(p + 1) / 1e3, 1) for p in (2, 4, 8, 10)] == [7.5, 41.4, 338.8, 711.7]
    # B = 16 p^3 + 192 + 4 n^3: p=2: 428 B, p=4: 1.72 KB, p=8: 11.3 KB
    assert bench.bytes_per_element(2) == 428 and bench.bytes_per_element(4) == 1716 and bench.bytes_per_element(8) == 11300
    assert bench.proc_grid(8) == (2, 2, 2) and bench.proc_grid(2) == (2, 1, 1)
