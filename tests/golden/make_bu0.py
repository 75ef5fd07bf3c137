"""Extracts the numbering-pinning vectors of the reference's golden src/fs/tests/refout/bu-0.refout
(bu -dmesh_gen block -dmeshgen_block_mnp 3,2,2 -dfs_ordering_type natural; src/fs/tests/CMakeLists.txt:47) into
tests/golden/bu0.json.  Run in the build container, where /root/reference is mounted:  python tests/golden/make_bu0.py"""
import json
import os

REF = "/root/reference/src/fs/tests/refout/bu-0.refout"
lines = open(REF).read().splitlines()


def block(title, count):
    i = next(k for k, l in enumerate(lines) if l.startswith(title))
    vals, k = [], i + 1
    while len(vals) < count:
        try:
            vals.append(float(lines[k]))
        except ValueError:
            pass
        k += 1
    return vals


out = {
    "source": "src/fs/tests/refout/bu-0.refout",
    "ne": [2, 1, 1], "p": 2, "dirichlet_sets": [100],
    "n": 9, "nc": 45, "expanded": 54,
    "expanded_homogeneous": block("Expanded vector projected into homogeneous space", 54),
    "expanded_inhomogeneous": block("Expanded vector projected into inhomogeneous space", 54),
    "closure_coordinates": block("Global coordinate vector", 135),
}
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "bu0.json"), "w") as f:
    json.dump(out, f)
print("wrote bu0.json")
