"""PCIe probe: host<->device bandwidth of every visible GPU alone and all together (pinned buffers, both directions at once), and the
topology nvidia-smi reports.  Explains the end-to-end (host-vector) scaling of bench.py.  python tools/pcie_probe.py"""
import subprocess
import threading
import time

import torch

print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout)
n = torch.cuda.device_count()
N = 1 << 28          # 256 MiB
bufs = []
for d in range(n):
    with torch.cuda.device(d):
        bufs.append((torch.empty(N, dtype=torch.uint8).pin_memory(), torch.empty(N, dtype=torch.uint8).pin_memory(),
                     torch.empty(N, dtype=torch.uint8, device="cuda"), torch.empty(N, dtype=torch.uint8, device="cuda"),
                     torch.cuda.Stream(), torch.cuda.Stream()))


def run(d, reps, out):
    hx, hy, dx, dy, s1, s2 = bufs[d]
    with torch.cuda.device(d):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            with torch.cuda.stream(s1):
                dx.copy_(hx, non_blocking=True)
            with torch.cuda.stream(s2):
                hy.copy_(dy, non_blocking=True)
        torch.cuda.synchronize()
        out[d] = reps * N / (time.perf_counter() - t0) / 1e9


for d in range(n):
    o = {}
    run(d, 2, o)
    run(d, 16, o)
    print("GPU %d alone: %.1f GB/s each way (H2D and D2H concurrently)" % (d, o[d]))
o = {}
th = [threading.Thread(target=run, args=(d, 16, o)) for d in range(n)]
[t.start() for t in th]
[t.join() for t in th]
print("all %d GPUs together: %s GB/s each way per GPU, %.1f total" % (n, {d: round(v, 1) for d, v in o.items()}, sum(o.values())))
