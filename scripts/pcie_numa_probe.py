"""Does the NUMA placement of pinned host buffers matter for host<->device copies on this box?  (development aid)"""
import os
import re
import subprocess
import torch

print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:3000])
print("affinity now:", len(os.sched_getaffinity(0)), "cpus", sorted(os.sched_getaffinity(0))[:8], "...")


def bw(label):
    n = 1 << 30
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_in.fill_(1)
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out.fill_(2)
    d1 = torch.empty(n, dtype=torch.uint8, device="cuda")
    d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for mode in ("h2d", "d2h", "both"):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d1.copy_(h_in, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_out.copy_(d2, non_blocking=True)
        s1.synchronize(); s2.synchronize()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        print("%s %-5s %.1f GB/s per direction" % (label, mode, n / ms / 1e6))


bw("default ")
topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout
m = re.search(r"^GPU0\s.*?(\d+(?:-\d+)?(?:,\d+(?:-\d+)?)*)\s+(\d+)", topo, flags=re.M)
if m:
    cpus = set()
    for part in m.group(1).split(","):
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    cpus &= os.sched_getaffinity(0) | cpus
    try:
        os.sched_setaffinity(0, cpus)
        print("affinity set to GPU0-local cpus:", len(cpus), "numa", m.group(2))
        bw("gpu-local")
    except OSError as e:
        print("cannot set affinity:", e)
