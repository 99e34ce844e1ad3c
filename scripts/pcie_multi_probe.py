"""One process, several GPUs: aggregate host<->device bandwidth for different NUMA placements of the pinned host buffers.

Development aid for the one-process host path (cb200_init_devices): where must a caller's pinned buffers live so that
every GPU's copies run at link speed at the same time?  Run under `gpurun --gpus N`; prints one line per placement.
"""
import os
import re
import subprocess
import time

import torch

topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout
print(topo[:4000])
ng = torch.cuda.device_count()
ALL = os.sched_getaffinity(0)


def cpus_of(gpu):
    m = re.search(r"^GPU%d\s.*?(\d+(?:-\d+)?(?:,\d+(?:-\d+)?)*)\s+(\d+)" % gpu, topo, flags=re.M)
    cpus = set()
    if m:
        for part in m.group(1).split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
    return (cpus & ALL) or ALL, (m.group(2) if m else "?")


N = 1 << 29


def alloc_near(gpu):
    """pinned in/out buffers allocated and first touched by this thread while it is bound next to `gpu` (None: unbound)"""
    os.sched_setaffinity(0, cpus_of(gpu)[0] if gpu is not None else ALL)
    time.sleep(0.01)
    a = torch.empty(N, dtype=torch.uint8).pin_memory()
    a.fill_(1)
    b = torch.empty(N, dtype=torch.uint8).pin_memory()
    b.fill_(2)
    os.sched_setaffinity(0, ALL)
    return a, b


def run(label, place):
    host = [alloc_near(place(g)) for g in range(ng)]
    dev = [(torch.empty(N, dtype=torch.uint8, device="cuda:%d" % g), torch.empty(N, dtype=torch.uint8, device="cuda:%d" % g))
           for g in range(ng)]
    st = [(torch.cuda.Stream(device=g), torch.cuda.Stream(device=g)) for g in range(ng)]
    for mode in ("h2d", "d2h", "both"):
        for g in range(ng):
            torch.cuda.synchronize(g)
        t0 = time.perf_counter()
        for _ in range(3):
            for g in range(ng):
                if mode in ("h2d", "both"):
                    with torch.cuda.stream(st[g][0]):
                        dev[g][0].copy_(host[g][0], non_blocking=True)
                if mode in ("d2h", "both"):
                    with torch.cuda.stream(st[g][1]):
                        host[g][1].copy_(dev[g][1], non_blocking=True)
        for g in range(ng):
            st[g][0].synchronize()
            st[g][1].synchronize()
        dt = (time.perf_counter() - t0) / 3
        print("%-28s %-5s %6.1f GB/s per direction per GPU, %7.1f aggregate per direction" % (label, mode, N / dt / 1e9, ng * N / dt / 1e9),
              flush=True)
    del host, dev


for g in range(ng):
    print("GPU%d: %d local cpus, numa %s" % (g, len(cpus_of(g)[0]), cpus_of(g)[1]))
run("unbound (first touch)", lambda g: None)
run("all next to GPU0", lambda g: 0)
run("all next to GPU%d" % (ng - 1), lambda g: ng - 1)
run("each next to its own GPU", lambda g: g)
