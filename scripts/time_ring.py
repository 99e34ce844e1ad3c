"""Quick device-side timing of the ring kernels (development aid; bench.py is the contract)."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import circl_b200
from circl_b200 import kyber

circl_b200.init(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
Q = 3329
d = (torch.randint(0, 2 * Q, (n, 256), device="cuda", dtype=torch.int32) - Q).to(torch.int16)
res = {}
for name, fn in (("ntt", lambda: kyber.ntt_(d)), ("invntt", lambda: kyber.inv_ntt_(d)),
                 ("barrett", lambda: kyber.barrett_reduce(d, out=d)),
                 ("mulhat", lambda: kyber.mul_hat(d, d, out=d))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)
    med = ms[len(ms) // 2]
    res[name] = {"ms_med": med, "ms_min": ms[0], "per_s": n / (med * 1e-3), "GBps_1024B": n * 1024 / (med * 1e-3) / 1e9}
from circl_b200 import dilithium as dl
QD = 8380417
e = torch.randint(0, QD, (n // 2, 256), device="cuda", dtype=torch.int32)
for name, fn in (("dil_ntt", lambda: dl.ntt_(e)), ("dil_invntt", lambda: dl.inv_ntt_(e)),
                 ("dil_reduce", lambda: dl.reduce_le2q(e, out=e)), ("dil_mulhat", lambda: dl.mul_hat(e, e, out=e))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)
    med = ms[len(ms) // 2]
    res[name] = {"ms_med": med, "per_s": (n // 2) / (med * 1e-3), "GBps_2048B": (n // 2) * 2048 / (med * 1e-3) / 1e9}
print(json.dumps(res, indent=1))
