"""Development aid: bit-exactness and device time of the Kyber NTT kernel variants (CB200_NTT_VARIANT /
CB200_INVNTT_VARIANT select the kernel inside launch_kyber_ntt).  bench.py is the contract; this only ranks variants."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import circl_b200
import oracle
from circl_b200 import kyber

circl_b200.init(0)
Q = 3329
g = torch.Generator(device="cuda").manual_seed(7)


def rnd(n, lo, hi):
    return torch.randint(lo, hi + 1, (n, 256), device="cuda", dtype=torch.int32, generator=g).to(torch.int16)


def run(x, inverse, var):
    os.environ["CB200_INVNTT_VARIANT" if inverse else "CB200_NTT_VARIANT"] = str(var)
    y = x.clone()
    (kyber.inv_ntt_ if inverse else kyber.ntt_)(y)
    torch.cuda.synchronize()
    return y


res = {"exact": {}, "ms": {}}
n = (1 << 16) + 5
cases = {
    "contract |c|<=q": rnd(n, -Q, Q),
    "normalised [0,q)": rnd(n, 0, Q - 1),
    "any int16": rnd(n, -32768, 32767),
    "edge of the fast range": None,
    "mixed": None,
}
for inverse, bound in ((0, 13561), (1, 3679)):
    cases["edge of the fast range"] = rnd(n, -bound - 1, bound + 1)
    m = rnd(n, -bound, bound)
    m[::7] = rnd((n + 6) // 7, -32768, 32767)
    m[:, 255] = bound
    m[5::11, 3] = bound + 1
    m[6::11, 128] = -bound - 1
    cases["mixed"] = m
    for cname, x in cases.items():
        ref = run(x, inverse, 0)
        idx = torch.arange(0, n, max(1, n // 1024), device="cuda")
        want = x[idx].cpu().numpy().copy()
        oracle.kyber_ntt_inplace_mt(want, bool(inverse), 4)
        res["exact"][f"{'inv' if inverse else 'fwd'} v0 {cname} vs oracle"] = bool(np.array_equal(ref[idx].cpu().numpy(), want))
        for var in (1, 2, 3):
            y = run(x, inverse, var)
            ok = bool(torch.equal(y, ref))
            res["exact"][f"{'inv' if inverse else 'fwd'} v{var} {cname}"] = ok

# oracle spot check of the general kernel on the same data (the GPU suite does this at length)
x = cases["contract |c|<=q"][:4096]
for inverse in (0, 1):
    y = run(x, inverse, 1).cpu().numpy()
    w = x.cpu().numpy().copy()
    oracle.kyber_ntt_inplace_mt(w, bool(inverse), 4)
    res["exact"][f"{'inv' if inverse else 'fwd'} v1 vs oracle"] = bool(np.array_equal(y, w))

npoly = 1 << 20
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for inverse in (0, 1):
    for dist, src in (("contract", rnd(npoly, -Q, Q)), ("any int16", rnd(npoly, -32768, 32767))):
        d = src.clone()
        for var in (0, 1, 2, 3):
            os.environ["CB200_INVNTT_VARIANT" if inverse else "CB200_NTT_VARIANT"] = str(var)
            fn = kyber.inv_ntt_ if inverse else kyber.ntt_
            ts = []
            for i in range(13):
                d.copy_(src)
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn(d)
                b.record()
                torch.cuda.synchronize()
                if i >= 3:
                    ts.append(a.elapsed_time(b))
            ts.sort()
            med = ts[len(ts) // 2]
            res["ms"][f"{'inv' if inverse else 'fwd'} v{var} {dist}"] = {
                "med": round(med, 4), "min": round(ts[0], 4), "hbm_frac": round(npoly * 1024 / (med * 1e-3) / 1e9 / 6574.5, 3)}
print(json.dumps(res, indent=1), flush=True)
circl_b200.shutdown()
