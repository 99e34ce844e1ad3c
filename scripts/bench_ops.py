"""Device-resident timings of the secondary entry points (CUDA events, >= 3 warm-ups, inputs > L2).
Writes one JSON object; the headline contract lives in bench.py."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import circl_b200
from circl_b200 import dilithium as dl, kyber, mlkem
import bench as B

circl_b200.init(0)
res = {}


def timed(fn, steps=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


for name, k, log2 in (("ML-KEM-768", 3, 20), ("ML-KEM-1024", 4, 20)):
    scheme = mlkem.ByName(name)
    n = 1 << log2
    g = torch.Generator(device="cuda").manual_seed(1)
    seeds = torch.randint(0, 256, (n, 64), generator=g, device="cuda", dtype=torch.uint8)
    ms = torch.randint(0, 256, (n, 32), generator=g, device="cuda", dtype=torch.uint8)
    t = timed(lambda: scheme.DeriveKeyPairBatch(seeds), steps=3)
    ek, dk = scheme.DeriveKeyPairBatch(seeds)
    res[f"{name} keygen"] = {"n": n, "ms": t, "per_s": n / (t * 1e-3)}
    ct = torch.empty((n, scheme.CiphertextSize()), dtype=torch.uint8, device="cuda")
    ss = torch.empty((n, 32), dtype=torch.uint8, device="cuda")
    t = timed(lambda: scheme.EncapsulateBatch(ek, ms, ct=ct, ss=ss))
    res[f"{name} encaps (per-op ek)"] = {"n": n, "ms": t, "per_s": n / (t * 1e-3)}
    pk0 = scheme.UnmarshalBinaryPublicKey(ek[0].cpu().numpy().tobytes())
    t = timed(lambda: scheme.EncapsulateBatch(pk0, ms, ct=ct, ss=ss))
    res[f"{name} encaps (shared ek)"] = {"n": n, "ms": t, "per_s": n / (t * 1e-3)}
    scheme.EncapsulateBatch(ek, ms, ct=ct, ss=ss)
    ss2 = torch.empty_like(ss)
    t = timed(lambda: scheme.DecapsulateBatch(dk, ct, ss=ss2), steps=3)
    res[f"{name} decaps"] = {"n": n, "ms": t, "per_s": n / (t * 1e-3), "roundtrip_ok": bool(torch.equal(ss, ss2))}
    del seeds, ms, ek, dk, ct, ss, ss2
    torch.cuda.empty_cache()

n = 1 << 19
e = torch.randint(0, 8380417, (n, 256), device="cuda", dtype=torch.int32)
for label, fn in (("Dilithium NTT", lambda: dl.ntt_(e)), ("Dilithium InvNTT", lambda: dl.inv_ntt_(e))):
    t = timed(fn, steps=10)
    res[label] = {"n": n, "ms": t, "per_s": n / (t * 1e-3), "GBps": n * 2048 / (t * 1e-3) / 1e9}
del e
torch.cuda.empty_cache()

# wire-side callers: X25519, X-Wing, X25519MLKEM768, round-3 Kyber768, all device-resident
from circl_b200 import hybrid  # noqa: E402
n = 1 << 18
g = torch.Generator(device="cuda").manual_seed(2)
k32 = torch.randint(0, 256, (n, 32), generator=g, device="cuda", dtype=torch.uint8)
t = timed(lambda: hybrid.x25519_keygen(k32), steps=3)
pub = hybrid.x25519_keygen(k32)
res["X25519 KeyGen"] = {"n": n, "ms": t, "per_s": n / (t * 1e-3)}
t = timed(lambda: hybrid.x25519_shared(k32, pub), steps=3)
res["X25519 Shared"] = {"n": n, "ms": t, "per_s": n / (t * 1e-3)}
for name in ("X-Wing", "X25519MLKEM768"):
    sch = hybrid.ByName(name)
    seeds = torch.randint(0, 256, (n, sch.SeedSize()), generator=g, device="cuda", dtype=torch.uint8)
    es = torch.randint(0, 256, (n, sch.EncapsulationSeedSize()), generator=g, device="cuda", dtype=torch.uint8)
    t = timed(lambda: sch.DeriveKeyPairBatch(seeds), steps=3)
    pk, sk = sch.DeriveKeyPairBatch(seeds)
    res[f"{name} keygen"] = {"n": n, "ms": t, "per_s": n / (t * 1e-3)}
    t = timed(lambda: sch.EncapsulateBatch(pk, es), steps=3)
    ct, ss = sch.EncapsulateBatch(pk, es)
    res[f"{name} encaps"] = {"n": n, "ms": t, "per_s": n / (t * 1e-3)}
    t = timed(lambda: sch.DecapsulateBatch(sk, ct), steps=3)
    res[f"{name} decaps"] = {"n": n, "ms": t, "per_s": n / (t * 1e-3),
                             "roundtrip_ok": bool(torch.equal(ss, sch.DecapsulateBatch(sk, ct)))}
    del seeds, es, pk, sk, ct, ss
    torch.cuda.empty_cache()
k768 = mlkem.ByName("Kyber768")
seeds = torch.randint(0, 256, (n, 64), generator=g, device="cuda", dtype=torch.uint8)
ms = torch.randint(0, 256, (n, 32), generator=g, device="cuda", dtype=torch.uint8)
ek, dk = k768.DeriveKeyPairBatch(seeds)
t = timed(lambda: k768.EncapsulateBatch(ek, ms), steps=3)
res["Kyber768 (round 3) encaps"] = {"n": n, "ms": t, "per_s": n / (t * 1e-3)}
del seeds, ms, ek, dk
torch.cuda.empty_cache()

# ML-DSA family, device-resident through the C ABI (CUDA events): key generation, signing and verification for
# ML-DSA-44/65/87 and round-3 Dilithium3; per-op keys, 32-byte messages
from circl_b200 import mldsa  # noqa: E402
from circl_b200._ffi import check, lib  # noqa: E402

L = lib()
n = 1 << 17
g = torch.Generator(device="cuda").manual_seed(9)
dseeds = torch.randint(0, 256, (n, 32), generator=g, device="cuda", dtype=torch.uint8)
msg_d = torch.randint(0, 256, (n * 32 + 8,), generator=g, device="cuda", dtype=torch.uint8)
off_d = (torch.arange(n + 1, device="cuda", dtype=torch.int64) * 32).contiguous()
check(L.cb200_set_stream(torch.cuda.current_stream().cuda_stream))
for name, mode in (("ML-DSA-44", 44), ("ML-DSA-65", 65), ("ML-DSA-87", 87), ("Dilithium3", 3)):
    sch = mldsa.ByName(name)
    pk = torch.empty((n, sch.PublicKeySize()), dtype=torch.uint8, device="cuda")
    sk = torch.empty((n, sch.PrivateKeySize()), dtype=torch.uint8, device="cuda")
    sig = torch.empty((n, sch.SignatureSize()), dtype=torch.uint8, device="cuda")
    st = torch.zeros((n,), dtype=torch.uint8, device="cuda")
    ok = torch.zeros((n,), dtype=torch.uint8, device="cuda")
    t = timed(lambda: check(L.cb200_mldsa_keygen(mode, dseeds.data_ptr(), pk.data_ptr(), sk.data_ptr(), n)), steps=3)
    res[f"{name} keygen"] = {"n": n, "ms": t, "per_s": n / (t * 1e-3)}
    t = timed(lambda: check(L.cb200_mldsa_sign(mode, sk.data_ptr(), sch.PrivateKeySize(), msg_d.data_ptr(), off_d.data_ptr(),
                                               None, 0, None, sig.data_ptr(), st.data_ptr(), n, 0, None)), steps=2, warm=2)
    res[f"{name} sign"] = {"n": n, "ms": t, "per_s": n / (t * 1e-3)}
    t = timed(lambda: check(L.cb200_mldsa_verify(mode, pk.data_ptr(), sch.PublicKeySize(), msg_d.data_ptr(), off_d.data_ptr(),
                                                 None, 0, sig.data_ptr(), ok.data_ptr(), n, 0)), steps=3)
    res[f"{name} verify"] = {"n": n, "ms": t, "per_s": n / (t * 1e-3), "all_valid": bool(ok.all().item())}
    del pk, sk, sig
    torch.cuda.empty_cache()
print(json.dumps(res, indent=1))
