"""One process, all visible GPUs: host-pointer cb200_mlkem_encaps sharded by index inside the library.

    python scripts/one_process_ngpu.py [log2 ops per GPU, default 19] > gpurun_out/r02_one_process.json

Host buffers are pinned (torch pin_memory) and filled before the timed region; the timed call is the public
EncapsulateBatch on numpy views of them, so host<->device copies of every shard are inside the timing.  A strided sample
of the result is compared with the oracle (tests-only code: this script is measurement infrastructure, like bench.py's
cpu_baseline leg).
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import circl_b200  # noqa: E402
from circl_b200 import mlkem  # noqa: E402
import oracle  # noqa: E402


def main():
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 19
    out = {"script": "one_process_ngpu", "log2_ops_per_gpu": lg, "runs": []}
    visible = circl_b200.device_count()
    for ndev in [d for d in (1, 2, 4, 8) if d <= visible]:
        for place in ("torch pin_memory (allocating thread's node)", "cb200_host_alloc_batch (each shard next to its GPU)"):
            got = circl_b200.init_devices(ndev)
            assert got == ndev
            scheme = mlkem.ByName("ML-KEM-768")
            n = ndev << lg
            rng = np.random.default_rng(5)
            pool, _ = scheme.DeriveKeyPairBatch(rng.integers(0, 256, size=(1024, 64), dtype=np.uint8))
            shapes = [(n, scheme.PublicKeySize()), (n, 32), (n, scheme.CiphertextSize()), (n, 32)]
            if place.startswith("torch"):
                keep = [torch.empty(sh, dtype=torch.uint8, pin_memory=True) for sh in shapes]
                eks, seeds, ct, ss = [t.numpy() for t in keep]
            else:
                keep = [circl_b200.host_batch(*sh) for sh in shapes]
                eks, seeds, ct, ss = keep
            idx = np.arange(n) % 1024
            for lo in range(0, n, 1 << 18):
                eks[lo:lo + (1 << 18)] = pool[idx[lo:lo + (1 << 18)]]
            seeds[:] = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
            for _ in range(2):
                scheme.EncapsulateBatch(eks, seeds, ct=ct, ss=ss)
            times = []
            for _ in range(4):
                t0 = time.perf_counter()
                scheme.EncapsulateBatch(eks, seeds, ct=ct, ss=ss)
                times.append(time.perf_counter() - t0)
            ok = True
            for i in list(range(0, n, max(1, n // 61))) + [n - 1]:
                wct, wss = oracle.mlkem_encaps(3, eks[i].tobytes(), seeds[i].tobytes())
                ok = ok and ct[i].tobytes() == wct and ss[i].tobytes() == wss
            best = min(times)
            out["runs"].append({"devices": ndev, "host_buffers": place, "ops": n, "ms_per_call_best": 1e3 * best,
                                "ms_per_call_all": [round(1e3 * t, 3) for t in times], "encaps_per_s": n / best,
                                "outputs_match_oracle": bool(ok), "h2d_bytes": int(n * (eks.shape[1] + 32)),
                                "d2h_bytes": int(n * (ct.shape[1] + 32))})
            if not place.startswith("torch"):
                for a in keep:
                    circl_b200.host_free(a)
            del keep, eks, seeds, ct, ss
            circl_b200.shutdown()
    base = out["runs"][0]["encaps_per_s"]
    for r in out["runs"]:
        r["efficiency_vs_1"] = r["encaps_per_s"] / (base * r["devices"])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
