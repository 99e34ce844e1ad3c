"""One ML-DSA-65 SignBatch call on per-op keys (development aid for ncu captures of the sign kernels)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hashlib
import numpy as np
import circl_b200
from circl_b200 import mldsa

circl_b200.init(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 15
s = mldsa.ByName("ML-DSA-65")
seeds = np.frombuffer(b"".join(hashlib.shake_256(b"k%d" % j).digest(32) for j in range(64)), dtype=np.uint8).reshape(64, 32)
pk, sk = s.DeriveKeyBatch(seeds)
sks = np.ascontiguousarray(np.tile(sk, (n // 64, 1)))
msgs = [hashlib.shake_256(b"m%d" % i).digest(32) for i in range(n)]
sig, att = s.SignBatch(sks, msgs, return_attempts=True)
print("attempts per signature", att / n, hashlib.sha256(sig.tobytes()).hexdigest())
