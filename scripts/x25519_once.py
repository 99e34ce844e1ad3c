"""X25519 KeyGen + Shared on device tensors, timed with CUDA events (development aid; also the ncu target for x25519_kernel)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import circl_b200
from circl_b200 import hybrid

circl_b200.init(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 17
g = torch.Generator(device="cuda").manual_seed(3)
k = torch.randint(0, 256, (n, 32), generator=g, device="cuda", dtype=torch.uint8)
p = hybrid.x25519_keygen(k)
s, ok = hybrid.x25519_shared(k, p)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    hybrid.x25519_keygen(k)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 3
print(n, bool(ok.all()), "KeyGen %.3f ms per batch, %.3e /s" % (ms, n / (ms * 1e-3)))
a.record()
for _ in range(3):
    hybrid.x25519_shared(k, p)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 3
print(n, "Shared %.3f ms per batch, %.3e /s" % (ms, n / (ms * 1e-3)))
