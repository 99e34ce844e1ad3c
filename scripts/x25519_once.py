"""One X25519 Shared batch on device tensors (development aid for an ncu capture of x25519_kernel)."""
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
print(n, bool(ok.all()))
