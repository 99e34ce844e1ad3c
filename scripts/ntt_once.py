"""One forward and one inverse Kyber NTT over 2^20 in-contract polynomials (|c| <= q): the launch ncu captures."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import circl_b200
from circl_b200 import kyber

circl_b200.init(0)
Q = 3329
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
src = (torch.randint(0, 2 * Q + 1, (n, 256), device="cuda", dtype=torch.int32) - Q).to(torch.int16)
for fn in (kyber.ntt_, kyber.inv_ntt_):
    for _ in range(2):
        d = src.clone()
        fn(d)
torch.cuda.synchronize()
circl_b200.shutdown()
