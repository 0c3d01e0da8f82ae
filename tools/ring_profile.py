"""Development probe for ncu: the general-ring transforms at d = 256 (q = 7681), 2^26 coefficients."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fusion_cryptography_b200 import native
from ring_times import table, root_of

d, q = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (256, 7681)
root = root_of(q, 2 * d)
r = native.get_ring(q, d, table(root, d, q), table(pow(root, q - 2, q), d, q))
n = (1 << 26) // d
a = torch.randint(-(q // 2), q // 2 + 1, (n, d), dtype=torch.int32, device="cuda")
o = torch.empty_like(a)
for _ in range(3):
    r.ntt_fwd(a, out=o); r.ntt_inv(a, out=o)
torch.cuda.synchronize()
