"""Development probe for ncu: one hash_ch launch (n signers) and one hash_ag launch (aggregates x signers), secpar 256."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fusion_cryptography_b200.batch import Engine

na, ns, nch = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
eng = Engine(256, seed=42)
ctx = eng.ctx
nk = 4096
kb = eng.keygen((10**6 + 2 * np.arange(nk)).astype(np.uint64))
dg = torch.randint(0, 256, (nk, 32), dtype=torch.uint8, device="cuda")
idx = torch.arange(nch, device="cuda") % nk
c = ctx.hash_ch(kb.vk_l[idx].contiguous(), kb.vk_r[idx].contiguous(), dg[idx].contiguous())
idx = torch.arange(na * ns, device="cuda") % nk
vl, vr, d8 = kb.vk_l[idx].contiguous(), kb.vk_r[idx].contiguous(), dg[idx].contiguous()
cc = ctx.hash_ch(vl, vr, d8)
a = ctx.hash_ag(vl, vr, d8, cc, [ns] * na)
torch.cuda.synchronize()
print("ok", a.shape)
