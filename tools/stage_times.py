"""Development probe: per-stage device times of keygen + sign at secpar 256."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fusion_cryptography_b200.batch import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
eng = Engine(256, seed=42); ctx = eng.ctx
seeds = (10**6 + 2 * np.arange(n)).astype(np.uint64)
msgs = ctx.pack_messages([f"msg-{i:07d}".encode() for i in range(n)])
def t(fn, reps=5):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): r = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, r
ms, kb = t(lambda: eng.keygen(seeds)); print(f"keygen total        {ms:8.3f} ms")
ms, co = t(lambda: ctx.sample_coef([int(s) for s in seeds[:4096]])); print(f"sample_coef (4096, host key packing incl.) {ms:8.3f} ms")
import ctypes as C
from fusion_cryptography_b200.native import _ptr
sd = torch.from_numpy(seeds.view(np.int64)).cuda()
ms, x = t(lambda: ctx.ntt_fwd(kb.sk_hat)); print(f"ntt_fwd 2n polys    {ms:8.3f} ms")
ms, x = t(lambda: ctx.hash_msg(msgs)); print(f"hash_msg            {ms:8.3f} ms"); dg = x
ms, c = t(lambda: ctx.hash_ch(kb.vk_l, kb.vk_r, dg)); print(f"hash_ch             {ms:8.3f} ms")
sig = torch.empty((n, eng.ell, eng.d), dtype=torch.int32, device="cuda")
ms, x = t(lambda: ctx.sign_mac_rows(kb.sk_hat, c, out=sig)); print(f"sign_mac_rows       {ms:8.3f} ms")
ms, x = t(lambda: eng.keygen_sign(seeds, msgs, chunk=16384)); print(f"keygen_sign total   {ms:8.3f} ms  -> {n/ms*1e3/1e6:.2f} M sign/s")

big = torch.empty(1 << 30, dtype=torch.int32, device="cuda")
ms, _ = t(lambda: big.zero_()); print(f"memset 4 GiB        {ms:8.3f} ms  -> {4 * 1.073741824 / ms * 1e3:.0f} GB/s write-only")
