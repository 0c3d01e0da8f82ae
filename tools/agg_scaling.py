"""Development probe: hashing-pipeline throughput vs number of aggregates in flight (secpar 256, capacity-sized)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fusion_cryptography_b200.batch import Engine

eng = Engine(256, seed=42)
ctx = eng.ctx
cap = ctx.capacity
nk = 8192
seeds = (10**6 + 2 * np.arange(nk)).astype(np.uint64)
kb = eng.keygen(seeds)
dg = torch.randint(0, 256, (nk, 32), dtype=torch.uint8, device="cuda")

def timed(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return time.perf_counter() - t0, r

for n in (1 << 13, 1 << 16, 1 << 18):
    idx = torch.arange(n, device="cuda") % nk
    vl, vr, d8 = kb.vk_l[idx].contiguous(), kb.vk_r[idx].contiguous(), dg[idx].contiguous()
    for mode in (0, 1):
        if mode == 1 and n > (1 << 16): continue
        ctx.set_option("buffered_hash", mode)
        ctx.hash_ch(vl[:64], vr[:64], d8[:64])
        t, c = timed(lambda: ctx.hash_ch(vl, vr, d8))
        print(f"hash_ch n={n} mode={'buffered' if mode else 'stream'}: {t*1e3:.1f} ms -> {n/t/1e6:.2f} M/s", flush=True)
ctx.set_option("buffered_hash", 0)
for na in [int(a) for a in (sys.argv[1:] or [32, 512, 4096])]:
    tot = na * cap
    idx = (torch.arange(na, device="cuda")[:, None] * 97 + torch.arange(cap, device="cuda")[None, :]).reshape(-1) % nk
    vl, vr, d8 = kb.vk_l[idx].contiguous(), kb.vk_r[idx].contiguous(), dg[idx].contiguous()
    c = ctx.hash_ch(vl, vr, d8)
    counts = [cap] * na
    for name, mode, impl in (("warp", 0, 3), ("thread", 0, 2), ("buffered", 1, 0)):
        if mode == 1 and na > 512: continue
        ctx.set_option("buffered_hash", mode); ctx.set_option("hash_ag_impl", impl)
        t, a = timed(lambda: ctx.hash_ag(vl, vr, d8, c, counts))
        print(f"hash_ag aggregates={na} impl={name}: {t:.2f} s -> {tot/t/1e6:.3f} M sigs/s", flush=True)
    ctx.set_option("buffered_hash", 0); ctx.set_option("hash_ag_impl", 0)
    del vl, vr, d8, c
