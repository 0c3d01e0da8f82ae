"""Development probe: the signature-row kernel (sig[b][r] = sk_hat[b][0] * c_hat[b] + sk_hat[b][1] for every row r: ell KiB written per key).
Round 2 measured variants behind a development switch "rows_impl" (since removed): the shipped kernel 5.64 TB/s (secpar 256, 189,440 keys) / 6.16 TB/s (secpar 128);
one key per CTA step, every store instruction of the CTA one contiguous 4 KiB piece: 5.28 / 5.65 / 5.83 TB/s at 8 / 16 / 32 CTAs per SM; plain, .cs, .wt and .cg
stores instead of L1::no_allocate: 5.55-5.63 TB/s.  torch's zero fill of the same buffer runs at 7.45 TB/s, but that is constant data (compressible in L2 / DRAM);
the copy peak of MEASURED_PEAKS.json (6.53 TB/s read + write) is the fair yardstick: the kernel is at 87 % of it and nothing tried moved it."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fusion_cryptography_b200.batch import Engine
for secpar in (256, 128):
    eng = Engine(secpar, seed=42); ctx = eng.ctx
    n = 189440 if secpar == 256 else 400000
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    h = (2147465729 - 1) // 2
    sk_hat = torch.randint(-h, h + 1, (n, 2, eng.d), dtype=torch.int32, device="cuda", generator=g)
    c_hat = torch.randint(-h, h + 1, (n, eng.d), dtype=torch.int32, device="cuda", generator=g)
    sig = torch.empty((n, eng.ell, eng.d), dtype=torch.int32, device="cuda")
    ref = None
    for impl in (0,):
        ctx.sign_mac_rows(sk_hat, c_hat, out=sig); torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ctx.sign_mac_rows(sk_hat, c_hat, out=sig); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        chk = (int(sig[:1000].sum().item()), int(sig[-1000:].sum().item()), int(sig[n // 2].sum().item()))
        if ref is None: ref = chk
        print(f"secpar {secpar} rows_impl {impl}: {best:.3f} ms = {n * eng.ell * eng.d * 4 / best / 1e9:.2f} TB/s  same={chk == ref}", flush=True)
    z = torch.empty_like(sig)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    z.zero_(); torch.cuda.synchronize(); a.record(); z.zero_(); b.record(); torch.cuda.synchronize()
    print(f"  memset of the same buffer: {a.elapsed_time(b):.3f} ms = {z.numel() * 4 / a.elapsed_time(b) / 1e9:.2f} TB/s")
    del sig, z
