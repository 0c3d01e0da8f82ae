"""Development probe: fz_ntt_pass with the Shoup quotient estimate of some long-stride stages on the FP64 pipe (fz_set_option "ntt_pass_f64").
Every mask must give the bytes of mask 0; prints the time of each over 2^20 secpar-256 (2^22 secpar-128) polynomials.

Needs the variant kernels: `git apply tools/ntt_pass_f64.patch && python -m fusion_cryptography_b200.build --force` (measured 1-3 % SLOWER than the
integer Shoup pair on B200, profiles/r02_pass_f64_probe.txt, so the library does not carry them)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fusion_cryptography_b200 import native
for secpar, n in ((256, 1 << 20), (128, 1 << 22)):
    ctx = native.get_context(secpar); d = ctx.d
    h = (native.Q - 1) // 2
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    a = torch.randint(-h, h + 1, (n, d), dtype=torch.int32, device="cuda", generator=g)
    a[0, :] = h; a[1, :] = -h; a[2, :] = 0; a[3, :] = torch.iinfo(torch.int32).max; a[4, :] = torch.iinfo(torch.int32).min   # any int32 is a legal input
    b = ctx.ntt_fwd(torch.randint(-h, h + 1, (n, d), dtype=torch.int32, device="cuda", generator=g))
    ref, o = torch.empty_like(a), torch.empty_like(a)
    ctx.set_option("ntt_pass_f64", 0)
    ctx.ntt_pass(a, b, out=ref)
    for rnd in range(2):
        for mask in (0, 0x7f, 0x0f, 0x70, 0x33, 0x55):
            ctx.set_option("ntt_pass_f64", mask)
            for _ in range(5): ctx.ntt_pass(a, b, out=o)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(30): ctx.ntt_pass(a, b, out=o)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 30
            print(f"secpar {secpar} round {rnd}: mask {mask:#04x}  {ms:.4f} ms  {n / ms / 1e6:.3f} G polys/s  identical to mask 0: {bool(torch.equal(o, ref))}", flush=True)
    ctx.set_option("ntt_pass_f64", 0)
