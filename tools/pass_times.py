"""Development probe: the NTT pass as three kernels vs the fused fz_ntt_pass, 2^20 secpar-256 polynomials."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fusion_cryptography_b200 import native
for secpar, n in ((256, 1 << 20), (128, 1 << 22)):
    ctx = native.get_context(secpar); d = ctx.d
    h = (native.Q - 1) // 2
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    a = torch.randint(-h, h + 1, (n, d), dtype=torch.int32, device="cuda", generator=g)
    b = ctx.ntt_fwd(torch.randint(-h, h + 1, (n, d), dtype=torch.int32, device="cuda", generator=g))
    t1, t2, o1, o2 = torch.empty_like(a), torch.empty_like(a), torch.empty_like(a), torch.empty_like(a)
    def three():
        ctx.ntt_fwd(a, out=t1); ctx.pointwise_mul(t1, b, out=t2); ctx.ntt_inv(t2, out=o1)
    def fused():
        ctx.ntt_pass(a, b, out=o2)
    for name, fn in (("three kernels", three), ("fused", fused)):
        for _ in range(5): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"secpar {secpar}: {name:14s} {ms:.3f} ms  -> {n / ms / 1e6:.3f} G polys/s")
    print("bit-identical:", bool(torch.equal(o1, o2)))
