"""Development probe: fz_polymul (INTT(NTT(a) o NTT(b)) in one kernel) on 2^20 secpar-256 / 2^22 secpar-128 polynomials."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fusion_cryptography_b200 import native
for secpar, n in ((256, 1 << 20), (128, 1 << 22)):
    ctx = native.get_context(secpar); d = ctx.d; h = (native.Q - 1) // 2
    a = torch.randint(-h, h + 1, (n, d), dtype=torch.int32, device="cuda")
    b = torch.randint(-h, h + 1, (n, d), dtype=torch.int32, device="cuda")
    o = torch.empty_like(a)
    for _ in range(5): ctx.polymul(a, b, out=o)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ctx.polymul(a, b, out=o)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    ref = ctx.ntt_inv(ctx.pointwise_mul(ctx.ntt_fwd(a), ctx.ntt_fwd(b)))
    print(f"secpar {secpar}: polymul {ms:.3f} ms -> {n / ms / 1e6:.3f} G products/s, equals the four kernels: {bool(torch.equal(o, ref))}")
