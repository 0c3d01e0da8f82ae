"""Development probe: throughput of the general ring kernels (fz_ring_*) on a few rings, 2^28 bytes of coefficients per operand."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fusion_cryptography_b200 import native


def table(w, d, q):
    bits = d.bit_length() - 1
    p = [pow(w, i, q) for i in range(d)]
    return [p[int(format(i, f"0{bits}b")[::-1], 2)] for i in range(d)]


def root_of(q, order):
    g = 2
    while True:
        x = pow(g, (q - 1) // order, q)
        if pow(x, order // 2, q) != 1:
            return x
        g += 1


def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


if __name__ == "__main__":
  for d, q in ((64, 257), (256, 7681), (256, 2013265921), (1024, 12289), (4096, 2013265921), (8192, 2013265921), (16384, 2013265921)):
      root = root_of(q, 2 * d)
      r = native.get_ring(q, d, table(root, d, q), table(pow(root, q - 2, q), d, q))
      n = (1 << 26) // d          # 2^26 coefficients = 256 MiB per operand (larger than the 126 MB L2)
      g = torch.Generator(device="cuda"); g.manual_seed(1)
      a = torch.randint(-(q // 2), q // 2 + 1, (n, d), dtype=torch.int32, device="cuda", generator=g)
      b = torch.randint(-(q // 2), q // 2 + 1, (n, d), dtype=torch.int32, device="cuda", generator=g)
      o = torch.empty_like(a)
      byt = a.numel() * 4
      res = []
      for name, fn, traffic in (("ntt_fwd", lambda: r.ntt_fwd(a, out=o), 2 * byt), ("ntt_inv", lambda: r.ntt_inv(a, out=o), 2 * byt),
                                ("pointwise_mul", lambda: r.pointwise_mul(a, b, out=o), 3 * byt), ("pointwise_add", lambda: r.pointwise_add(a, b, out=o), 3 * byt)):
          ms = timed(fn)
          res.append(f"{name} {ms:.3f} ms {traffic / ms / 1e6:.0f} GB/s")
      print(f"d={d} q={q} n={n}: " + "; ".join(res))
