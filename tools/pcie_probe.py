import torch, time
n = 1 << 28  # 1 GiB int32
h_in = torch.empty(n, dtype=torch.int32).pin_memory(); h_out = torch.empty(n, dtype=torch.int32).pin_memory()
d_a = torch.empty(n, dtype=torch.int32, device="cuda"); d_b = torch.empty(n, dtype=torch.int32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
def h2d():
    with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
for name, fn, gib in (("h2d", h2d, 1), ("d2h", d2h, 1), ("both", both, 2)):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / 5
    print(f"{name}: {dt*1e3:.2f} ms per GiB-direction -> {gib * 1.0737 / dt:.1f} GB/s total")
