"""Development probe: what the box's host links give -- bare pinned-memory copies of 1 GiB per direction, per rank.

    python tools/pcie_probe.py                                                               # one GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29544 tools/pcie_probe.py     # N ranks AT ONCE

With N ranks every rank copies at the same time (barrier before, max over ranks after): the aggregate is the ceiling of any end-to-end leg that
moves its data through host memory (bench.py reports the same figure next to its e2e line as e2e.pcie_ceiling)."""
import os
import time

import torch

rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 28  # 1 GiB int32
h_in = torch.empty(n, dtype=torch.int32, pin_memory=True)
h_out = torch.empty(n, dtype=torch.int32, pin_memory=True)
d_a = torch.empty(n, dtype=torch.int32, device="cuda")
d_b = torch.empty(n, dtype=torch.int32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


def both():
    h2d()
    d2h()


for name, fn, gib in (("h2d", h2d, 1), ("d2h", d2h, 1), ("both", both, 2)):
    fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t = time.perf_counter()
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / 5
    if world > 1:
        x = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(x, op=dist.ReduceOp.MAX)
        dt = float(x.item())
    if rank == 0:
        print(f"{world} rank(s) at once, {name}: {dt * 1e3:.2f} ms per GiB and direction and rank -> {world * gib * 1.0737 / dt:.1f} GB/s over all ranks "
              f"({gib * 1.0737 / dt:.1f} GB/s per rank)", flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
