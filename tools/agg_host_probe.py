"""Development probe: host-side behaviour of the windowed aggregate / verify pipeline (allocator activity, per-call wall time)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fusion_cryptography_b200.batch import Engine

na = int(sys.argv[1]) if len(sys.argv) > 1 else 1250
eng = Engine(256, seed=42); ctx = eng.ctx; dev = ctx.device
cap = ctx.capacity; nk = 65536
seeds = (10**6 + 2 * np.arange(nk)).astype(np.uint64)
msgs = [f"msg-{i:07d}".encode() for i in range(nk)]
vk_l, vk_r, sigs = eng.keygen_sign(seeds, msgs)
base = (np.arange(na)[:, None] * 97 + np.arange(cap)[None, :]) % nk
sel = torch.from_numpy(base.reshape(-1).astype(np.int32)).to(dev)
dg = ctx.hash_msg(msgs)[sel.long()].contiguous()
counts = [cap] * na
aggs = eng.aggregate(vk_l, vk_r, None, sigs, counts, sel=sel, sig_sel=sel, digests=dg)
for it in range(3):
    st0 = torch.cuda.memory_stats()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    codes = eng.verify(vk_l, vk_r, None, aggs, counts, sel=sel, digests=dg)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    st1 = torch.cuda.memory_stats()
    print(f"verify {it}: host returned after {t1 - t0:.3f} s, device done after {t2 - t0:.3f} s; cudaMalloc calls {st1['num_device_alloc'] - st0['num_device_alloc']}, "
          f"cudaFree {st1['num_device_free'] - st0['num_device_free']}, retries {st1['num_alloc_retries'] - st0['num_alloc_retries']}, reserved {st1['reserved_bytes.all.current'] / 2**30:.1f} GiB")
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    a2 = eng.aggregate(vk_l, vk_r, None, sigs, counts, sel=sel, sig_sel=sel, digests=dg)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"aggregate {it}: {t2 - t0:.3f} s")
