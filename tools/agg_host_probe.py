"""Development probe: host-side behaviour of the windowed aggregate / verify pipeline (allocator activity, per-call wall time)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fusion_cryptography_b200.batch import Engine

na = int(sys.argv[1]) if len(sys.argv) > 1 else 1250
from fusion_cryptography_b200.batch import Tuning
eng = Engine(256, seed=42, tuning=Tuning(side_priority=int(os.environ.get("FZ_SIDE_PRIO", "0")), text_stream=os.environ.get("FZ_TEXT_STREAM", "1") != "0",
                                            lookahead=int(os.environ.get("FZ_LOOKAHEAD", "2")))); ctx = eng.ctx; dev = ctx.device
if os.environ.get("FZ_AG_IMPL"): ctx.set_option("hash_ag_impl", int(os.environ["FZ_AG_IMPL"]))
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

# where does the time beyond the bare chain go?  events around every chain kernel inside the real pipeline
ev = []
orig_a, orig_s = ctx.hash_ag_absorb_ring, ctx.hash_ag_squeeze_xof
def wrap(fn, tag):
    def f(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(*a, **k); e1.record(); ev.append((tag, e0, e1)); return r
    return f
ctx.hash_ag_absorb_ring = wrap(orig_a, "absorb"); ctx.hash_ag_squeeze_xof = wrap(orig_s, "squeeze")
torch.cuda.synchronize(); t0 = time.perf_counter()
e_begin = torch.cuda.Event(enable_timing=True); e_begin.record()
codes = eng.verify(vk_l, vk_r, None, aggs, counts, sel=sel, digests=dg)
e_end = torch.cuda.Event(enable_timing=True); e_end.record()
torch.cuda.synchronize()
tot = {"absorb": 0.0, "squeeze": 0.0}
for tag, e0, e1 in ev: tot[tag] += e0.elapsed_time(e1)
first = ev[0][1]; last = ev[-1][2]
print(f"verify total {e_begin.elapsed_time(e_end):.1f} ms: chain kernels absorb {tot['absorb']:.1f} + squeeze {tot['squeeze']:.1f} ms; "
      f"before the first chain kernel {e_begin.elapsed_time(first):.1f} ms; after the last {last.elapsed_time(e_end):.1f} ms; "
      f"gaps inside {first.elapsed_time(last) - tot['absorb'] - tot['squeeze']:.1f} ms")
for i, (tag, e0, e1) in enumerate(ev):
    if i < 4 or i > len(ev) - 4: print(f"   {tag} {i}: {e0.elapsed_time(e1):.2f} ms")
gaps = [(ev[i][0], ev[i + 1][0], ev[i][2].elapsed_time(ev[i + 1][1])) for i in range(len(ev) - 1)]
print("gaps between consecutive chain kernels (ms):", " ".join(f"{g[2]:.1f}" for g in gaps))

ctx.hash_ag_absorb_ring, ctx.hash_ag_squeeze_xof = orig_a, orig_s
pev = []
def wrap2(fn, tag):
    def f(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(*a, **k); e1.record(); pev.append((tag, e0, e1)); return r
    return f
o1, o2, o3 = ctx.hash_ch, ctx.hash_ag_text, ctx.gather_sorted
ctx.hash_ch, ctx.hash_ag_text, ctx.gather_sorted = wrap2(o1, "hash_ch"), wrap2(o2, "text"), wrap2(o3, "gather")
codes = eng.verify(vk_l, vk_r, None, aggs, counts, sel=sel, digests=dg)
torch.cuda.synchronize()
print("side-stream stages inside the pipeline (ms):", " ".join(f"{t}:{a.elapsed_time(b):.1f}" for t, a, b in pev[:12]))
ctx.hash_ch, ctx.hash_ag_text, ctx.gather_sorted = o1, o2, o3

# one common timeline: chain kernels (main stream) and side-stream stages, ms since the start of the call
ev.clear(); pev.clear()
ctx.hash_ag_absorb_ring, ctx.hash_ag_squeeze_xof = wrap(orig_a, "absorb"), wrap(orig_s, "squeeze")
ctx.hash_ch, ctx.hash_ag_text, ctx.gather_sorted = wrap2(o1, "hash_ch"), wrap2(o2, "text"), wrap2(o3, "gather")
torch.cuda.synchronize()
base_ev = torch.cuda.Event(enable_timing=True); base_ev.record()
codes = eng.verify(vk_l, vk_r, None, aggs, counts, sel=sel, digests=dg)
torch.cuda.synchronize()
rows_ = [(base_ev.elapsed_time(a), base_ev.elapsed_time(b), t) for t, a, b in ev + pev]
rows_.sort()
for a, b, t in rows_[:26]:
    print(f"  {a:8.1f} -> {b:8.1f}  {t}")

