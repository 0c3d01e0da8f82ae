"""Development probe: per-stage device times of the windowed aggregate / verify pipeline at secpar 256.

    python tools/agg_stages.py [n_aggs] [window_positions]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from fusion_cryptography_b200.batch import Engine

na = int(sys.argv[1]) if len(sys.argv) > 1 else 1250
wp = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 18
eng = Engine(256, seed=42)
ctx = eng.ctx
dev = ctx.device
cap, d, ell = ctx.capacity, eng.d, eng.ell
nk = 16384
seeds = (10**6 + 2 * np.arange(nk)).astype(np.uint64)
msgs = [f"msg-{i:07d}".encode() for i in range(nk)]
vk_l, vk_r, sigs = eng.keygen_sign(seeds, msgs, chunk=8192)
base = (np.arange(na)[:, None] * 97 + np.arange(cap)[None, :]) % nk
sel = torch.from_numpy(base.reshape(-1).astype(np.int32)).to(dev)
dg_pos = ctx.hash_msg(msgs)[sel.long()].contiguous()
counts = [cap] * na


def timed(fn, reps=1):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, out


# whole pipeline, second call (allocator warm)
eng.verify(vk_l, vk_r, None, eng.aggregate(vk_l, vk_r, None, sigs, counts[:8], sel=sel[:8 * cap], sig_sel=sel[:8 * cap], digests=dg_pos[:8 * cap]),
           counts[:8], sel=sel[:8 * cap], digests=dg_pos[:8 * cap])
ms, aggs = timed(lambda: eng.aggregate(vk_l, vk_r, None, sigs, counts, sel=sel, sig_sel=sel, digests=dg_pos, max_positions=wp))
ms, aggs = timed(lambda: eng.aggregate(vk_l, vk_r, None, sigs, counts, sel=sel, sig_sel=sel, digests=dg_pos, max_positions=wp))
print(f"aggregate {na} x {cap}: {ms:9.2f} ms  -> {na * cap / ms / 1e3:.2f} M sigs/s")
ms, codes = timed(lambda: eng.verify(vk_l, vk_r, None, aggs, counts, sel=sel, digests=dg_pos, max_positions=wp))
print(f"verify (cold allocator)   : {ms:9.2f} ms")
ms, codes = timed(lambda: eng.verify(vk_l, vk_r, None, aggs, counts, sel=sel, digests=dg_pos, max_positions=wp))
print(f"verify    {na} x {cap}: {ms:9.2f} ms  -> {na * cap / ms / 1e3:.2f} M sigs/s   ok={bool((codes == 0).all())}")

# stages of one window, serial
P = max(1, min(cap, wp // na))
n_win = na * P
agg_off = torch.arange(na + 1, dtype=torch.int64, device=dev) * cap
win_off = torch.arange(na + 1, dtype=torch.int64, device=dev) * P
ms, perm = timed(lambda: ctx.sort_keys(vk_l, vk_r, counts, sel=sel))
print(f"sort (all positions)            {ms:9.3f} ms")
idx = (agg_off[:-1, None] + torch.arange(P, device=dev)[None, :]).reshape(-1)
pos = perm.index_select(0, idx)
ms, (vl_w, vr_w, dg_w) = timed(lambda: ctx.gather_sorted(vk_l, vk_r, dg_pos, pos, sel=sel))
print(f"gather window ({n_win} pos)     {ms:9.3f} ms")
ms, c_w = timed(lambda: ctx.hash_ch(vl_w, vr_w, dg_w))
print(f"hash_ch window                  {ms:9.3f} ms  -> {n_win / ms / 1e3:.2f} M/s")
agg_of = torch.arange(na, dtype=torch.int32, device=dev).repeat_interleave(P)
state = ctx.hash_ag_begin(win_off, na)          # aggregates of exactly P signers: the window closes them
ring, ring_cap, text_pos = ctx.hash_ag_ring(na, P)
ms, wend = timed(lambda: ctx.hash_ag_text(vl_w, vr_w, dg_w, c_w, win_off, agg_of, win_off, na, n_win, 0, text_pos.zero_(), ring, ring_cap))
print(f"hash_ag text window             {ms:9.3f} ms  -> {n_win / ms / 1e3:.2f} M signers/s")
ms, _ = timed(lambda: ctx.hash_ag_absorb_ring(state, ring, ring_cap, wend, win_off, na))
print(f"hash_ag absorb (ring) (P={P})   {ms:9.3f} ms  -> {ms * 1e3 / P:.1f} us per signer position")
xof, xof_cap = ctx.hash_ag_xof_buffer(na, P)
ms, bpos = timed(lambda: ctx.hash_ag_squeeze_xof(state, win_off, na, xof, xof_cap))
print(f"hash_ag squeeze (xof)           {ms:9.3f} ms  -> {ms * 1e3 / P:.1f} us per signer position")
ms, alpha_w = timed(lambda: ctx.hash_ag_decode(xof, xof_cap, bpos, win_off, agg_of, n_win))
print(f"hash_ag decode + NTT window     {ms:9.3f} ms")
rows = sel.index_select(0, pos.long()).contiguous()
ms, part = timed(lambda: ctx.aggregate_sum_off(sigs, rows, alpha_w, win_off, na, P))
print(f"aggregate_sum window            {ms:9.3f} ms  -> {n_win * (ell + 1) * d * 4 / ms / 1e6:.0f} GB/s")
ms, _ = timed(lambda: ctx.verify_target(vl_w, vr_w, c_w, alpha_w, win_off, na))
print(f"verify_target window            {ms:9.3f} ms")
nwin = (cap + P - 1) // P
print(f"windows: {nwin}")
