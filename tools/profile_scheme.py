"""Development probe for ncu: one launch of every scheme kernel at a size that fills the GPU (secpar 256).

    python tools/profile_scheme.py            # plain run
    ncu --set full -k regex:<kernel> ... python tools/profile_scheme.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from fusion_cryptography_b200.batch import Engine

eng = Engine(256, seed=42)
ctx = eng.ctx
dev = ctx.device
nk = 32768
seeds = (10**6 + 2 * np.arange(nk)).astype(np.uint64)
msgs = ctx.pack_messages([f"msg-{i:07d}".encode() for i in range(nk)])
# keygen (sample_kernel, ntt_fwd, matvec) + sign (sponge_bytes, hash_ch_stream, ntt_fwd, sign_mac)
vk_l, vk_r, sigs = eng.keygen_sign(seeds, msgs, chunk=nk)
torch.cuda.synchronize()
# windowed aggregate / verify: a warp per aggregate (few aggregates), a lane pair per aggregate (the middle range) and a thread per aggregate (many)
dg = ctx.hash_msg(msgs)
for impl, na, ns in ((3, 1184, 96), (4, 9472, 12), (2, 9472, 12)):
    ctx.set_option("hash_ag_impl", impl)
    base = (np.arange(na)[:, None] * 97 + np.arange(ns)[None, :]) % nk
    sel = torch.from_numpy(base.reshape(-1).astype(np.int32)).to(dev)
    dgp = dg[sel.long()].contiguous()
    aggs = eng.aggregate(vk_l, vk_r, None, sigs, [ns] * na, sel=sel, sig_sel=sel, digests=dgp, max_positions=na * ns)
    codes = eng.verify(vk_l, vk_r, None, aggs, [ns] * na, sel=sel, digests=dgp, max_positions=na * ns)
    torch.cuda.synchronize()
    assert bool((codes == 0).all()), impl
ctx.set_option("hash_ag_impl", 0)
print("ok")
