import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fusion_cryptography_b200.batch import Engine
n = 10**6
eng = Engine(256, seed=42); ctx = eng.ctx
seeds = (10**6 + 2 * np.arange(n)).astype(np.uint64)
pm = ctx.pack_messages([f"msg-{i:07d}".encode() for i in range(n)])
for chunk in (94720, 94720 * 2, 94720 * 3, 65536):
    eng.keygen_sign(seeds, pm, chunk=chunk, keep_sigs=False); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); eng.keygen_sign(seeds, pm, chunk=chunk, keep_sigs=False); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    print(f"chunk={chunk}: {best:.2f} ms = {n / best / 1e3:.2f} M sign/s", flush=True)
