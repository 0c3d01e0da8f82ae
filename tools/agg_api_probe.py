"""Development probe: where a single aggregate() call through the drop-in API spends its host time (secpar 128, 32 signers)."""
import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fusion_cryptography_b200.fusion.fusion import aggregate, fusion_setup, keygen, sign, verify
secpar = int(sys.argv[1]) if len(sys.argv) > 1 else 128
p = fusion_setup(secpar, 3)
keys = [keygen(p, 100 + 2 * i) for i in range(32)]
msgs = [f"m{i}" for i in range(32)]
sigs = [sign(p, k, m) for k, m in zip(keys, msgs)]
vks = [k[1] for k in keys]
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); agg = aggregate(p, vks, msgs, sigs); torch.cuda.synchronize(); print("aggregate", time.perf_counter() - t0)
pr = cProfile.Profile(); pr.enable(); agg = aggregate(p, vks, msgs, sigs); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
