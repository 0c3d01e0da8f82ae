"""The reference's own benchmark harness (benchmarks/benchmarks.py: fusion_setup, keygen, sign, aggregate, verify through the public API for
2, 4, ..., 32 signers at both security levels), without loguru / memory_profiler, for either implementation:

    python tools/api_benchmarks.py                         # the drop-in on the GPU (fusion_cryptography_b200.fusion.fusion)
    python tools/api_benchmarks.py --backend reference     # the unmodified reference (needs its checkout: --reference /root/reference)

Prints mean / median seconds per call, like the reference's summary, and one JSON line."""
import argparse
import json
import os
import random
import statistics
import string
import sys
import time

ap = argparse.ArgumentParser()
ap.add_argument("--backend", default="b200", choices=["b200", "reference"])
ap.add_argument("--reference", default=os.environ.get("FUSION_REFERENCE", "/root/reference"))
ap.add_argument("--max-size", type=int, default=32)
ap.add_argument("--secpars", default="128,256")
args = ap.parse_args()

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if args.backend == "reference":
    sys.dont_write_bytecode = True
    sys.path.insert(0, args.reference)
    from fusion.fusion import aggregate, fusion_setup, keygen, sign, verify
    sync = lambda: None  # noqa: E731
else:
    sys.path.insert(0, ROOT)
    import torch
    from fusion_cryptography_b200.fusion.fusion import aggregate, fusion_setup, keygen, sign, verify
    sync = torch.cuda.synchronize
    p0 = fusion_setup(128, 1)   # library load, context creation and kernel warm-up are not what the harness measures
    k0 = keygen(p0, 1)
    verify(p0, [k0[1]], ["w"], aggregate(p0, [k0[1]], ["w"], [sign(p0, k0, "w")]))


def timed(fn, *a):
    sync()
    t0 = time.perf_counter()
    out = fn(*a)
    sync()
    return out, time.perf_counter() - t0


rng = random.Random(20231018)
report = {}
for secpar in [int(x) for x in args.secpars.split(",")]:
    times = {k: [] for k in ("fusion_setup", "keygen", "sign", "aggregate", "verify")}
    n = 2
    while n <= args.max_size:
        params, t = timed(fusion_setup, secpar, rng.randint(0, 1000000))
        times["fusion_setup"].append(t)
        keys, msgs, sigs = [], [], []
        for _ in range(n):
            k, t = timed(keygen, params, rng.randint(0, 1000000))
            times["keygen"].append(t)
            keys.append(k)
            msgs.append("".join(rng.choices(string.ascii_letters + string.digits, k=20)))
        for k, m in zip(keys, msgs):
            s, t = timed(sign, params, k, m)
            times["sign"].append(t)
            sigs.append(s)
        vks = [k[1] for k in keys]
        agg, t = timed(aggregate, params, vks, msgs, sigs)
        times["aggregate"].append(t)
        ok, t = timed(verify, params, vks, msgs, agg)
        times["verify"].append(t)
        assert ok[0] is True, ok
        print(f"secpar {secpar} n={n}: aggregate {times['aggregate'][-1]:.4f} s, verify {times['verify'][-1]:.4f} s", file=sys.stderr)
        n *= 2
    report[secpar] = {k: {"mean": statistics.mean(v), "median": statistics.median(v), "calls": len(v)} for k, v in times.items()}
    report[secpar]["aggregate_by_size"] = times["aggregate"]
    report[secpar]["verify_by_size"] = times["verify"]
    print(f"Security Parameter: {secpar}  ({args.backend})")
    for k in ("fusion_setup", "keygen", "sign", "aggregate", "verify"):
        print(f"  {k:13s} mean {report[secpar][k]['mean']:.6f} s   median {report[secpar][k]['median']:.6f} s   ({report[secpar][k]['calls']} calls)")
print(json.dumps({"backend": args.backend, "sizes": "2..%d doubling" % args.max_size, "report": report}))
