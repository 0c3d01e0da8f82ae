"""The reference's own benchmark harness (benchmarks/benchmarks.py: fusion_setup, keygen, sign, aggregate, verify through the public API for
2, 4, ..., 32 signers at both security levels), without loguru / memory_profiler, for either implementation:

    python tools/api_benchmarks.py                         # the drop-in on the GPU (fusion_cryptography_b200.fusion.fusion)
    python tools/api_benchmarks.py --backend reference     # the unmodified reference (baseline/_ref, else --reference /root/reference)
    python tools/api_benchmarks.py --backend both          # both on THIS host, one comparison table (the CPU column is the same box's CPU)
    python -m torch.distributed.run --nproc-per-node N ... tools/api_benchmarks.py    # N ranks, one GPU each: the harness is a serial stream of
                                                           # single calls, so N GPUs serve N independent callers (calls/s x N is printed)

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
ap.add_argument("--backend", default="b200", choices=["b200", "reference", "both"])
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_REF = os.path.join(_ROOT, "baseline", "_ref")
ap.add_argument("--reference", default=os.environ.get("FUSION_REFERENCE", _REF if os.path.isdir(os.path.join(_REF, "fusion")) else "/root/reference"))
ap.add_argument("--max-size", type=int, default=32)
ap.add_argument("--secpars", default="128,256")
args = ap.parse_args()

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RANK, WORLD, LOCAL = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
if args.backend == "both":
    import subprocess

    rows = {}
    for be in ("reference", "b200"):
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--backend", be, "--reference", args.reference, "--max-size", str(args.max_size),
                            "--secpars", args.secpars], capture_output=True, text=True, check=True)
        rows[be] = json.loads(r.stdout.strip().splitlines()[-1])["report"]
    print(f"# tools/api_benchmarks.py --backend both: the reference's benchmark harness (benchmarks/benchmarks.py: 2, 4, ..., {args.max_size} signers through the public API),")
    print(f"# seconds per call; b200 = the drop-in on one GPU of this box, reference = the unmodified Python reference ({args.reference}) on ONE core of the SAME box")
    print("secpar function        b200_mean   b200_median  reference_mean  reference_median  ratio_of_means")
    for sp in rows["b200"]:
        for fn in ("fusion_setup", "keygen", "sign", "aggregate", "verify"):
            b, rf = rows["b200"][sp][fn], rows["reference"][sp][fn]
            print(f"{sp:<6s} {fn:13s} {b['mean']:11.6f}  {b['median']:11.6f}  {rf['mean']:14.6f}  {rf['median']:16.6f}  {rf['mean'] / b['mean']:14.0f}")
        for what in ("aggregate", "verify"):
            print(f"{sp:<6s} {what:9s} by size (2,4,..): b200 {[round(x, 4) for x in rows['b200'][sp][what + '_by_size']]}  reference {[round(x, 3) for x in rows['reference'][sp][what + '_by_size']]}")
    print(json.dumps({"backend": "both", "report": rows}))
    sys.exit(0)
if args.backend == "reference":
    sys.dont_write_bytecode = True
    sys.path.insert(0, args.reference)
    from fusion.fusion import aggregate, fusion_setup, keygen, sign, verify
    sync = lambda: None  # noqa: E731
else:
    sys.path.insert(0, ROOT)
    import torch
    torch.cuda.set_device(LOCAL)
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
if WORLD > 1:   # N independent callers, one GPU each: gather every rank's means on rank 0
    import torch.distributed as dist

    dist.init_process_group("gloo")
    allr = [None] * WORLD
    dist.all_gather_object(allr, report)
    if RANK == 0:
        print(f"# {WORLD} ranks, one GPU each (independent callers): calls per second, all ranks together")
        for sp in report:
            for fn in ("fusion_setup", "keygen", "sign", "aggregate", "verify"):
                print(f"  secpar {sp} {fn:13s} {sum(1.0 / r[sp][fn]['mean'] for r in allr):12.1f} calls/s on {WORLD} GPUs   (rank 0 alone: {1.0 / report[sp][fn]['mean']:.1f})")
        print(json.dumps({"backend": args.backend, "n_gpus": WORLD, "sizes": "2..%d doubling" % args.max_size, "report_per_rank": allr}))
    dist.barrier()
    dist.destroy_process_group()
else:
    print(json.dumps({"backend": args.backend, "sizes": "2..%d doubling" % args.max_size, "report": report}))
