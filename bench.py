#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200-native Fusion hot path.

Contract (see the task brief): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON
line on rank 0.  For N > 1 it is launched by torchrun (one rank per GPU, NCCL).

Workload (BASELINE.json configs[1]): 2^20 secpar-256 polynomials per GPU; one step =
forward NTT -> pointwise product with a resident NTT-domain operand -> inverse NTT
(algebra/ntt.py cooley_tukey_ntt / PolynomialNTTRepresentation.__mul__ / gentleman_sande_intt).
`value` = polynomials pushed through that pass per second, whole job, inputs resident in HBM.
`e2e` = the same pass through the C-ABI host-buffer entry point (pinned host memory in,
host memory out, copies inside the timed region).

`--impl reference` times the CPU restatement of the reference (oracle/, kind "port": the
reference is pure Python and cannot travel to the GPU box) on all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SECPAR = 256
D = 256
N_POLYS = 1 << 20
METRIC = "ntt_pass_polys_per_sec"
UNIT = "polys/s"
WORKLOAD = "configs[1]: fwd NTT + pointwise mul + inv NTT over 2^20 secpar-256 polynomials per GPU (one fused kernel: INTT(NTT(a) o b_hat))"
# algorithmic HBM bytes per polynomial: NTT/INTT 1 KiB in + 1 KiB out, pointwise 3 KiB
BYTES = {"ntt_pass": 3072, "ntt_fwd": 2048, "pointwise_mul": 3072, "ntt_inv": 2048}
# measured on B200 with tools/keccak_lab.cu: Keccak-f[1600] per second with one sponge per thread and the ALU pipe (LOP3/SHF) saturated
KECCAK_PEAK = 4.28e9


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons.  Started before warm-up and stopped after a
    short same-load tail so that even a very short timed region (K steps of ~1.5 ms) is covered;
    samples whose host timestamp falls inside the timed window are preferred."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None
        self.window = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def mark(self, t0: float, t1: float):
        self.window = (t0, t1)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        parsed = []
        for ts, r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                parsed.append((ts, float(f[0]), float(f[1]), float(f[2]), [n for n, v in zip(names, f[3:7]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        inside = [p for p in parsed if self.window and self.window[0] <= p[0] <= self.window[1] + 0.06]
        # "under load" = drawing clearly more than idle power while the same step loops
        loaded = [p for p in parsed if p[3] > 250.0]
        use, which = (inside, "timed region") if len(inside) >= 2 else (loaded, "same step looping around the timed region (region shorter than the sampling period)")
        if not use:
            use, which = parsed, "all samples"
        reasons = sorted({n for p in use for n in p[4]})
        return {"sm_mhz": statistics.median([p[1] for p in use]) if use else None, "sm_max_mhz": max([p[2] for p in use]) if use else None,
                "power_w_max": max([p[3] for p in use]) if use else None, "reasons": reasons, "samples": len(use), "window": which}


def cpu_pass(orc, a, bhat):
    """The same pass on the CPU oracle (algebra/ntt.py + polynomials.py restated in C)."""
    fa = orc.ntt_fwd(SECPAR, a)
    fc = orc.pointwise_mul(fa, bhat)
    return orc.ntt_inv(SECPAR, fc)


class numa_local:
    """Runs its body on the CPUs nearest to GPU `index` (NVML's ideal affinity), so that host memory pinned inside is placed on that GPU's NUMA node;
    the process's affinity is restored afterwards (the CPU legs use every core)."""

    def __init__(self, index: int):
        self.index, self.note, self.saved = index, {"note": "default placement"}, None

    def __enter__(self):
        try:
            import pynvml

            self.saved = os.sched_getaffinity(0)
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES is not remapped by NVML: translate through the device's UUID
            import torch

            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            pynvml.nvmlDeviceSetCpuAffinity(h)
            self.note["note"] = f"pinned on the GPU's NUMA node ({len(os.sched_getaffinity(0))} of {len(self.saved)} CPUs are local)"
        except Exception as e:  # no NVML / no permission: keep the default placement and say so
            self.note["note"] = f"default placement ({type(e).__name__})"
        return self.note

    def __exit__(self, *exc):
        if self.saved is not None:
            try:
                os.sched_setaffinity(0, self.saved)
            except Exception:
                pass
        return False


def cpu_baseline(target_seconds: float = 10.0):
    import numpy as np
    from oracle import oracle as orc

    threads = max(orc.max_threads(), len(os.sched_getaffinity(0)))
    orc.set_threads(threads)
    rng = np.random.default_rng(0xF0510)
    h = (2147465729 - 1) // 2
    n0 = 1 << 13
    a = rng.integers(-h, h + 1, size=(n0, D)).astype(np.int32)
    b = rng.integers(-h, h + 1, size=(n0, D)).astype(np.int32)
    cpu_pass(orc, a, b)
    t0 = time.perf_counter(); cpu_pass(orc, a, b); t1 = time.perf_counter()
    rate = n0 / (t1 - t0)
    n = int(min(1 << 20, max(n0, rate * target_seconds)))
    a = rng.integers(-h, h + 1, size=(n, D)).astype(np.int32)
    b = rng.integers(-h, h + 1, size=(n, D)).astype(np.int32)
    t0 = time.perf_counter(); cpu_pass(orc, a, b); t1 = time.perf_counter()
    return {"value": n / (t1 - t0), "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n} secpar-256 polynomials through fwd NTT + pointwise mul + inv NTT, C oracle (oracle/fz_oracle.c), {threads} OpenMP threads, {t1 - t0:.2f} s"}


def run_reference(args, rank, world):
    """Reference arm: the reference's CPU implementation of the pass (C port of the pure-Python
    reference, all host threads), each step a bounded sample of the workload."""
    if rank != 0:
        return
    import numpy as np
    from oracle import oracle as orc

    # torchrun exports OMP_NUM_THREADS=1 to its workers: take the cores this process may run on instead
    threads = max(orc.max_threads(), len(os.sched_getaffinity(0)))
    orc.set_threads(threads)
    rng = np.random.default_rng(0xF0510)
    h = (2147465729 - 1) // 2
    n = 1 << 15
    a = rng.integers(-h, h + 1, size=(n, D)).astype(np.int32)
    b = rng.integers(-h, h + 1, size=(n, D)).astype(np.int32)
    t0 = time.perf_counter(); cpu_pass(orc, a, b); dt = time.perf_counter() - t0
    # size a step so that warmup + steps stay within ~2 minutes
    budget = 100.0 / max(1, args.steps + args.warmup)
    n = int(min(1 << 20, max(1 << 12, n * budget / dt / 2)))
    a = rng.integers(-h, h + 1, size=(n, D)).astype(np.int32)
    b = rng.integers(-h, h + 1, size=(n, D)).astype(np.int32)
    for _ in range(args.warmup):
        cpu_pass(orc, a, b)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_pass(orc, a, b)
    dt = time.perf_counter() - t0
    v = n * args.steps / dt
    sample = f"{n} polynomials per step (bounded sample of the 2^20 workload), C oracle port, {threads} threads"
    emit({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
        "data": "synthetic", "config": {"workload": WORKLOAD, "sample_polys_per_step": n},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


def scheme_extras(args, ctx, rank, world, dist):
    """BASELINE metric, second half: sign/s and aggregate-verify sigs/s (configs[2]-[4]) at secpar 256, per-GPU shards
    (by signer / by aggregate, no data-path collective), whole-job numbers = sum over ranks / max time."""
    import numpy as np
    import torch

    from fusion_cryptography_b200.batch import Engine, KeyBatch

    eng = Engine(SECPAR, seed=42, device=ctx.device_index)
    dev = ctx.device
    ell, d, cap = eng.ell, eng.d, ctx.capacity
    nk = args.extra_keys
    seeds = (10**6 + 2 * (rank * nk + np.arange(nk))).astype(np.uint64)
    msgs = [f"msg-{rank * nk + i:07d}".encode() for i in range(nk)]

    def timed(fn, reps=1):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            out = fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / 1000.0 / reps, out

    def allmax(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    eng.keygen_sign(seeds[:256], msgs[:256])  # warm-up
    pm = ctx.pack_messages(msgs)              # messages resident in HBM before the timed region
    eng.keygen_sign(seeds, pm, chunk=65536)    # first full-size call pays cudaMalloc for the outputs; the timed one reuses them
    t_ks, (vk_l, vk_r, sigs) = timed(lambda: eng.keygen_sign(seeds, pm, chunk=65536))
    t_ks = allmax(t_ks)
    kb = eng.keygen(seeds)
    dg = ctx.hash_msg(pm)
    c_hat = ctx.hash_ch(kb.vk_l, kb.vk_r, dg)
    t_hash, _ = timed(lambda: ctx.hash_ch(kb.vk_l, kb.vk_r, dg), reps=3)
    t_mac, _ = timed(lambda: ctx.sign_mac_rows(kb.sk_hat, c_hat, out=sigs), reps=5)
    # general (ell distinct rows) sign MAC, the roofline kernel of SURVEY 8d: reads 2 * ell KiB, writes ell KiB per key
    nkg = min(nk, 4096)
    kg = eng.keygen(seeds[:nkg], want_sk=True)
    t_macg, _ = timed(lambda: ctx.sign_mac(kg.sk_l, kg.sk_r, c_hat[:nkg], out=sigs[:nkg]), reps=5)
    del kg
    # aggregates at capacity over sliding windows of the key pool
    na = args.extra_aggs
    counts = [cap] * na
    base = (np.arange(na)[:, None] * 97 + np.arange(cap)[None, :]) % nk
    sel = torch.from_numpy(base.reshape(-1).astype(np.int32)).to(dev)
    # key i always signs msgs[i]: its one signature (and prehash) is reused by every window that contains it
    dg_pos = ctx.hash_msg(msgs)[sel.long()].contiguous()
    # first call: buffers come from cudaMalloc; the timed call reuses them (a verifier service runs warm)
    eng.aggregate(kb.vk_l, kb.vk_r, None, sigs, counts, sel=sel, sig_sel=sel, digests=dg_pos)
    # these two legs are bound by the latency of the serial sponge chains, i.e. proportional to the SM clock: record it
    cs = ClockSampler(ctx.device_index)
    cs.start()
    w0 = time.time()
    t_agg, aggs = timed(lambda: eng.aggregate(kb.vk_l, kb.vk_r, None, sigs, counts, sel=sel, sig_sel=sel, digests=dg_pos))
    w1 = time.time()
    cs.mark(w0, w1)
    clk_agg = cs.stop()
    eng.verify(kb.vk_l, kb.vk_r, None, aggs, counts, sel=sel, digests=dg_pos)
    cs = ClockSampler(ctx.device_index)
    cs.start()
    w0 = time.time()
    t_ver, codes = timed(lambda: eng.verify(kb.vk_l, kb.vk_r, None, aggs, counts, sel=sel, digests=dg_pos))
    w1 = time.time()
    cs.mark(w0, w1)
    clk_ver = cs.stop()
    ok = bool((codes == 0).all().item())
    # 1 % tampered aggregates must be rejected with the reference's reason code
    bad = aggs.clone()
    bad[::100, 0, 0] += 1
    codes_bad = eng.verify(kb.vk_l, kb.vk_r, None, bad, counts, sel=sel, digests=dg_pos)
    tamper_ok = bool((codes_bad[::100] == 3).all().item()) and int((codes_bad != 0).sum().item()) == len(range(0, na, 100))
    del bad
    # C4: ONE capacity aggregate end to end (sort, 2818 challenges, the serial aggregation hash, the weighted sum) and its verification:
    # the latency of one SHAKE256 chain of 279 K permutations
    one = slice(0, cap)
    eng.aggregate(kb.vk_l, kb.vk_r, None, sigs, [cap], sel=sel[one], sig_sel=sel[one], digests=dg_pos[one])
    t_one_a, agg1 = timed(lambda: eng.aggregate(kb.vk_l, kb.vk_r, None, sigs, [cap], sel=sel[one], sig_sel=sel[one], digests=dg_pos[one]))
    eng.verify(kb.vk_l, kb.vk_r, None, agg1, [cap], sel=sel[one], digests=dg_pos[one])
    t_one_v, code1 = timed(lambda: eng.verify(kb.vk_l, kb.vk_r, None, agg1, [cap], sel=sel[one], digests=dg_pos[one]))
    single = {"signers": cap, "aggregate_seconds": t_one_a, "verify_seconds": t_one_v, "accepted": bool((code1 == 0).all().item()),
              "equals_batched_result": bool(torch.equal(agg1[0], aggs[0])),
              "note": "one serial SHAKE256 chain (197 K absorb + 82 K squeeze permutations) on one warp: latency-bound by construction"}
    # the whole C5 batch (10^4 capacity aggregates) on ONE GPU: with every sponge chain of the batch in flight the hash
    # kernels are bound by the integer ALU pipe instead of by the latency of a single chain
    full = None
    if world == 1 and args.full_batch_aggs > 0:
        nf = args.full_batch_aggs
        cf = [cap] * nf
        basef = (np.arange(nf)[:, None] * 97 + np.arange(cap)[None, :]) % nk
        self_ = torch.from_numpy(basef.reshape(-1).astype(np.int32)).to(dev)
        dgf = ctx.hash_msg(msgs)[self_.long()].contiguous()
        torch.cuda.empty_cache()
        aggf = eng.aggregate(kb.vk_l, kb.vk_r, None, sigs, cf, sel=self_, sig_sel=self_, digests=dgf)     # warm the allocator (tens of GB)
        t_fa, aggf = timed(lambda: eng.aggregate(kb.vk_l, kb.vk_r, None, sigs, cf, sel=self_, sig_sel=self_, digests=dgf))
        eng.verify(kb.vk_l, kb.vk_r, None, aggf, cf, sel=self_, digests=dgf)
        t_fv, codef = timed(lambda: eng.verify(kb.vk_l, kb.vk_r, None, aggf, cf, sel=self_, digests=dgf))
        full = {"aggregates": nf, "signers_per_aggregate": cap,
                "aggregate": {"value": nf * cap / t_fa, "unit": "sigs/s", "seconds": t_fa},
                "verify": {"value": nf * cap / t_fv, "unit": "sigs/s", "seconds": t_fv, "all_accept": bool((codef == 0).all().item())},
                "keccak_f_per_sec": nf * cap * 224.0 / t_fv,
                "note": "hash_ch 125 + hash_ag ~99 Keccak-f per signer (SURVEY 8d); measured ALU-pipe peak 4.28e9 Keccak-f/s (profiles/r01_keccak_lab_v2.txt)"}
        del aggf, codef, self_, dgf
    # the streaming aggregation sum alone (alpha precomputed): 86,016 B per signer
    perm, vl_s, vr_s, c_s, alpha = eng._sorted_hash(kb.vk_l, kb.vk_r, dg_pos, counts, sel)
    del vl_s, vr_s, c_s
    rows = sel[perm.long()].to(torch.int32).contiguous()
    t_sum, _ = timed(lambda: ctx.aggregate_sum(sigs, rows, alpha, counts), reps=5)
    t_sum1, _ = timed(lambda: ctx.aggregate_sum(sigs, rows[:cap].contiguous(), alpha[:cap].contiguous(), [cap]), reps=10)
    t_agg, t_ver = allmax(t_agg), allmax(t_ver)
    tot = na * cap
    return {
        "secpar": SECPAR, "n_gpus": world,
        "keygen_sign": {"keys_per_gpu": nk, "value": nk * world / t_ks, "unit": "sign/s", "seconds": t_ks,
                        "note": "seeds -> MT19937 sampler -> NTT -> A.sk -> SHA3/SHAKE challenge -> decode -> NTT -> sk_l*c+sk_r, sk never materialised as ell rows"},
        "aggregate": {"aggregates_per_gpu": na, "signers_per_aggregate": cap, "value": tot * world / t_agg, "unit": "sigs/s", "seconds": t_agg,
                      "clocks_rank0": clk_agg},
        "verify": {"aggregates_per_gpu": na, "signers_per_aggregate": cap, "value": tot * world / t_ver, "unit": "sigs/s", "seconds": t_ver, "all_accept": ok,
                   "tampered_1pct_rejected": tamper_ok, "clocks_rank0": clk_ver},
        "single_capacity_aggregate": single,
        "full_batch_one_gpu": full,
        "kernels": {
            "hash_ch": {"value": nk / t_hash, "unit": "challenges/s/GPU", "bound": "alu (LOP3/SHF)", "keccak_f_per_sec": 125.0 * nk / t_hash,
                        "alu_peak_keccak_f_per_sec": KECCAK_PEAK, "frac_of_alu_peak": 125.0 * nk / t_hash / KECCAK_PEAK,
                        "peak_source": "tools/keccak_lab.cu thread-per-sponge rate on B200 (profiles/r01_keccak_lab_v2.txt)"},
            "sign_mac_rows": {"GB/s": nk * ell * d * 4 / t_mac / 1e9, "ms": 1000 * t_mac, "bytes": "ell KiB written per key (keys as one row)"},
            "sign_mac": {"GB/s": nkg * 3 * ell * d * 4 / t_macg / 1e9, "ms": 1000 * t_macg, "bytes": "3 * ell KiB per key (SURVEY 8d: 257,024 B)"},
            "aggregate_sum": {"GB/s": tot * (ell + 1) * d * 4 / t_sum / 1e9, "ms": 1000 * t_sum, "bytes": "(ell + 1) KiB per signer (SURVEY 8d: 86,016 B)",
                              "single_capacity_aggregate_us": 1e6 * t_sum1, "single_GB/s": cap * (ell + 1) * d * 4 / t_sum1 / 1e9},
        },
    }


def verify_128(args, rank, world, dist, device_index):
    """BASELINE configs[4], secpar 128 half: capacity-sized (1796 signers) aggregates, aggregate + verify sigs/s per GPU share."""
    import numpy as np
    import torch

    from fusion_cryptography_b200 import native
    from fusion_cryptography_b200.batch import Engine

    ctx = native.get_context(128, device_index)
    eng = Engine(128, seed=42, device=device_index)
    dev, cap = ctx.device, ctx.capacity
    nk = min(args.extra_keys, 32768)
    seeds = (10**6 + 2 * (rank * nk + np.arange(nk))).astype(np.uint64)
    msgs = [f"msg-{rank * nk + i:07d}".encode() for i in range(nk)]
    vk_l, vk_r, sigs = eng.keygen_sign(seeds, ctx.pack_messages(msgs))
    na = args.extra_aggs
    counts = [cap] * na
    base = (np.arange(na)[:, None] * 97 + np.arange(cap)[None, :]) % nk
    sel = torch.from_numpy(base.reshape(-1).astype(np.int32)).to(dev)
    dg_pos = ctx.hash_msg(msgs)[sel.long()].contiguous()

    def timed(fn):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / 1000.0, out

    def allmax(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    eng.aggregate(vk_l, vk_r, None, sigs, counts, sel=sel, sig_sel=sel, digests=dg_pos)
    t_agg, aggs = timed(lambda: eng.aggregate(vk_l, vk_r, None, sigs, counts, sel=sel, sig_sel=sel, digests=dg_pos))
    eng.verify(vk_l, vk_r, None, aggs, counts, sel=sel, digests=dg_pos)
    t_ver, codes = timed(lambda: eng.verify(vk_l, vk_r, None, aggs, counts, sel=sel, digests=dg_pos))
    ok = bool((codes == 0).all().item())
    t_agg, t_ver = allmax(t_agg), allmax(t_ver)
    tot = na * cap
    # d = 64 transforms: 2^22 polynomials (1 GiB), 512 B of HBM traffic per polynomial and direction
    npoly = 1 << 22
    hq = (native.Q - 1) // 2
    x = torch.randint(-hq, hq + 1, (npoly, 64), dtype=torch.int32, device=dev)
    y = torch.empty_like(x)
    ntt = {}
    for name, fn in (("ntt_fwd", ctx.ntt_fwd), ("ntt_inv", ctx.ntt_inv)):
        for _ in range(3):
            fn(x, out=y)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            fn(x, out=y)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        ntt[name] = {"ms": ms, "GB/s": 2 * npoly * 256 / (ms / 1000.0) / 1e9, "polys": npoly}
    del x, y
    return {"secpar": 128, "aggregates_per_gpu": na, "signers_per_aggregate": cap, "ntt_d64": ntt,
            "aggregate": {"value": tot * world / t_agg, "unit": "sigs/s", "seconds": t_agg},
            "verify": {"value": tot * world / t_ver, "unit": "sigs/s", "seconds": t_ver, "all_accept": ok}}


def readme_demo():
    """BASELINE configs[0]: the README demo through the drop-in fusion.fusion API (single-item calls, Python objects in and out)."""
    import hashlib
    import random
    import string

    from fusion_cryptography_b200.fusion import fusion as ff

    def run():
        a = ff.fusion_setup(256, 42)
        keys = [ff.keygen(a, 42) for _ in range(2)]
        messages = ["".join(random.choices(string.ascii_letters + string.digits, k=20)) for _ in range(2)]
        sigs = [ff.sign(a, k, m) for k, m in zip(keys, messages)]
        vks = [k[1] for k in keys]
        agg = ff.aggregate(a, vks, messages, sigs)
        return messages, agg, ff.verify(a, vks, messages, agg)

    run()
    t0 = time.perf_counter()
    messages, agg, res = run()
    dt = time.perf_counter() - t0
    digest = hashlib.sha256(str(agg).encode()).hexdigest()
    return {"seconds": dt, "verify": list(res), "messages": messages, "agg_sha256": digest,
            "matches_reference_digest": digest == "130985dc44f2699f1fcffb8fc95a2fbe61682ea97b51ba1d03c2c83c1d735105",
            "reference_cpu_seconds": "about 5 s in CPython on one core (SURVEY section 6: keygen 1.2 s, sign 0.22 s, aggregate 0.16 s per signer, verify 0.7 s)"}


_JSON_OUT = None


def emit(line: dict) -> None:
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


# SASS mix of the loop body of ntt_pass_kernel<8> (one iteration = one warp = 2 polynomials; cuobjdump -sass of this build, profiles/r01_sass_mix_ntt_pass.txt):
# 356 IMAD-class instructions (2 issue cycles each on the FMA-heavy pipe of an SM sub-partition, measured with tools/ubench_int.cu) and 186 IMAD.HI / IMAD.WIDE
# (4 cycles each); 1,233 instructions in all.
PASS_FMAHEAVY_CYCLES_PER_ITER = 2 * 356 + 4 * 186
PASS_INSTRUCTIONS_PER_ITER = 1233


def int_pipe(ms, n_polys, clocks):
    """Fraction of the FMA-heavy (IMAD) pipe's peak the fused pass runs at: cycles the pipe needs for the launch (SASS count) / cycles the launch took
    at the SM clock sampled during the timed region; the ncu counter of the same kernel is quoted beside it."""
    import torch

    props = torch.cuda.get_device_properties(torch.cuda.current_device())
    mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0
    iters_per_smsp = n_polys / 2.0 / (props.multi_processor_count * 4)
    need = iters_per_smsp * PASS_FMAHEAVY_CYCLES_PER_ITER
    took = ms * 1e-3 * mhz * 1e6
    out = {"pipe": "fmaheavy (IMAD / IMAD.HI / IMAD.WIDE)", "cycles_needed_per_sub_partition": need, "cycles_elapsed": took, "frac": need / took,
           "issue_slot_frac": iters_per_smsp * PASS_INSTRUCTIONS_PER_ITER / took, "sm_mhz_used": mhz}
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        rec = json.load(open(tp)).get("ntt_pass_kernel<8>", {})
        for k in ("ncu_fmaheavy_pct", "ncu_alu_pct", "ncu_issue_active_pct", "ncu_dram_pct"):
            if k in rec:
                out[k] = rec[k]
    except Exception:
        pass
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--polys", type=int, default=N_POLYS, help="polynomials per GPU (default: the BASELINE size 2^20)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the keygen+sign / aggregate / verify throughput section")
    ap.add_argument("--extra-keys", type=int, default=262144, help="one-time keys per GPU for the sign leg")
    ap.add_argument("--extra-aggs", type=int, default=1250, help="capacity-sized aggregates per GPU for the aggregate/verify leg (10^4 / 8 GPUs)")
    ap.add_argument("--full-batch-aggs", type=int, default=10000, help="N = 1 only: also run the whole C5 batch of capacity aggregates on one GPU (0 = skip)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    # stdout carries exactly one JSON line: everything libraries print to file descriptor 1 (NCCL's version banner ...) goes to stderr
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from fusion_cryptography_b200 import native

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = native.get_context(SECPAR, local_rank)
    dev = ctx.device
    n = args.polys
    h = (native.Q - 1) // 2
    g = torch.Generator(device=dev)
    g.manual_seed(0xF0510 + rank)
    a = torch.randint(-h, h + 1, (n, D), dtype=torch.int32, device=dev, generator=g)
    bhat = torch.randint(-h, h + 1, (n, D), dtype=torch.int32, device=dev, generator=g)
    fa = torch.empty_like(a)
    fc = torch.empty_like(a)
    c = torch.empty_like(a)
    c3 = torch.empty_like(a)

    def step(events=None):
        # one pass of the hot path: INTT(NTT(a) o b_hat), fused in one kernel (bit-identical to the three separate kernels, checked below)
        if events is not None:
            events[0].record()
        ctx.ntt_pass(a, bhat, out=c)
        if events is not None:
            events[1].record()

    def three_kernels(events=None):
        if events is not None:
            events[0].record()
        ctx.ntt_fwd(a, out=fa)
        if events is not None:
            events[1].record()
        ctx.pointwise_mul(fa, bhat, out=fc)
        if events is not None:
            events[2].record()
        ctx.ntt_inv(fc, out=c3)
        if events is not None:
            events[3].record()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    t_pre = time.time()
    while time.time() - t_pre < 0.4:  # untimed: lets the clock sampler see this load before timing
        step()
        torch.cuda.synchronize()
    barrier()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_w0 = time.time()
    start.record()
    for i in range(args.steps):
        step(evs[i])
    end.record()
    barrier()
    t_w1 = time.time()
    elapsed_ms = start.elapsed_time(end)
    t_post = time.time()
    while time.time() - t_post < 0.3:  # untimed same-load tail for the sampler
        step()
        torch.cuda.synchronize()
    sampler.mark(t_w0, t_w1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    kern_ms = {"ntt_pass": sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps}
    # the same pass as three separate kernels, timed one by one outside the timed region (the fused kernel must give the same bits)
    for _ in range(3):
        three_kernels()
    ev3 = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(10)]
    for e in ev3:
        three_kernels(e)
    torch.cuda.synchronize()
    for i, k in enumerate(("ntt_fwd", "pointwise_mul", "ntt_inv")):
        kern_ms[k] = sum(e[i].elapsed_time(e[i + 1]) for e in ev3) / len(ev3)
    fused_ok = bool(torch.equal(c, c3))

    # end to end: pinned host in -> device pipeline -> host out, through the C-ABI host entry point
    with numa_local(local_rank) as numa_note:   # the pinned buffers of a rank live on its GPU's NUMA node (matters from 2 GPUs per box on)
        hin = torch.empty((n, D), dtype=torch.int32).pin_memory()
        hin.copy_(a.cpu())
        hout = torch.empty((n, D), dtype=torch.int32).pin_memory()
        hout.zero_()
    e2e_steps = max(3, min(args.steps, 10))
    ctx.pass_host(hin, bhat, hout)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.pass_host(hin, bhat, hout)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    same = bool(torch.equal(hout.to(dev), c))

    extra = scheme_extras(args, ctx, rank, world, dist if world > 1 else None) if not args.no_extras else None
    if extra is not None:
        torch.cuda.empty_cache()
        extra["secpar_128"] = verify_128(args, rank, world, dist if world > 1 else None, local_rank)
        if rank == 0:
            extra["readme_demo"] = readme_demo()
    if rank == 0:
        peak, peak_src = peaks()
        value = n * world * args.steps / (elapsed_ms / 1000.0)
        dom = "ntt_pass"
        achieved = BYTES[dom] * n / (kern_ms[dom] / 1000.0) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("ntt_pass_kernel<8>", {}).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "polys_per_gpu": n, "secpar": SECPAR, "degree": D,
                       "l2": "inputs (1 GiB per operand) far exceed the 126 MB L2; no flush needed",
                       "sharding": "by polynomial index, no data-path collective"},
            "roofline": {"bound": "hbm", "kernel": "ntt_pass_kernel<8>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": BYTES[dom] * n,
                         "note": "2,048 modular butterflies + 256 products per 3 KiB: the kernel is bound by the FMA-heavy (IMAD) pipe, not by HBM -- see int_pipe; "
                                 "the three-kernel form of the same pass moves 7 KiB per polynomial",
                         "int_pipe": int_pipe(kern_ms[dom], n, clocks),
                         "fused_equals_three_kernels": fused_ok,
                         "three_kernel_pass_ms": kern_ms["ntt_fwd"] + kern_ms["pointwise_mul"] + kern_ms["ntt_inv"],
                         "all_kernels": {k: {"ms": kern_ms[k], "GB/s": BYTES[k] * n / (kern_ms[k] / 1000.0) / 1e9,
                                             "frac": BYTES[k] * n / (kern_ms[k] / 1000.0) / 1e9 / peak} for k in kern_ms}},
            "e2e": {"value": n * world * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": n * D * 4, "d2h_bytes_per_step": n * D * 4,
                    "ms_per_step": 1000 * e2e_s / e2e_steps, "matches_device_path": same, "host_buffers": numa_note["note"]},
            "gpu_launches": args.steps,
            "clocks": clocks,
        }
        if extra is not None:
            for k in ("sign_mac", "aggregate_sum"):
                if k in extra.get("kernels", {}):
                    extra["kernels"][k]["frac_of_hbm_peak"] = extra["kernels"][k]["GB/s"] / peak
            line["extra"] = extra
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline()
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
