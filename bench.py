#!/usr/bin/env python3
"""bench.py -- benchmark of the B200-native Fusion hot path (BASELINE.json: "aggregate-verify sigs/sec & sign/sec at 1/2/4/8 B200;
NTT polys/sec vs roofline").

Contract (task brief): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line on rank 0; for N > 1 it is launched by torchrun
(one rank per GPU, NCCL).  The line's top level is BASELINE configs[1] -- the NTT pass -- and it carries the two other legs of the metric
as first-class, baselined numbers under `legs` (copied into `roofline.legs`, `e2e.legs`, `cpu_baseline.legs`, which the driver keeps):

  ntt     configs[1]  2^20 secpar-256 polynomials per GPU; one step = INTT(NTT(a) o b_hat) (algebra/ntt.py cooley_tukey_ntt,
                      PolynomialNTTRepresentation.__mul__, gentleman_sande_intt) in one fused kernel.  Weak scaling.
  sign    configs[2]  keygen + sign of 10^6 one-time keys / messages at secpar 256 (fusion.py:338-373, 534-557), sharded by signer index over
                      the ranks (fixed total: strong scaling), the reference's seeded keys (quirk Q1) AND the general ell-row path.
  verify  configs[3-4] aggregate + verify of 10^4 capacity-sized aggregates (2818 signers @256, 1796 @128; fusion.py:655-728), sharded by
                      aggregate index (fixed total: strong scaling); ONE capacity aggregate alone; and, for N > 1, that one aggregate with
                      its signers sharded over the ranks (the NCCL exchange paths parallel.aggregate_one_sharded / verify_one_sharded).

Every leg has: `value` (device-resident inputs, CUDA events, max over ranks), `e2e` (the same leg through HOST buffers, copies inside the
timed region), `roofline` (the dominant kernel against its bound) and `cpu_baseline` (N = 1: the UNMODIFIED Python reference from
baseline/_ref in a process pool over all host cores -- kind "reference" -- and its C restatement oracle/ on all host threads -- kind "port").

`--impl reference` times the reference's CPU implementation of the same legs on the box's host cores: the unmodified Python reference when
baseline/_ref is present (tools/install_reference.py puts it there; it travels with the snapshot), else the C port.
`--leg sign|verify` prints that leg as the top-level line instead (used for the per-leg records under profiles/).
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
REF_DIR = os.path.join(ROOT, "baseline", "_ref")

SECPAR = 256
D = 256
Q = 2147465729
N_POLYS = 1 << 20
SIGN_KEYS = 10**6
VERIFY_AGGS = 10**4
UNIT = {"ntt": "polys/s", "sign": "sign/s", "verify": "sigs/s"}
METRIC = {"ntt": "ntt_pass_polys_per_sec", "sign": "keygen_sign_per_sec", "verify": "aggregate_verify_sigs_per_sec"}
WORKLOAD = {
    "ntt": "configs[1]: fwd NTT + pointwise mul + inv NTT over 2^20 secpar-256 polynomials per GPU (one fused kernel: INTT(NTT(a) o b_hat))",
    "sign": "configs[2]: keygen + sign of 10^6 one-time keys / messages at secpar 256, seeds 10^6 + 2i, messages msg-%07d, sharded by signer index",
    "verify": "configs[4]: verify of 10^4 capacity-sized aggregates (2818 signers) at secpar 256, sharded by aggregate index; "
              "configs[3]: aggregate of the same batch and of ONE capacity aggregate",
}
# algorithmic HBM bytes per polynomial: NTT/INTT 1 KiB in + 1 KiB out, pointwise 3 KiB; the fused pass reads a and b_hat, writes the result
BYTES = {"ntt_pass": 3072, "ntt_fwd": 2048, "pointwise_mul": 3072, "ntt_inv": 2048}
# Keccak-f[1600] EXECUTED per signer.  hash_ch at secpar 256: 48 absorb (the text is 6.4 KB) + 61 squeeze -- the reference asks SHAKE256 for 10,436 bytes
# = 77 blocks (fusion.py:520-524, the figure of SURVEY 8d), but its decoder reads only the first 8,423 (fusion.py:436-438, 473: signums, omega magnitudes,
# d - 1 - omega index draws), so the kernel stops after the 62nd block (ctx.n_ch_used); at 128: 15 + 7 (1,075 of 1,551 bytes are read).  The roofline
# counts the permutations that run, not the ones the reference's digest() call computes and throws away.  hash_ag: about 70 absorb + 29.2 squeeze, all read.
PERMS_CH = {256: 109.0, 128: 22.0}
PERMS_CH_REFERENCE = {256: 125.0, 128: 27.0}
PERMS_AG = {256: 99.2, 128: 30.0}
# thread-per-sponge permutation: 179 ALU-pipe (LOP3 / SHF / IADD3) instructions per round once theta is folded into three-input LOP3s (cuobjdump of
# hash_ch_stream_kernel, tools/keccak_lab.cu V4); an SM sub-partition issues one ALU-pipe warp instruction every 2 cycles (tools/ubench_int.cu)
KECCAK_ALU_INSTR_PER_ROUND = 179
ALU_CYCLES_PER_WARP_INSTR = 2


def static_config(args, leg="ntt"):
    """The workload description: a function of the command line only, so that both arms print the same `config`."""
    cfg = {"workload": WORKLOAD[leg], "secpar": SECPAR, "degree": D}
    if leg == "ntt":
        cfg.update({"polys_per_gpu": args.polys, "l2": "inputs (1 GiB per operand) far exceed the 126 MB L2; no flush needed",
                    "sharding": "by polynomial index, no data-path collective"})
        cfg["legs"] = {"sign": static_config(args, "sign"), "verify": static_config(args, "verify")}
    elif leg == "sign":
        cfg.update({"keys_total": args.sign_keys, "sharding": "by signer index, no data-path collective", "scaling": "strong (fixed total)",
                    "general_row_keys_per_gpu": args.general_keys})
    else:
        cfg.update({"aggregates_total": args.verify_aggs, "signers_per_aggregate": {"256": 2818, "128": 1796},
                    "sharding": "by aggregate index, no data-path collective; one-aggregate paths: all-gather of challenges + all-gather/fold of partials (NCCL)",
                    "scaling": "strong (fixed total)", "key_pool_per_gpu": args.pool_keys})
    return cfg


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_record(kernel):
    """Counters of one kernel from the committed ncu captures (profiles/ncu_traffic.json), or {}."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(kernel, {})
    except Exception:
        return {}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons.  Started before warm-up and stopped after a short same-load tail so that even a very short
    timed region is covered; samples whose host timestamp falls inside the timed window are preferred."""

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
        return self

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
        loaded = [p for p in parsed if p[3] > 250.0]   # "under load" = drawing clearly more than idle power while the same step loops
        use, which = (inside, "timed region") if len(inside) >= 2 else (loaded, "same step looping around the timed region (region shorter than the sampling period)")
        if not use:
            use, which = parsed, "all samples"
        reasons = sorted({n for p in use for n in p[4]})
        return {"sm_mhz": statistics.median([p[1] for p in use]) if use else None, "sm_max_mhz": max([p[2] for p in use]) if use else None,
                "power_w_max": max([p[3] for p in use]) if use else None, "reasons": reasons, "samples": len(use), "window": which}


class numa_local:
    """Runs its body on the CPUs nearest to GPU `index` (NVML's ideal affinity), so that host memory pinned inside is placed on that GPU's NUMA node;
    the process's affinity is restored afterwards (the CPU legs use every core)."""

    def __init__(self, index: int):
        self.index, self.note, self.saved = index, {"note": "default placement"}, None

    def __enter__(self):
        try:
            import pynvml
            import torch

            self.saved = os.sched_getaffinity(0)
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)   # CUDA_VISIBLE_DEVICES is not remapped by NVML: go through the UUID
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


_JSON_OUT = None


def emit(line: dict) -> None:
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


# =====================================================================================================================================
# CPU side: the unmodified Python reference (baseline/_ref) in a process pool, and its C restatement (oracle/) on all host threads.
# Workers are module-level functions (spawned processes import this file; nothing below touches CUDA).
# =====================================================================================================================================
def host_cores() -> int:
    return len(os.sched_getaffinity(0))


def reference_present() -> bool:
    return os.path.isdir(os.path.join(REF_DIR, "fusion")) and os.path.isdir(os.path.join(REF_DIR, "algebra"))


def _ref_path():
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)


_W = {}


def _w_ntt_tables():
    """Per-worker state of the NTT leg: the reference's own bit-reversed power tables (built once, as BASELINE.md 3 prescribes) and inputs."""
    if "ntt" not in _W:
        _ref_path()
        import random

        from algebra.ntt import bit_reverse_copy
        from fusion.fusion import fusion_setup

        p = fusion_setup(256, 42)
        q, d = p.modulus, p.degree
        tw = bit_reverse_copy([pow(p.root, i, q) for i in range(d)])
        itw = bit_reverse_copy([pow(p.inv_root, i, q) for i in range(d)])
        rng = random.Random(0xF0510 + os.getpid())
        h = q // 2
        a = [[rng.randrange(-h, h + 1) for _ in range(d)] for _ in range(8)]
        b = [[rng.randrange(-h, h + 1) for _ in range(d)] for _ in range(8)]
        _W["ntt"] = (p, q, d, tw, itw, a, b)
    return _W["ntt"]


def _w_ntt_raw(n: int) -> float:
    """n passes INTT(NTT(a) o b_hat) with the reference's cooley_tukey_ntt / cent / gentleman_sande_intt and prebuilt tables."""
    p, q, d, tw, itw, a, b = _w_ntt_tables()
    from algebra.ntt import cent, cooley_tukey_ntt, gentleman_sande_intt

    t0 = time.perf_counter()
    for i in range(n):
        v = list(a[i & 7])
        cooley_tukey_ntt(val=v, modulus=q, root_order=p.root_order, bit_rev_root_powers=tw)
        bh = b[i & 7]
        v = [cent(val=x * y, modulus=q, halfmod=q // 2, logmod=q.bit_length() - 1) for x, y in zip(v, bh)]
        gentleman_sande_intt(val=v, modulus=q, root_order=p.root_order, bit_rev_inv_root_powers=itw)
    return time.perf_counter() - t0


def _w_ntt_api(n: int) -> float:
    """The same pass as a user of the reference writes it: transform(transform(a) * b_hat) on the polynomial classes (tables rebuilt per call)."""
    p, q, d, tw, itw, a, b = _w_ntt_tables()
    from algebra.polynomials import PolynomialCoefficientRepresentation as PC, PolynomialNTTRepresentation as PN, transform

    kw = dict(modulus=q, degree=d, root=p.root, inv_root=p.inv_root, root_order=p.root_order)
    pa = [PC(coefficients=list(x), **kw) for x in a]
    pb = [PN(values=list(x), **kw) for x in b]
    t0 = time.perf_counter()
    for i in range(n):
        transform(transform(pa[i & 7]) * pb[i & 7])
    return time.perf_counter() - t0


def _w_scheme(job):
    """One worker's share of the scheme legs with the unmodified reference API: n_keys keygen, n sign (the signers take the keys in turn, each with its
    own message -- the one-time keys are reused only to bound the keygen time of this sample, as the GPU leg's key pool is), ONE aggregate over
    them and its verify."""
    secpar, n_keys, n, first = job
    _ref_path()
    from fusion.fusion import aggregate, fusion_setup, keygen, sign, verify

    a = fusion_setup(secpar, 42)
    t0 = time.perf_counter()
    keys = [keygen(a, 10**6 + 2 * (first + i)) for i in range(n_keys)]
    t1 = time.perf_counter()
    msgs = [f"msg-{first + i:07d}" for i in range(n)]
    signers = [keys[i % n_keys] for i in range(n)]
    sigs = [sign(a, k, m) for k, m in zip(signers, msgs)]
    t2 = time.perf_counter()
    vks = [k[1] for k in signers]
    agg = aggregate(a, vks, msgs, sigs)
    t3 = time.perf_counter()
    ok = verify(a, vks, msgs, agg)
    t4 = time.perf_counter()
    return {"n": n, "n_keys": n_keys, "keygen_s": t1 - t0, "sign_s": t2 - t1, "aggregate_s": t3 - t2, "verify_s": t4 - t3, "ok": bool(ok[0])}


class RefPool:
    """multiprocessing pool over all host cores running the unmodified reference (spawned: the parent may hold a CUDA context)."""

    def __init__(self):
        import multiprocessing as mp

        self.cores = host_cores()
        self.pool = mp.get_context("spawn").Pool(self.cores)
        self.pool.map(_w_ntt_raw, [1] * self.cores)   # start the workers, build their tables

    def close(self):
        self.pool.close()
        self.pool.join()

    def ntt_step(self, per_worker: int) -> float:
        """One step of the NTT leg: every core does `per_worker` passes.  Returns wall seconds."""
        t0 = time.perf_counter()
        self.pool.map(_w_ntt_raw, [per_worker] * self.cores)
        return time.perf_counter() - t0

    def ntt_api(self, per_worker: int) -> float:
        t0 = time.perf_counter()
        self.pool.map(_w_ntt_api, [per_worker] * self.cores)
        return time.perf_counter() - t0

    def scheme(self, secpar: int, n_small: int, n_large: int, n_keys: int = 2):
        """Scheme legs: every core runs keygen of n_keys keys, sign of n signers and one aggregate + verify over them; two aggregate sizes give the fixed
        and the per-signer cost of aggregate / verify (t = t0 + t1 * N), from which the capacity-sized figure is EXTRAPOLATED (and labelled so)."""
        out = {}
        for tag, n in (("small", n_small), ("large", n_large)):
            t0 = time.perf_counter()
            res = self.pool.map(_w_scheme, [(secpar, n_keys, n, w * n) for w in range(self.cores)])
            wall = time.perf_counter() - t0
            assert all(r["ok"] for r in res), "the reference rejected its own aggregate"
            m = lambda k: statistics.mean(r[k] for r in res)
            out[tag] = {"n": n, "n_keys": n_keys, "wall_s": wall, "keygen_s": m("keygen_s"), "sign_s": m("sign_s"), "aggregate_s": m("aggregate_s"), "verify_s": m("verify_s")}
        return out


def reference_scheme_legs(pool: RefPool, secpar: int = 256, n_small: int = 8, n_large: int = 32):
    """cpu_baseline entries (kind "reference") of the sign and verify legs."""
    cap = {256: 2818, 128: 1796}[secpar]
    r = pool.scheme(secpar, n_small, n_large)
    s, l = r["small"], r["large"]
    cores = pool.cores
    t_keygen = (s["keygen_s"] + l["keygen_s"]) / (s["n_keys"] + l["n_keys"])     # seconds per keygen, one core
    t_sign = (s["sign_s"] + l["sign_s"]) / (s["n"] + l["n"])                     # seconds per sign, one core
    fit = lambda k: ((l[k] - s[k]) / (l["n"] - s["n"]), s[k] - (l[k] - s[k]) / (l["n"] - s["n"]) * s["n"])   # (per signer, fixed)
    ag1, ag0 = fit("aggregate_s")
    vf1, vf0 = fit("verify_s")
    ag0, vf0 = max(ag0, 0.0), max(vf0, 0.0)
    sample = (f"unmodified Python reference (baseline/_ref) in a pool of {cores} processes, secpar {secpar}: per process {s['n_keys'] + l['n_keys']} keygen, {s['n']} + {l['n']} sign, "
              f"and one aggregate + verify of {s['n']} and of {l['n']} signers ({s['wall_s'] + l['wall_s']:.1f} s wall)")
    sign = {"value": cores / (t_keygen + t_sign), "unit": UNIT["sign"], "cores": cores, "kind": "reference", "sample": sample,
            "seconds_per_keygen_one_core": t_keygen, "seconds_per_sign_one_core": t_sign}
    t_ver_cap = vf0 + vf1 * cap
    t_agg_cap = ag0 + ag1 * cap
    verify = {"value": cores * l["n"] / l["verify_s"], "unit": UNIT["verify"], "cores": cores, "kind": "reference", "sample": sample,
              "measured": {"signers": l["n"], "verify_seconds_one_core": l["verify_s"], "aggregate_seconds_one_core": l["aggregate_s"]},
              "fit": {"verify_fixed_s": vf0, "verify_per_signer_s": vf1, "aggregate_fixed_s": ag0, "aggregate_per_signer_s": ag1},
              "extrapolated_to_capacity": {"signers": cap, "verify_sigs_per_sec_all_cores": cores * cap / t_ver_cap, "aggregate_sigs_per_sec_all_cores": cores * cap / t_agg_cap,
                                           "note": "linear extrapolation t = fixed + per_signer * N of the two measured sizes; NOT run (one capacity aggregate is about 10 minutes of CPython)"},
              "aggregate_value": cores * l["n"] / l["aggregate_s"]}
    return sign, verify


def port_threads() -> int:
    from oracle import oracle as orc

    n = max(orc.max_threads(), host_cores())   # torchrun exports OMP_NUM_THREADS=1 to its workers: take the cores this process may run on
    orc.set_threads(n)
    return n


def port_ntt_pass(orc, a, bhat):
    return orc.ntt_inv(SECPAR, orc.pointwise_mul(orc.ntt_fwd(SECPAR, a), bhat))


def port_ntt(target_seconds: float = 6.0):
    """cpu_baseline (kind "port") of the NTT leg: the C restatement of the same pass on all host threads."""
    import numpy as np
    from oracle import oracle as orc

    threads = port_threads()
    rng = np.random.default_rng(0xF0510)
    h = (Q - 1) // 2
    n0 = 1 << 13
    a = rng.integers(-h, h + 1, size=(n0, D)).astype(np.int32)
    b = rng.integers(-h, h + 1, size=(n0, D)).astype(np.int32)
    port_ntt_pass(orc, a, b)
    t0 = time.perf_counter(); port_ntt_pass(orc, a, b); t1 = time.perf_counter()
    n = int(min(1 << 20, max(n0, n0 / (t1 - t0) * target_seconds)))
    a = rng.integers(-h, h + 1, size=(n, D)).astype(np.int32)
    b = rng.integers(-h, h + 1, size=(n, D)).astype(np.int32)
    t0 = time.perf_counter(); port_ntt_pass(orc, a, b); t1 = time.perf_counter()
    return {"value": n / (t1 - t0), "unit": UNIT["ntt"], "cores": threads, "kind": "port",
            "sample": f"{n} secpar-256 polynomials through fwd NTT + pointwise mul + inv NTT, C oracle (oracle/fz_oracle.c), {threads} OpenMP threads, {t1 - t0:.2f} s"}


def port_scheme_legs(secpar: int = 256, n_keys: int = 0, n_aggs: int = 0, signers: int = 64):
    """cpu_baseline entries (kind "port") of the sign and verify legs: fzo_keygen_batch / fzo_sign_batch over all host threads; aggregates of `signers`
    signers verified one per thread (the port's aggregate / verify are serial inside, like the reference's)."""
    from concurrent.futures import ThreadPoolExecutor

    import numpy as np
    from oracle import oracle as orc

    threads = port_threads()
    n_keys = n_keys or 16 * threads
    n_aggs = n_aggs or threads
    A = orc.setup_A(secpar, 42)
    seeds = (10**6 + 2 * np.arange(n_keys)).astype(np.uint64)
    msgs = [f"msg-{i:07d}".encode() for i in range(n_keys)]
    t0 = time.perf_counter()
    sk_l, sk_r, vl, vr = orc.keygen_batch(secpar, A, seeds)
    sigs = orc.sign_batch(secpar, sk_l, sk_r, vl, vr, msgs)
    t1 = time.perf_counter()
    sign = {"value": n_keys / (t1 - t0), "unit": UNIT["sign"], "cores": threads, "kind": "port",
            "sample": f"{n_keys} keygen + sign at secpar {secpar}, C oracle, {threads} OpenMP threads, {t1 - t0:.2f} s"}
    orc.set_threads(1)
    signers = min(signers, n_keys)

    def one(j):
        ks = (np.arange(signers) + 7 * j) % n_keys
        m = [msgs[i] for i in ks]
        t0 = time.perf_counter()
        agg = orc.aggregate(secpar, vl[ks], vr[ks], m, sigs[ks])
        t1 = time.perf_counter()
        code = orc.verify(secpar, A, vl[ks], vr[ks], m, agg)
        return t1 - t0, time.perf_counter() - t1, code

    w0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        res = list(ex.map(one, range(n_aggs)))
    wall = time.perf_counter() - w0
    orc.set_threads(threads)
    assert all(r[2] == 0 for r in res)
    t_agg, t_ver = sum(r[0] for r in res), sum(r[1] for r in res)
    verify = {"value": n_aggs * signers / wall * (t_agg + t_ver) / t_ver, "unit": UNIT["verify"], "cores": threads, "kind": "port",
              "aggregate_value": n_aggs * signers / wall * (t_agg + t_ver) / t_agg,
              "sample": f"{n_aggs} aggregates of {signers} signers at secpar {secpar}, one per thread over {threads} threads (aggregate then verify, {wall:.2f} s wall; "
                        "each rate = signatures / that stage's share of the wall time)"}
    return sign, verify


def run_reference(args, rank, world):
    """Reference arm: the reference's own CPU implementation of the legs on the box's host cores (rank 0 only; other ranks exit)."""
    if rank != 0:
        return
    leg = args.leg
    use_py = reference_present() and not args.port_only
    cfg = static_config(args, leg)
    legs_cpu = {}
    if use_py:
        pool = RefPool()
        cores = pool.cores
        kind = "reference"
        # NTT leg: steps of `per` passes per core, sized so that warm-up + steps stay around a minute
        t = pool.ntt_step(4)
        per = max(2, int(4 * (40.0 / max(1, args.steps + args.warmup)) / t))
        per = min(per, 256)
        for _ in range(args.warmup):
            pool.ntt_step(per)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.ntt_step(per)
        dt = time.perf_counter() - t0
        n_step = per * cores
        ntt_v = n_step * args.steps / dt
        api_n = max(1, per // 8)
        api_v = api_n * cores / pool.ntt_api(api_n)
        ntt = {"value": ntt_v, "unit": UNIT["ntt"], "cores": cores, "kind": kind,
               "sample": f"{n_step} polynomials per step ({per} per process, {cores} processes): the reference's cooley_tukey_ntt, cent and gentleman_sande_intt "
                         f"(algebra/ntt.py) with prebuilt tables; as a user writes it -- transform(transform(a) * b_hat) on the polynomial classes -- {api_v:.1f} polys/s",
               "as_used_class_api_value": api_v}
        sgn, ver = reference_scheme_legs(pool)
        pool.close()
        p_sgn, p_ver = port_scheme_legs()
        ntt["port"] = port_ntt()
        sgn["port"], ver["port"] = p_sgn, p_ver
        ms_step = 1000 * dt / args.steps
    else:
        import numpy as np
        from oracle import oracle as orc

        threads = port_threads()
        kind = "port"
        rng = np.random.default_rng(0xF0510)
        h = (Q - 1) // 2
        n = 1 << 15
        a = rng.integers(-h, h + 1, size=(n, D)).astype(np.int32)
        b = rng.integers(-h, h + 1, size=(n, D)).astype(np.int32)
        t0 = time.perf_counter(); port_ntt_pass(orc, a, b); dt = time.perf_counter() - t0
        n = int(min(1 << 20, max(1 << 12, n * (100.0 / max(1, args.steps + args.warmup)) / dt / 2)))
        a = rng.integers(-h, h + 1, size=(n, D)).astype(np.int32)
        b = rng.integers(-h, h + 1, size=(n, D)).astype(np.int32)
        for _ in range(args.warmup):
            port_ntt_pass(orc, a, b)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            port_ntt_pass(orc, a, b)
        dt = time.perf_counter() - t0
        ntt = {"value": n * args.steps / dt, "unit": UNIT["ntt"], "cores": threads, "kind": kind,
               "sample": f"{n} polynomials per step (bounded sample of the 2^20 workload), C oracle port, {threads} threads (baseline/_ref absent)"}
        sgn, ver = port_scheme_legs()
        ms_step = 1000 * dt / args.steps
    legs_cpu = {"ntt": ntt, "sign": sgn, "verify": ver}
    top = legs_cpu[leg]
    line = {
        "impl": "reference", "metric": METRIC[leg], "value": top["value"], "unit": UNIT[leg], "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak" if leg == "ntt" else "strong", "vs_baseline": None,
        "dtype": "int (CPython arbitrary precision)" if kind == "reference" else "int64", "data": "synthetic", "config": cfg,
        "cpu_baseline": dict(top, legs={k: v for k, v in legs_cpu.items() if k != leg}),
        "e2e": {"value": top["value"], "unit": UNIT[leg], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "legs": {k: {"value": v["value"], "unit": v["unit"]} for k, v in legs_cpu.items() if k != leg}},
        "gpu_launches": 0,
        "legs": {k: {"value": v["value"], "unit": v["unit"], "kind": v["kind"], "cores": v["cores"]} for k, v in legs_cpu.items()},
    }
    emit(line)


# =====================================================================================================================================
# GPU side
# =====================================================================================================================================
class Dev:
    """What every GPU leg needs: rank layout, timers and reductions."""

    def __init__(self, args, rank, world, local_rank, dist):
        import torch

        self.args, self.rank, self.world, self.local, self.dist, self.torch = args, rank, world, local_rank, dist, torch
        self.device = torch.device("cuda", local_rank)
        self.props = torch.cuda.get_device_properties(local_rank)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, reps=1):
        torch = self.torch
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            out = fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / 1000.0 / reps, out

    def allmax(self, x: float) -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def allmin_int(self, x: int) -> int:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.int64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return int(t.item())

    def alu_peak_perms(self, mhz) -> float:
        """Keccak-f[1600] per second at which the integer ALU pipe (LOP3 / SHF / IADD3) of the whole GPU is saturated at `mhz`."""
        cyc_per_perm_per_thread = 24 * KECCAK_ALU_INSTR_PER_ROUND * ALU_CYCLES_PER_WARP_INSTR / 32.0
        return self.props.multi_processor_count * 4 * mhz * 1e6 / cyc_per_perm_per_thread


# SASS mix of the loop body of ntt_pass_kernel<8> (one iteration = one warp = 2 polynomials; cuobjdump -sass of this build, profiles/r01_sass_mix_ntt_pass.txt):
# 356 IMAD-class instructions (2 issue cycles each on the FMA-heavy pipe of an SM sub-partition, measured with tools/ubench_int.cu) and 186 IMAD.HI / IMAD.WIDE
# (4 cycles each); 1,233 instructions in all.
PASS_FMAHEAVY_CYCLES_PER_ITER = 2 * 356 + 4 * 186
PASS_INSTRUCTIONS_PER_ITER = 1233


def leg_ntt(dv: Dev, ctx):
    """configs[1]: K timed steps of the fused pass; e2e through fz_ntt_pass_host; the three-kernel form beside it."""
    import torch

    from fusion_cryptography_b200 import native

    args, dev, rank, world, dist = dv.args, dv.device, dv.rank, dv.world, dv.dist
    n = args.polys
    h = (native.Q - 1) // 2
    g = torch.Generator(device=dev)
    g.manual_seed(0xF0510 + rank)
    a = torch.randint(-h, h + 1, (n, D), dtype=torch.int32, device=dev, generator=g)
    bhat = torch.randint(-h, h + 1, (n, D), dtype=torch.int32, device=dev, generator=g)
    fa, fc, c, c3 = (torch.empty_like(a) for _ in range(4))

    def step(events=None):
        if events is not None:
            events[0].record()
        ctx.ntt_pass(a, bhat, out=c)
        if events is not None:
            events[1].record()

    def three_kernels(events=None):
        fns = ((ctx.ntt_fwd, (a,), fa), (ctx.pointwise_mul, (fa, bhat), fc), (ctx.ntt_inv, (fc,), c3))
        for i, (fn, ins, out) in enumerate(fns):
            if events is not None:
                events[i].record()
            fn(*ins, out=out)
        if events is not None:
            events[3].record()

    sampler = ClockSampler(dv.local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    t_pre = time.time()
    while time.time() - t_pre < 0.4:  # untimed: lets the clock sampler see this load before timing
        step()
        torch.cuda.synchronize()
    dv.barrier()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(args.steps)]
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dv.barrier()
    t_w0 = time.time()
    start.record()
    for i in range(args.steps):
        step(evs[i])
    end.record()
    dv.barrier()
    t_w1 = time.time()
    elapsed_ms = dv.allmax(start.elapsed_time(end))
    t_post = time.time()
    while time.time() - t_post < 0.3:  # untimed same-load tail for the sampler
        step()
        torch.cuda.synchronize()
    sampler.mark(t_w0, t_w1)
    clocks = sampler.stop() if rank == 0 else None
    kern_ms = {"ntt_pass": sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps}
    for _ in range(3):
        three_kernels()
    ev3 = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(10)]
    for e in ev3:
        three_kernels(e)
    torch.cuda.synchronize()
    for i, k in enumerate(("ntt_fwd", "pointwise_mul", "ntt_inv")):
        kern_ms[k] = sum(e[i].elapsed_time(e[i + 1]) for e in ev3) / len(ev3)
    fused_ok = bool(torch.equal(c, c3))
    del fa, fc, c3

    # end to end: pinned host in -> device pipeline -> host out, through the C-ABI host entry point
    with numa_local(dv.local) as numa_note:   # the pinned buffers of a rank live on its GPU's NUMA node
        hin = torch.empty((n, D), dtype=torch.int32, pin_memory=True)
        hin.copy_(a)
        hout = torch.empty((n, D), dtype=torch.int32, pin_memory=True)
        hout.zero_()
    e2e_steps = max(3, min(args.steps, 10))
    # the ceiling of that leg on this box: the bare pinned copies (1 GiB up and 1 GiB down, concurrently, nothing computed), every rank at once
    s_up, s_dn = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def bare_copies():
        with torch.cuda.stream(s_up):
            a.copy_(hin, non_blocking=True)
        with torch.cuda.stream(s_dn):
            hout.copy_(c, non_blocking=True)

    bare_copies()
    torch.cuda.synchronize()
    dv.barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        bare_copies()
    torch.cuda.synchronize()
    bare_s = dv.allmax(time.perf_counter() - t0) / 3
    ctx.pass_host(hin, bhat, hout)
    dv.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.pass_host(hin, bhat, hout)
    torch.cuda.synchronize()
    e2e_s = dv.allmax(time.perf_counter() - t0)
    same = bool(torch.equal(hout.to(dev), c))
    del hin, hout, a, bhat, c
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    peak, peak_src = peaks()
    ms = kern_ms["ntt_pass"]
    mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0
    iters_per_smsp = n / 2.0 / (dv.props.multi_processor_count * 4)
    need, took = iters_per_smsp * PASS_FMAHEAVY_CYCLES_PER_ITER, ms * 1e-3 * mhz * 1e6
    rec = ncu_record("ntt_pass_kernel<8>")
    hbm = BYTES["ntt_pass"] * n / (ms / 1000.0) / 1e9
    roofline = {
        "bound": "int32 imad pipe", "kernel": "ntt_pass_kernel<8>",
        "achieved": need / took * 100.0, "peak": 100.0, "unit": "% of FMA-heavy (IMAD / IMAD.HI / IMAD.WIDE) pipe issue cycles", "frac": need / took,
        "how": "pipe cycles the launch needs (SASS count of the loop body: 356 two-cycle + 186 four-cycle instructions per warp iteration of 2 polynomials) / "
               "cycles the launch took at the SM clock sampled in the timed region; ncu's sm__pipe_fmaheavy_cycles_active of the same kernel is quoted beside it",
        "ncu_fmaheavy_pct": rec.get("ncu_fmaheavy_pct"), "ncu_alu_pct": rec.get("ncu_alu_pct"), "ncu_issue_active_pct": rec.get("ncu_issue_active_pct"),
        "ncu_dram_pct": rec.get("ncu_dram_pct"), "issue_slot_frac": iters_per_smsp * PASS_INSTRUCTIONS_PER_ITER / took, "sm_mhz_used": mhz,
        "traffic": rec.get("dram_bytes_per_launch"), "algorithmic_bytes_per_launch": BYTES["ntt_pass"] * n,
        "hbm": {"achieved": hbm, "peak": peak, "unit": "GB/s", "frac": hbm / peak, "peak_source": peak_src,
                "note": "second limiter: 3,072 algorithmic bytes per polynomial; DRAM traffic from ncu equals the algorithmic bytes (no re-reads)"},
        "fused_equals_three_kernels": fused_ok,
        "three_kernel_pass_ms": kern_ms["ntt_fwd"] + kern_ms["pointwise_mul"] + kern_ms["ntt_inv"],
        "all_kernels": {k: {"ms": kern_ms[k], "GB/s": BYTES[k] * n / (kern_ms[k] / 1000.0) / 1e9, "frac_of_hbm_peak": BYTES[k] * n / (kern_ms[k] / 1000.0) / 1e9 / peak}
                        for k in kern_ms},
    }
    return {
        "metric": METRIC["ntt"], "value": n * world * args.steps / (elapsed_ms / 1000.0), "unit": UNIT["ntt"], "ms_per_step": elapsed_ms / args.steps,
        "scaling": "weak", "dtype": "u32", "gpu_launches": args.steps, "clocks": clocks, "roofline": roofline,
        "e2e": {"value": n * world * e2e_steps / e2e_s, "unit": UNIT["ntt"], "h2d_bytes_per_step": n * D * 4, "d2h_bytes_per_step": n * D * 4,
                "ms_per_step": 1000 * e2e_s / e2e_steps, "matches_device_path": same, "host_buffers": numa_note["note"],
                "pcie_ceiling": {"ms_per_step": 1000 * bare_s, "GB/s_per_direction_all_ranks": world * n * D * 4 / bare_s / 1e9,
                                 "frac": bare_s / (e2e_s / e2e_steps),
                                 "note": "the same pinned buffers copied up and down concurrently with nothing computed, all ranks at once (max over ranks): what the box's "
                                         "host links give this leg; frac = that time / the pipeline's time"}},
    }


def leg_sign(dv: Dev, ctx, eng):
    """configs[2]: keygen + sign of args.sign_keys one-time keys in all, this rank's block of signer indices."""
    import numpy as np
    import torch

    from fusion_cryptography_b200 import parallel

    args, dev, rank, world = dv.args, dv.device, dv.rank, dv.world
    ell, d = eng.ell, eng.d
    lo, hi = parallel.shard_bounds(args.sign_keys, rank, world)
    nk = hi - lo
    idx = np.arange(lo, hi)
    seeds = (10**6 + 2 * idx).astype(np.uint64)
    msgs = [f"msg-{i:07d}".encode() for i in idx]
    chunk = 2 * 640 * int(ctx.param("sm_count"))                    # Engine.keygen_sign's default: two full waves of the challenge-hash kernel
    eng.keygen_sign(seeds[:256], msgs[:256])                        # warm-up
    pm = ctx.pack_messages(msgs)                                    # messages resident in HBM before the timed region
    eng.keygen_sign(seeds, pm, chunk=chunk, keep_sigs=False)        # first full-size call pays cudaMalloc; the timed one reuses the buffers
    cs = ClockSampler(dv.local).start() if rank == 0 else None
    dv.barrier()
    w0 = time.time()
    t_ks, (vk_l, vk_r, _) = dv.timed(lambda: eng.keygen_sign(seeds, pm, chunk=chunk, keep_sigs=False))
    w1 = time.time()
    t_ks = dv.allmax(t_ks)
    clocks = None
    if cs:
        cs.mark(w0, w1)
        clocks = cs.stop()
    launches_per_chunk = 9   # seeds->keys, sampler, NTT, 2 x (A-sum, product), prehash, challenge hash + decode, NTT, signature rows (the rows on a second stream)
    # kernels of the leg, one by one, on a resident chunk
    nc = min(nk, chunk)
    kb = eng.keygen(seeds[:nc])
    dg = ctx.hash_msg(pm[:nc])
    c_hat = ctx.hash_ch(kb.vk_l, kb.vk_r, dg)
    sig = torch.empty((nc, ell, d), dtype=torch.int32, device=dev)
    t_hash, _ = dv.timed(lambda: ctx.hash_ch(kb.vk_l, kb.vk_r, dg), reps=3)
    t_rows, _ = dv.timed(lambda: ctx.sign_mac_rows(kb.sk_hat, c_hat, out=sig), reps=5)
    t_kg, _ = dv.timed(lambda: eng.keygen(seeds[:nc]), reps=3)
    del sig, kb, c_hat, dg
    # the general ell-row path (no use of quirk Q1): 2*ell generators, 2*ell transforms, ell-term products, ell-row signature MAC
    ng = min(args.general_keys, nk)
    rs = np.stack([np.repeat(seeds[:ng, None], ell, 1), np.repeat(seeds[:ng, None] + 1, ell, 1)], axis=1)     # the same keys, the long way round
    rs_d = torch.from_numpy(np.ascontiguousarray(rs).view(np.int64)).to(dev)
    gchunk = 8192
    vl_g, vr_g, _ = eng.keygen_sign_rows(rs_d, pm[:ng], chunk=gchunk, keep_sigs=False)
    t_gen, (vl_g, vr_g, _) = dv.timed(lambda: eng.keygen_sign_rows(rs_d, pm[:ng], chunk=gchunk, keep_sigs=False))
    t_gen = dv.allmax(t_gen)
    general_equal = bool(torch.equal(vl_g, vk_l[:ng]) and torch.equal(vr_g, vk_r[:ng]))
    ngm = min(ng, gchunk)
    kg = eng.keygen_rows(rs_d[:ngm])
    c_g = ctx.hash_ch(kg.vk_l, kg.vk_r, ctx.hash_msg(pm[:ngm]))
    sg = torch.empty((ngm, ell, d), dtype=torch.int32, device=dev)
    t_macg, _ = dv.timed(lambda: ctx.sign_mac(kg.sk_l, kg.sk_r, c_g, out=sg), reps=5)
    t_kgr, _ = dv.timed(lambda: ctx.keygen_rows(eng.A, rs_d[:ngm], out=(kg.sk_l, kg.sk_r)), reps=2)
    del kg, c_g, sg, rs_d, vl_g, vr_g
    torch.cuda.empty_cache()
    # end to end through host buffers: seeds + messages up, verification keys + signatures down, chunk by chunk
    e2e_n = min(nk, args.sign_e2e_keys)
    with numa_local(dv.local) as numa_note:
        seeds_h = torch.from_numpy(seeds[:e2e_n].view(np.int64).copy()).pin_memory()
        offs = np.zeros(e2e_n + 1, dtype=np.int64)
        np.cumsum([len(m) for m in msgs[:e2e_n]], out=offs[1:])
        buf_h = torch.frombuffer(bytearray(b"".join(msgs[:e2e_n]) + b"\0"), dtype=torch.uint8).pin_memory()
        offs_h = torch.from_numpy(offs).pin_memory()
        check = {"sum": 0}

        def sink(lo_, hi_, vl_h, vr_h, sig_h):     # the consumer of the host data: touches one word per chunk
            check["sum"] += int(sig_h[0, 0, 0]) + int(vl_h[-1, -1])

        eng.keygen_sign_host(seeds_h[:4096], buf_h, offs_h[:4097], sink, chunk=2048)     # warm-up, allocates and pins the staging buffers of this chunk size
        e2e_chunk = 16384
        eng.keygen_sign_host(seeds_h[:e2e_chunk], buf_h, offs_h[:e2e_chunk + 1], sink, chunk=e2e_chunk)
    dv.barrier()
    t0 = time.perf_counter()
    moved = eng.keygen_sign_host(seeds_h, buf_h, offs_h, sink, chunk=e2e_chunk)
    torch.cuda.synchronize()
    e2e_s = dv.allmax(time.perf_counter() - t0)
    eng._ksh = None
    del seeds_h, buf_h, offs_h, pm, vk_l, vk_r
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    peak, peak_src = peaks()
    mhz = (clocks or {}).get("sm_mhz") or 1965.0
    alu_peak = dv.alu_peak_perms(mhz)
    perms = PERMS_CH[SECPAR] * nc / t_hash
    rec = ncu_record("hash_ch_stream_kernel")
    n_chunks = -(-nk // chunk)
    return {
        "metric": METRIC["sign"], "value": args.sign_keys / t_ks, "unit": UNIT["sign"], "seconds": t_ks, "scaling": "strong", "dtype": "u32 / u64 (Keccak lanes)",
        "keys_total": args.sign_keys, "keys_this_rank": nk, "gpu_launches": launches_per_chunk * n_chunks, "clocks": clocks,
        "note": "the reference's seeded keys: under a seed all ell rows of a key matrix are the same polynomial (SURVEY quirk Q1), so keygen is 2 generators + 2 transforms "
                "per key and the signature is ell copies of one product -- bit-exact with keygen(params, seed) / sign(); the general ell-row cost is `general_rows`",
        "general_rows": {"value": ng * world / t_gen, "unit": "sign/s", "keys_per_gpu": ng, "seconds": t_gen, "equals_seeded_keys": general_equal,
                         "note": "fz_keygen_rows -> fz_hash_ch -> fz_sign_mac: 2*ell sampler generators, 2*ell transforms, two ell-term products and an ell-row MAC per key "
                                 "(row seeds chosen so that the keys are the seeded ones, computed the long way round)",
                         "keygen_rows_ms_per_chunk": 1000 * t_kgr, "chunk": ngm, "generators_per_sec": 2 * ell * ngm / t_kgr},
        "roofline": {"bound": "int32 alu pipe (LOP3 / SHF)", "kernel": "hash_ch_stream_kernel", "achieved": perms, "peak": alu_peak, "unit": "Keccak-f[1600]/s",
                     "frac": perms / alu_peak, "traffic": None,
                     "how": f"{PERMS_CH[SECPAR]:.0f} permutations executed per challenge (48 absorb + 61 squeeze; the reference's digest() call computes {PERMS_CH_REFERENCE[SECPAR]:.0f}, "
                            f"the last 15 squeezed blocks are never read by its decoder and are not computed here); peak = SMs * 4 sub-partitions * SM clock / (24 rounds * {KECCAK_ALU_INSTR_PER_ROUND} ALU-pipe instructions "
                            f"* {ALU_CYCLES_PER_WARP_INSTR} issue cycles / 32 lanes) at the sampled clock",
                     "sm_mhz_used": mhz, "ncu_alu_pipe_pct": rec.get("ncu_alu_pct"), "ncu_issue_active_pct": rec.get("ncu_issue_active_pct"),
                     "share_of_leg": t_hash * (nk / nc) / t_ks,
                     "other_kernels": {
                         "sign_rows_kernel": {"bound": "hbm (writes)", "GB/s": nc * ell * d * 4 / t_rows / 1e9, "frac": nc * ell * d * 4 / t_rows / 1e9 / peak, "ms": 1000 * t_rows,
                                              "bytes": "ell KiB written per key"},
                         "sign_mac_kernel (general rows)": {"bound": "hbm", "GB/s": ngm * 3 * ell * d * 4 / t_macg / 1e9, "frac": ngm * 3 * ell * d * 4 / t_macg / 1e9 / peak,
                                                            "ms": 1000 * t_macg, "bytes": "3 * ell KiB per key (SURVEY 8d: 257,024 B)",
                                                            "note": "the peak is a 1:1 read / write copy (torch copy_); this kernel reads two bytes per byte it writes, and "
                                                                    "HBM reads are faster than writes, so a fraction slightly above 1 of THAT peak is possible (nominal HBM3e: 8 TB/s)"},
                         "keygen (sampler + NTT + A.sk)": {"ms_per_chunk": 1000 * t_kg, "keys": nc}},
                     "hbm_peak": peak, "hbm_peak_source": peak_src},
        "e2e": {"value": e2e_n * world / e2e_s, "unit": UNIT["sign"], "h2d_bytes_per_step": moved["h2d_bytes"], "d2h_bytes_per_step": moved["d2h_bytes"],
                "seconds": e2e_s, "keys_per_gpu": e2e_n, "GB/s_d2h": moved["d2h_bytes"] / e2e_s / 1e9, "host_buffers": numa_note["note"],
                "note": "seeds and messages from pinned host memory, every verification key (2 KiB) and signature (ell KiB) copied back to pinned host memory: PCIe-bound"},
    }


def leg_verify(dv: Dev, ctx, eng, secpar=256, full=True):
    """configs[3]-[4]: aggregate + verify of args.verify_aggs capacity aggregates in all (this rank's block), sliding windows over a key pool."""
    import numpy as np
    import torch

    from fusion_cryptography_b200 import parallel

    args, dev, rank, world, dist = dv.args, dv.device, dv.rank, dv.world, dv.dist
    ell, d, cap = eng.ell, eng.d, ctx.capacity
    nk = args.pool_keys if secpar == 256 else min(args.pool_keys, 32768)
    seeds = (10**6 + 2 * (rank * nk + np.arange(nk))).astype(np.uint64)
    msgs = [f"msg-{rank * nk + i:07d}".encode() for i in range(nk)]
    pm = ctx.pack_messages(msgs)
    vk_l, vk_r, sigs = eng.keygen_sign(seeds, pm, chunk=65536)
    j0, j1 = parallel.shard_bounds(args.verify_aggs, rank, world)
    na = j1 - j0
    counts = [cap] * na
    base = ((j0 + np.arange(na))[:, None] * 97 + np.arange(cap)[None, :]) % nk
    sel_np = base.reshape(-1).astype(np.int32)
    sel = torch.from_numpy(sel_np).to(dev)
    # key i always signs msgs[i]: its one signature (and prehash) is reused by every window that contains it
    dg_pos = ctx.hash_msg(pm)[sel.long()].contiguous()
    kw = dict(sel=sel, digests=dg_pos)

    def clocked(fn):
        cs = ClockSampler(dv.local).start() if rank == 0 else None
        dv.barrier()
        w0 = time.time()
        t, out = dv.timed(fn)
        w1 = time.time()
        clk = None
        if cs:
            cs.mark(w0, w1)
            clk = cs.stop()
        return dv.allmax(t), out, clk

    eng.aggregate(vk_l, vk_r, None, sigs, counts, sig_sel=sel, **kw)          # first call: buffers come from cudaMalloc; a verifier service runs warm
    t_agg, aggs, clk_agg = clocked(lambda: eng.aggregate(vk_l, vk_r, None, sigs, counts, sig_sel=sel, **kw))
    eng.verify(vk_l, vk_r, None, aggs, counts, **kw)
    t_ver, codes, clk_ver = clocked(lambda: eng.verify(vk_l, vk_r, None, aggs, counts, **kw))
    ok = bool((codes == 0).all().item())
    out = {"secpar": secpar, "aggregates_total": args.verify_aggs, "aggregates_this_rank": na, "signers_per_aggregate": cap,
           "value": args.verify_aggs * cap / t_ver, "unit": UNIT["verify"], "seconds": t_ver, "all_accept": bool(dv.allmin_int(int(ok))),
           "aggregate": {"value": args.verify_aggs * cap / t_agg, "unit": UNIT["verify"], "seconds": t_agg, "clocks": clk_agg}, "clocks": clk_ver}
    if not full:
        del sigs, aggs
        torch.cuda.empty_cache()
        return out if rank == 0 else None
    # 1 % tampered aggregates must be rejected with the reference's reason code
    bad = aggs.clone()
    bad[::100, 0, 0] += 1
    codes_bad = eng.verify(vk_l, vk_r, None, bad, counts, **kw)
    tamper_ok = bool((codes_bad[::100] == 3).all().item()) and int((codes_bad != 0).sum().item()) == len(range(0, na, 100))
    del bad
    # C4: ONE capacity aggregate end to end (sort, challenges, the serial aggregation hash, the weighted sum) and its verification
    one = slice(0, cap)
    k1 = dict(sel=sel[one], digests=dg_pos[one])
    eng.aggregate(vk_l, vk_r, None, sigs, [cap], sig_sel=sel[one], **k1)
    t_one_a, agg1 = dv.timed(lambda: eng.aggregate(vk_l, vk_r, None, sigs, [cap], sig_sel=sel[one], **k1))
    eng.verify(vk_l, vk_r, None, agg1, [cap], **k1)
    t_one_v, code1 = dv.timed(lambda: eng.verify(vk_l, vk_r, None, agg1, [cap], **k1))
    chain_perms = cap * PERMS_AG[secpar]
    single = {"signers": cap, "aggregate_seconds": t_one_a, "verify_seconds": t_one_v, "accepted": bool((code1 == 0).all().item()),
              "equals_batched_result": bool(torch.equal(agg1[0], aggs[0])) if j0 == 0 else None,
              "note": "one serial SHAKE256 chain (about 197 K absorb + 82 K squeeze permutations) on one warp: latency-bound by construction"}
    # the streaming aggregation sum alone, on aggregates that share NO signature (nothing re-hits L2): 86,016 B per signer
    n_dis = max(1, min(64, nk // cap))
    sel_dis = torch.arange(n_dis * cap, dtype=torch.int32, device=dev)
    alpha_dis = ctx.hash_ch(vk_l[: n_dis * cap], vk_r[: n_dis * cap], ctx.hash_msg(pm[: n_dis * cap]))      # any NTT-domain weights do for the timing
    t_sum, _ = dv.timed(lambda: ctx.aggregate_sum(sigs, sel_dis, alpha_dis, [cap] * n_dis), reps=5)
    t_sum1, _ = dv.timed(lambda: ctx.aggregate_sum(sigs, sel_dis[:cap].contiguous(), alpha_dis[:cap].contiguous(), [cap]), reps=10)
    del alpha_dis, sel_dis
    # end to end through host buffers: key pool, message pool, key / message index per position and the aggregate signatures up, codes down
    with numa_local(dv.local) as numa_note:
        vl_h, vr_h = vk_l.cpu().pin_memory(), vk_r.cpu().pin_memory()
        sel_h = torch.from_numpy(sel_np).pin_memory()
        aggs_h = aggs.cpu().pin_memory()
        codes_h = torch.empty(na, dtype=torch.uint8, pin_memory=True)
    dv.barrier()
    t0 = time.perf_counter()
    moved = eng.verify_host(vl_h, vr_h, msgs, sel_h, sel_h, aggs_h, counts, codes_h)
    e2e_s = dv.allmax(time.perf_counter() - t0)
    e2e_ok = bool((codes_h == 0).all().item())
    del vl_h, vr_h, sel_h, aggs_h
    # N > 1: ONE capacity aggregate whose signers are sharded over the ranks (the NCCL exchange paths), against the single-GPU result
    sharded = None
    if world > 1 and secpar == 256:
        n1 = cap
        s1 = (10**6 + 2 * np.arange(n1)).astype(np.uint64)
        m1 = [f"msg-{i:07d}".encode() for i in range(n1)]
        kb = eng.keygen(s1)
        lo, hi = parallel.shard_bounds(n1, rank, world)
        from fusion_cryptography_b200.batch import KeyBatch

        loc = KeyBatch(sk_hat=kb.sk_hat[lo:hi].contiguous(), vk_l=kb.vk_l[lo:hi].contiguous(), vk_r=kb.vk_r[lo:hi].contiguous())
        local_sigs = eng.sign(loc, m1[lo:hi])
        parallel.aggregate_one_sharded(eng, kb.vk_l, kb.vk_r, m1, local_sigs)
        dv.barrier()
        t_sa, agg_s = dv.timed(lambda: parallel.aggregate_one_sharded(eng, kb.vk_l, kb.vk_r, m1, local_sigs))
        t_sa = dv.allmax(t_sa)
        parallel.verify_one_sharded(eng, kb.vk_l, kb.vk_r, m1, agg_s)
        dv.barrier()
        t_sv, code_s = dv.timed(lambda: parallel.verify_one_sharded(eng, kb.vk_l, kb.vk_r, m1, agg_s))
        t_sv = dv.allmax(t_sv)
        eq = 1
        if rank == 0:
            all_sigs = eng.sign(kb, m1)
            want = eng.aggregate(kb.vk_l, kb.vk_r, m1, all_sigs, [n1])
            eq = int(torch.equal(want, agg_s)) & int(eng.verify(kb.vk_l, kb.vk_r, m1, agg_s, [n1])[0].item() == 0)
            del all_sigs
        ref = agg_s.clone()
        dist.broadcast(ref, 0)
        same_all = dv.allmin_int(int(torch.equal(ref, agg_s)) & int(code_s[0].item() == 0))
        eqt = torch.tensor([eq], dtype=torch.int64, device=dev)
        dist.broadcast(eqt, 0)
        sharded = {"signers": n1, "ranks": world, "aggregate_one_sharded_seconds": t_sa, "verify_one_sharded_seconds": t_sv,
                   "equals_single_gpu": bool(eqt.item()), "identical_on_all_ranks": bool(same_all),
                   "collectives": "all-gather of the challenges (1 KiB per signer), all-gather + modular fold of the partial sums (ell KiB / 1 KiB per rank)",
                   "note": "the serial aggregation hash runs redundantly on every rank and is the latency floor: sharding ONE aggregate spreads its data, not its hash"}
        del kb, local_sigs, agg_s
    del sigs, aggs
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    peak, peak_src = peaks()
    mhz = (clk_ver or {}).get("sm_mhz") or 1965.0
    alu_peak = dv.alu_peak_perms(mhz)
    perms_per_signer = PERMS_CH[secpar] + PERMS_AG[secpar]
    perms = args.verify_aggs * cap * perms_per_signer / t_ver
    gbs = n_dis * cap * (ell + 1) * d * 4 / t_sum / 1e9
    rec = ncu_record("aggregate_kernel")
    # the chain layout launch_agw_absorb / launch_agw_squeeze pick for this many aggregates on one GPU (fz_hashstream.cu chain_layout)
    layout = "warp" if na <= 1776 else ("pair" if na < 40960 else "thread")
    thread_kernels = layout != "warp"
    chain_kernel = {"warp": "warpag::agw_absorb/squeeze_warp_kernel (a warp per aggregate)", "pair": "pairag::agw_absorb/squeeze_pair_kernel (two lanes per aggregate)",
                    "thread": "agw_absorb/squeeze_thread_kernel (a thread per aggregate)"}[layout]
    out.update({
        "metric": METRIC["verify"], "scaling": "strong", "dtype": "u32 / u64 (Keccak lanes)", "tampered_1pct_rejected": tamper_ok,
        "single_capacity_aggregate": single, "one_aggregate_sharded": sharded,
        "roofline": {
            "bound": "int32 alu pipe (LOP3 / SHF)" if thread_kernels else "latency of the serial SHAKE256 chains (one warp per aggregate)",
            "kernel": "hash_ch_stream_kernel + " + chain_kernel, "chain_layout": layout,
            "achieved": perms, "peak": alu_peak, "unit": "Keccak-f[1600]/s", "frac": perms / alu_peak, "traffic": None, "sm_mhz_used": mhz,
            "how": f"whole leg: {perms_per_signer:.1f} permutations per verified signature (hash_ch {PERMS_CH[secpar]:.0f} + hash_ag {PERMS_AG[secpar]:.1f}) / leg time, against the ALU-pipe "
                   "peak at the sampled clock (see the sign leg; a permutation in the pair layout costs 1.31 x the ALU-pipe instructions of the thread layout the peak "
                   "is counted in); with at most 1,776 aggregates per GPU the aggregation hash runs a warp per chain and the leg is bound by chain latency instead",
            "chain_latency": {"cycles_per_round_single_chain": t_one_v * mhz * 1e6 / (chain_perms * 24.0), "lab_floor_cycles_per_round": 168,
                              "note": "single capacity aggregate: verify seconds * SM clock / (chain permutations * 24 rounds), includes sort, challenges and the weighted sum; "
                                      "lab floor = tools/keccak_lab.cu V17 alone on a sub-partition (profiles/r01_keccak_lab_v3.txt); the five-lanes-per-chain layout "
                                      "measured 290 (profiles/r02_keccak_lab5_v21.txt)"},
            "aggregate_kernel": {"bound": "hbm", "GB/s": gbs, "frac": gbs / peak, "ms": 1000 * t_sum, "aggregates": n_dis, "bytes": "(ell + 1) KiB per signer (SURVEY 8d: 86,016 B)",
                                 "signatures_shared_between_aggregates": 0, "single_capacity_aggregate_us": 1e6 * t_sum1,
                                 "single_GB/s": cap * (ell + 1) * d * 4 / t_sum1 / 1e9, "ncu_dram_bytes_per_launch": rec.get("dram_bytes_per_launch"),
                                 "peak": peak, "peak_source": peak_src}},
        "e2e": {"value": args.verify_aggs * cap / e2e_s, "unit": UNIT["verify"], "h2d_bytes_per_step": moved["h2d_bytes"], "d2h_bytes_per_step": moved["d2h_bytes"],
                "seconds": e2e_s, "all_accept": e2e_ok, "host_buffers": numa_note["note"],
                "note": "key pool, message pool, key / message index per position and the aggregate signatures copied from pinned host memory, reason codes copied back"},
    })
    return out


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
            "matches_reference_digest": digest == "130985dc44f2699f1fcffb8fc95a2fbe61682ea97b51ba1d03c2c83c1d735105"}


def cpu_baselines(args):
    """N = 1, rank 0: the CPU figures of all three legs -- the unmodified reference in a process pool (when baseline/_ref travelled) and the C port."""
    out = {}
    p_sgn, p_ver = port_scheme_legs()
    port = {"ntt": port_ntt(), "sign": p_sgn, "verify": p_ver}
    if reference_present() and not args.port_only:
        pool = RefPool()
        per = 64
        t = pool.ntt_step(per)
        api_n = 3
        api_v = api_n * pool.cores / pool.ntt_api(api_n)
        ntt = {"value": per * pool.cores / t, "unit": UNIT["ntt"], "cores": pool.cores, "kind": "reference",
               "sample": f"{per * pool.cores} polynomials ({per} per process, {pool.cores} processes) through the reference's cooley_tukey_ntt, cent and gentleman_sande_intt "
                         f"(algebra/ntt.py) with prebuilt tables, {t:.2f} s; transform(transform(a) * b_hat) on the polynomial classes: {api_v:.1f} polys/s",
               "as_used_class_api_value": api_v}
        sgn, ver = reference_scheme_legs(pool)
        pool.close()
        out = {"ntt": ntt, "sign": sgn, "verify": ver}
        for k in out:
            out[k]["port"] = port[k]
    else:
        out = port
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--leg", default="ntt", choices=["ntt", "sign", "verify"], help="which leg is the top level of the JSON line (default: the NTT pass, configs[1])")
    ap.add_argument("--polys", type=int, default=N_POLYS, help="polynomials per GPU (default: the BASELINE size 2^20)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--port-only", action="store_true", help="CPU figures from the C port only, even when baseline/_ref is present")
    ap.add_argument("--no-legs", action="store_true", help="NTT leg only")
    ap.add_argument("--sign-keys", type=int, default=SIGN_KEYS, help="one-time keys in all for the sign leg (BASELINE configs[2]: 10^6)")
    ap.add_argument("--sign-e2e-keys", type=int, default=262144, help="keys per GPU pushed through the host-buffer pipeline of the sign leg")
    ap.add_argument("--general-keys", type=int, default=65536, help="keys per GPU for the general ell-row keygen + sign figure")
    ap.add_argument("--verify-aggs", type=int, default=VERIFY_AGGS, help="capacity aggregates in all for the verify leg (BASELINE configs[4]: 10^4)")
    ap.add_argument("--pool-keys", type=int, default=262144, help="one-time keys per GPU the aggregates' sliding windows are drawn from")
    ap.add_argument("--no-128", action="store_true", help="skip the secpar-128 half of configs[4]")
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
    from fusion_cryptography_b200.batch import Engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dv = Dev(args, rank, world, local_rank, dist if world > 1 else None)
    ctx = native.get_context(SECPAR, local_rank)
    legs = {}
    if args.leg == "ntt" or not args.no_legs:
        legs["ntt"] = leg_ntt(dv, ctx)
    if not args.no_legs or args.leg != "ntt":
        eng = Engine(SECPAR, seed=42, device=local_rank)
        if args.leg in ("ntt", "sign") or not args.no_legs:
            legs["sign"] = leg_sign(dv, ctx, eng)
        if args.leg in ("ntt", "verify") or not args.no_legs:
            legs["verify"] = leg_verify(dv, ctx, eng, 256, full=True)
            if not args.no_128:
                torch.cuda.empty_cache()
                eng128 = Engine(128, seed=42, device=local_rank)
                v128 = leg_verify(dv, eng128.ctx, eng128, 128, full=False)
                if rank == 0:
                    legs["verify"]["secpar_128"] = v128
    if rank == 0:
        demo = readme_demo() if not args.no_legs else None
        cpu = cpu_baselines(args) if (world == 1 and not args.no_cpu_baseline) else {}
        top = legs[args.leg]
        others = {k: v for k, v in legs.items() if k != args.leg and v is not None}
        for k, v in legs.items():
            if v is not None and k in cpu:
                v["cpu_baseline"] = cpu[k]
        line = {
            "metric": top["metric"], "value": top["value"], "unit": top["unit"], "n_gpus": world, "steps": args.steps if args.leg == "ntt" else 1,
            "warmup": args.warmup if args.leg == "ntt" else 1,
            "ms_per_step": top["ms_per_step"] if args.leg == "ntt" else 1000 * top["seconds"], "higher_is_better": True, "scaling": top["scaling"],
            "vs_baseline": None, "dtype": top["dtype"], "data": "synthetic", "config": static_config(args, args.leg),
            "roofline": dict(top["roofline"]), "e2e": dict(top["e2e"]), "gpu_launches": top.get("gpu_launches"), "clocks": top.get("clocks"),
        }
        if args.leg in cpu:
            line["cpu_baseline"] = dict(cpu[args.leg])
        # the other legs, whole, under the keys the driver keeps -- and once more as a top-level object
        slim = lambda v: {k: x for k, x in v.items() if k not in ("roofline", "e2e", "cpu_baseline")}
        line["roofline"]["legs"] = {k: dict(slim(v), roofline=v.get("roofline")) for k, v in others.items()}
        line["e2e"]["legs"] = {k: v.get("e2e") for k, v in others.items()}
        if "cpu_baseline" in line:
            line["cpu_baseline"]["legs"] = {k: cpu[k] for k in others if k in cpu}
        line["legs"] = {k: slim(v) for k, v in legs.items() if v is not None}
        if demo is not None:
            line["readme_demo"] = demo
            line["e2e"]["legs"]["readme_demo (configs[0], drop-in fusion.fusion API)"] = demo
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
