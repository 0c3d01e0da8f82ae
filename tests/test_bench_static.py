"""CPU-side checks of bench.py's static parts: both arms describe the same workload, and the permutation counts its hash rooflines divide by are
the ones the scheme's parameters give (fusion.py:436-438, 473, 517-524), not a remembered constant."""
import argparse
import importlib.util
import os
from math import ceil, log2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)   # importing it needs neither a GPU nor the extension
    return mod


def _args(b, **over):
    ns = dict(gpus=1, steps=20, warmup=5, impl="b200", leg="ntt", polys=b.N_POLYS, sign_keys=b.SIGN_KEYS, sign_e2e_keys=262144, general_keys=65536,
              verify_aggs=b.VERIFY_AGGS, pool_keys=262144, no_128=False, no_legs=False, no_cpu_baseline=False, port_only=False)
    ns.update(over)
    return argparse.Namespace(**ns)


def test_both_arms_print_the_same_config():
    b = _bench()
    for leg in ("ntt", "sign", "verify"):
        assert b.static_config(_args(b, impl="b200"), leg) == b.static_config(_args(b, impl="reference"), leg)
    cfg = b.static_config(_args(b))
    assert cfg["polys_per_gpu"] == 1 << 20 and set(cfg["legs"]) == {"sign", "verify"}
    assert "model" not in cfg and "workload" in cfg


def test_hash_ch_permutation_counts_follow_from_the_parameters():
    b = _bench()
    # secpar -> (degree, omega_ch, typical bytes of "dst,str(vk),str(prehash)" as the reference formats it: measured 6,426-6,456 at 256 and
    # 2,016-2,041 at 128 over seeded keys; the block count is 48 for every length in [6,392, 6,527], 15 in [1,904, 2,039])
    levels = {256: (256, 60, (6426, 6456)), 128: (64, 27, (2016, 2039))}
    for secpar, (d, omega, (lo, hi)) in levels.items():
        bpc = ceil((log2(1) + 1 + secpar) / 8)          # beta_ch = 1
        bpi = ceil((log2(d) + secpar) / 8)
        n_requested = ceil(omega / 8) + bpc * omega + d * bpi               # fusion.py:520-524: what digest(n) computes
        n_read = ceil(omega / 8) + bpc * omega + (d - 1 - omega) * bpi      # fusion.py:436-438, 473: what the decoder reads
        absorb_lo, absorb_hi = lo // 136 + 1, hi // 136 + 1
        assert absorb_lo == absorb_hi, "the text length range must sit inside one block count"
        executed = absorb_lo + ceil(n_read / 136) - 1        # the first squeezed block needs no further permutation
        reference = absorb_lo + ceil(n_requested / 136) - 1
        assert b.PERMS_CH[secpar] == executed, (secpar, executed)
        assert b.PERMS_CH_REFERENCE[secpar] >= reference - 1   # SURVEY 8d's figure counts the absorb and squeeze blocks (48 + 77, 15 + 12)
        assert executed < reference
