#!/usr/bin/env python3
"""Digests of every KIND of known-answer row the reference's KAT generator writes (KATs/generate_KAT_values.py), produced by the UNMODIFIED
reference on fixed seeds -- including the kinds whose CSVs are missing from the reference's repository (.MISSING_LARGE_BLOBS: keygen, sign,
hash_ag, hash_vks_and_ints_and_challs_to_bytes, aggregate at 256, ...) and the generator's habit of passing (sk, vk) TUPLES as keys to
hash_ag / hash_vks_and_ints_and_challs_to_bytes (:120-133; the reference only takes str(key) there).

Run in the build container only (needs /root/reference):   python tests/golden/make_kat_golden.py      -> tests/golden/kat_kinds_golden.json"""
import hashlib
import json
import os
import sys
from math import ceil, log2

REF = os.environ.get("FUSION_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
import fusion.fusion as ff  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def sha(x) -> str:
    return hashlib.sha256(str(x).encode()).hexdigest()


def main():
    J = {}
    for secpar, seed_a, key_seeds in ((128, 3141592653, [271828182 + 2 * i * 7919 for i in range(10)]), (256, 2718281828, [314159265 + 2 * i * 104729 for i in range(10)])):
        p = ff.fusion_setup(secpar, seed_a)
        msgs = [str(i) for i in range(10)]
        otks = [ff.keygen(p, s) for s in key_seeds]
        vks = [k[1] for k in otks]
        pre = [ff.hash_message_to_int(p, m) for m in msgs]
        bound = max(0, min(p.modulus // 2, p.beta_ch))
        n = ceil(p.omega_ch / 8) + ceil((log2(bound) + 1 + p.secpar) / 8) * max(0, min(p.degree, p.omega_ch)) + p.degree * ceil((log2(p.degree) + p.secpar) / 8)
        xof = [ff.hash_vk_and_int_to_bytes(p, v, i, n) for v, i in zip(vks, pre)]
        challs = [ff.hash_ch(p, v, m) for v, m in zip(vks, msgs)]
        sigs = [ff.sign(p, k, m) for k, m in zip(otks, msgs)]
        agg = ff.aggregate(p, vks, msgs, sigs)
        J[str(secpar)] = {
            "seed_a": seed_a, "key_seeds": key_seeds, "n_ch": n,
            "setup": sha(p), "keygen": [sha(k) for k in otks], "prehash": [str(i) for i in pre], "xof": [sha(x) for x in xof],
            "hash_ch": [sha(c) for c in challs], "sign": [sha(s) for s in sigs],
            "ag_bytes_vks": sha(ff.hash_vks_and_ints_and_challs_to_bytes(p, vks, pre, challs)),
            "ag_bytes_tuples": sha(ff.hash_vks_and_ints_and_challs_to_bytes(p, otks, pre, challs)),
            "hash_ag_vks": sha(ff.hash_ag(p, vks, msgs)), "hash_ag_tuples": sha(ff.hash_ag(p, otks, msgs)),
            "aggregate": sha(agg), "verify": list(ff.verify(p, vks, msgs, agg)),
        }
        print(secpar, "ok")
    json.dump(J, open(os.path.join(OUT, "kat_kinds_golden.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
