#!/usr/bin/env python3
"""Generate the golden fixtures in tests/golden/ from the UNMODIFIED Python reference.

Run in the build container only (needs /root/reference, which does not exist on the
GPU box):

    python tests/golden/make_golden.py

It imports the reference (fusion.fusion, algebra.*) from /root/reference, runs it on
seeded inputs, parses the reference's own valid KAT CSVs (KATs/KAT_values)
and stores compact fixtures:

    ntt_golden.npz      cent / twiddles / cooley_tukey_ntt / gentleman_sande_intt / pointwise
    scheme_golden.npz   setup, keygen, hash_ch, sign, hash_ag, aggregate, verify (both secpars)
    kat_golden.npz      rows of the reference's committed KAT CSVs (setup 128/256, hash_* 128)
    golden.json         strings, digests and byte strings that do not fit an int array

Nothing here is used by the product; tests read the fixtures, never /root/reference.
"""
import csv
import hashlib
import json
import os
import random
import re
import sys

import numpy as np

REF = os.environ.get("FUSION_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)

from algebra.ntt import bit_reverse_copy, cent, cooley_tukey_ntt, gentleman_sande_intt  # noqa: E402
from algebra.polynomials import (  # noqa: E402
    PolynomialCoefficientRepresentation,
    PolynomialNTTRepresentation,
    sample_polynomial_coefficient_representation,
    transform,
)
import fusion.fusion as ff  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
Q = 2147465729
J = {}


def sha(s) -> str:
    if isinstance(s, str):
        s = s.encode()
    return hashlib.sha256(s).hexdigest()


def tw_tables(p):
    rp = [pow(p["root"], i, Q) for i in range(p["degree"])]
    irp = [pow(p["inv_root"], i, Q) for i in range(p["degree"])]
    return bit_reverse_copy(rp), bit_reverse_copy(irp)


def gen_ntt():
    rng = np.random.default_rng(0xF0510)
    out = {}
    vals = [0, 1, -1, Q, -Q, Q // 2, Q // 2 + 1, -(Q // 2), -(Q // 2) - 1, Q * Q, -Q * Q + 5, 2**61 - 1, -(2**61) + 3]
    vals += [int(x) for x in rng.integers(-(2**62), 2**62, size=500)]
    out["cent_in"] = np.array(vals, dtype=np.int64)
    out["cent_out_30"] = np.array([cent(v, Q, Q // 2, 30) for v in vals], dtype=np.int64)
    out["cent_out_31"] = np.array([cent(v, Q, Q // 2, 31) for v in vals], dtype=np.int64)
    for secpar in (128, 256):
        p = ff.PREFIX_PARAMETERS[secpar]
        d = p["degree"]
        tw, itw = tw_tables(p)
        out[f"tw_{secpar}"] = np.array(tw, dtype=np.int64)
        out[f"itw_{secpar}"] = np.array(itw, dtype=np.int64)
        n = 24
        a = rng.integers(-(Q // 2), Q // 2 + 1, size=(n, d)).astype(np.int64)
        # edge rows: zeros, all max, all min, delta, non-canonical int32 extremes
        a[0] = 0
        a[1] = Q // 2
        a[2] = -(Q // 2)
        a[3] = 0
        a[3, 0] = 1
        a[4] = 0
        a[4, d - 1] = -1
        a[5] = rng.integers(-(2**31), 2**31, size=d)  # any int32 (Q8: same result as reduced input)
        a[6] = 2**31 - 1
        a[7] = -(2**31)
        fwd = np.zeros_like(a)
        inv = np.zeros_like(a)
        for i in range(n):
            v = [int(x) for x in a[i]]
            cooley_tukey_ntt(val=v, modulus=Q, root_order=2 * d, bit_rev_root_powers=tw)
            fwd[i] = v
            w = [int(x) for x in a[i]]
            gentleman_sande_intt(val=w, modulus=Q, root_order=2 * d, bit_rev_inv_root_powers=itw)
            inv[i] = w
        b = rng.integers(-(Q // 2), Q // 2 + 1, size=(n, d)).astype(np.int64)
        mk = lambda vals: PolynomialNTTRepresentation(modulus=Q, degree=d, root=p["root"], inv_root=p["inv_root"], root_order=2 * d, values=[int(x) for x in vals])
        mul = np.zeros_like(a)
        muladd = np.zeros_like(a)
        ac = np.array([[cent(int(x), Q, Q // 2, 30) for x in row] for row in a], dtype=np.int64)
        for i in range(n):
            if i in (0, 3, 4):  # x * zero-poly returns int 0 in the reference (Q6 edge) -- skip
                mul[i] = [cent(int(x) * int(y), Q, Q // 2, 30) for x, y in zip(ac[i], b[i])]
                muladd[i] = [cent(int(m) + int(z), Q, Q // 2, 30) for m, z in zip(mul[i], fwd[i])]
                continue
            m = mk(ac[i]) * mk(b[i])
            mul[i] = m.values
            muladd[i] = (m + mk(fwd[i])).values
        # convolution through the reference: transform(transform(x)*transform(y))
        conv = np.zeros((4, d), dtype=np.int64)
        for i in range(4):
            mkc = lambda vals: PolynomialCoefficientRepresentation(modulus=Q, degree=d, root=p["root"], inv_root=p["inv_root"], root_order=2 * d, coefficients=[int(x) for x in vals])
            conv[i] = transform(transform(mkc(ac[8 + i])) * transform(mkc(b[8 + i]))).coefficients
        out.update({f"a_{secpar}": a, f"b_{secpar}": b, f"fwd_{secpar}": fwd, f"inv_{secpar}": inv,
                    f"mul_{secpar}": mul, f"muladd_{secpar}": muladd, f"conv_{secpar}": conv})
    np.savez_compressed(os.path.join(OUT, "ntt_golden.npz"), **out)


def poly_vals(gm):
    return [[list(e.values) for e in row] for row in gm.matrix]


def gen_scheme():
    out = {}
    for secpar, nkeys in ((128, 5), (256, 3)):
        tag = str(secpar)
        setup_seed = 42
        params = ff.fusion_setup(secpar, setup_seed)
        d, ell = params.degree, params.num_rows_sk
        A = np.array(poly_vals(params.public_challenge)[0], dtype=np.int64)  # [ell][d]
        out[f"A_{tag}"] = A
        J[f"params_str_sha_{tag}"] = sha(str(params))
        seeds = [1_000_000 + 2 * i for i in range(nkeys)]
        seeds[-1] = 2**40 + 12345  # two-limb seed
        msgs = [f"msg-{i:07d}" for i in range(nkeys)]
        msgs[1] = "héllo wörld ✓ " + "x" * 200  # multi-block, non-ASCII prehash input
        msgs[2] = ""
        keys = [ff.keygen(params, s) for s in seeds]
        out[f"seeds_{tag}"] = np.array(seeds, dtype=np.uint64)
        J[f"msgs_{tag}"] = msgs
        out[f"sk_l_row0_{tag}"] = np.array([poly_vals(k[0].left_sk_hat)[0][0] for k in keys], dtype=np.int64)
        out[f"sk_r_row0_{tag}"] = np.array([poly_vals(k[0].right_sk_hat)[0][0] for k in keys], dtype=np.int64)
        J[f"sk_rows_identical_{tag}"] = all(
            poly_vals(k[0].left_sk_hat)[r][0] == poly_vals(k[0].left_sk_hat)[0][0] for k in keys for r in range(ell))
        out[f"vk_l_{tag}"] = np.array([poly_vals(k[1].left_vk_hat)[0][0] for k in keys], dtype=np.int64)
        out[f"vk_r_{tag}"] = np.array([poly_vals(k[1].right_vk_hat)[0][0] for k in keys], dtype=np.int64)
        J[f"key_str_sha_{tag}"] = [sha(str(k)) for k in keys]
        J[f"vk_str_{tag}_0"] = str(keys[0][1])
        # coefficient-domain secret key samples (sampler pin)
        out[f"sk_coef_{tag}"] = np.array([
            sample_polynomial_coefficient_representation(modulus=Q, degree=d, root=params.root, inv_root=params.inv_root,
                                                         root_order=params.root_order, norm_bound=params.beta_sk,
                                                         weight_bound=params.omega_sk, seed=s).coefficients
            for s in seeds], dtype=np.int64)
        pre = [ff.hash_message_to_int(params, m) for m in msgs]
        J[f"prehash_{tag}"] = [str(x) for x in pre]
        n_ch = (-(-params.omega_ch // 8) + params.bytes_for_one_coef_bdd_by_beta_ch * params.omega_ch
                + params.degree * (params.bytes_for_poly_shuffle // params.degree))
        J[f"n_ch_{tag}"] = n_ch
        xofs = [ff.hash_vk_and_int_to_bytes(params, k[1], x, n_ch) for k, x in zip(keys, pre)]
        J[f"ch_xof_hex_{tag}_0"] = xofs[0].hex()
        J[f"ch_xof_sha_{tag}"] = [sha(x) for x in xofs]
        c_coefs = [ff.decode_bytes_to_polynomial_coefficients(b=x, log2_bias=secpar, modulus=Q, degree=d,
                                                              norm_bound=params.beta_ch, weight_bound=params.omega_ch) for x in xofs]
        out[f"c_coef_{tag}"] = np.array(c_coefs, dtype=np.int64)
        challs = [ff.hash_ch(params, k[1], m) for k, m in zip(keys, msgs)]
        out[f"c_hat_{tag}"] = np.array([c.c_hat.values for c in challs], dtype=np.int64)
        J[f"ch_str_{tag}_0"] = str(challs[0])
        sigs = [ff.sign(params, k, m) for k, m in zip(keys, msgs)]
        out[f"sig_row0_{tag}"] = np.array([poly_vals(s.signature_hat)[0][0] for s in sigs], dtype=np.int64)
        J[f"sig_str_sha_{tag}"] = [sha(str(s)) for s in sigs]
        vks = [k[1] for k in keys]
        order = sorted(range(nkeys), key=lambda i: str(vks[i]))
        out[f"order_{tag}"] = np.array(order, dtype=np.int64)
        svks = [vks[i] for i in order]
        smsgs = [msgs[i] for i in order]
        alphas = ff.hash_ag(params, svks, smsgs)
        out[f"alpha_hat_sorted_{tag}"] = np.array([a.alpha_hat.values for a in alphas], dtype=np.int64)
        ag_bytes = ff.hash_vks_and_ints_and_challs_to_bytes(params, svks, [pre[i] for i in order], [challs[i] for i in order])
        J[f"ag_xof_sha_{tag}"] = sha(ag_bytes)
        J[f"ag_xof_len_{tag}"] = len(ag_bytes)
        n_ag = len(ag_bytes) // nkeys
        out[f"alpha_coef_sorted_{tag}"] = np.array([
            ff.decode_bytes_to_polynomial_coefficients(b=ag_bytes[i * n_ag:(i + 1) * n_ag], log2_bias=secpar, modulus=Q, degree=d,
                                                       norm_bound=params.beta_ag, weight_bound=params.omega_ag)
            for i in range(nkeys)], dtype=np.int64)
        agg = ff.aggregate(params, vks, msgs, sigs)
        out[f"agg_{tag}"] = np.array([r[0] for r in poly_vals(agg.signature_hat)], dtype=np.int64)  # [ell][d]
        J[f"agg_str_sha_{tag}"] = sha(str(agg))
        ok = ff.verify(params, vks, msgs, agg)
        J[f"verify_ok_{tag}"] = [bool(ok[0]), ok[1]]
        # tampered aggregate (one value + 1): tests/test_fusion.py:861-873
        import copy
        bad = copy.deepcopy(agg)
        bad.signature_hat.matrix[0][0].values[0] += 1
        rb = ff.verify(params, vks, msgs, bad)
        J[f"verify_tamper_{tag}"] = [bool(rb[0]), rb[1]]
        # wrong message
        wm = list(msgs)
        wm[0] = "other"
        rw = ff.verify(params, vks, wm, agg)
        J[f"verify_wrongmsg_{tag}"] = [bool(rw[0]), rw[1]]
        rc = ff.verify(params, vks, msgs[:-1], agg)
        J[f"verify_count_{tag}"] = [bool(rc[0]), rc[1]]
        # norm failure: aggregate whose coefficient representation is large but consistent is hard to
        # forge; instead record beta_vf for the norm check
        J[f"beta_vf_{tag}"] = params.beta_vf
        # single-signer aggregate
        agg1 = ff.aggregate(params, vks[:1], msgs[:1], sigs[:1])
        out[f"agg1_row0_{tag}"] = np.array(poly_vals(agg1.signature_hat)[0][0], dtype=np.int64)
        J[f"verify1_{tag}"] = list(ff.verify(params, vks[:1], msgs[:1], agg1))
    # README demo, secpar 256 seed 42
    import string
    a = ff.fusion_setup(256, 42)
    keys = [ff.keygen(a, 42) for _ in range(2)]
    messages = ["".join(random.choices(string.ascii_letters + string.digits, k=20)) for _ in range(2)]
    sigs = [ff.sign(a, k, m) for k, m in zip(keys, messages)]
    vks = [k[1] for k in keys]
    agg = ff.aggregate(a, vks, messages, sigs)
    J["demo"] = {"messages": messages, "params_sha": sha(str(a)), "key0_sha": sha(str(keys[0])), "sigs_sha": sha(str(sigs)),
                 "agg_sha": sha(str(agg)), "verify": list(ff.verify(a, vks, messages, agg))}
    # decoder goldens: tests/test_fusion.py:475-558 style (all-0 / all-1 bytes) + random + short (Q4)
    rng = random.Random(1234)
    dec = []
    for secpar in (128, 256):
        p = ff.PREFIX_PARAMETERS[secpar]
        d = p["degree"]
        for wb, nb, blen, fill in ((p["omega_ch"], 1, 12000, "rand"), (p["omega_ag"], 1, 3968 if secpar == 256 else 1195, "rand"),
                                   (p["omega_ch"], 1, 12000, "zero"), (p["omega_ch"], 1, 12000, "ones"),
                                   (p["omega_ch"], 52, 12000, "rand"), (5, 3, 12000, "rand"), (d, 7, 30000, "rand")):
            if fill == "rand":
                b = bytes(rng.getrandbits(8) for _ in range(blen))
            elif fill == "zero":
                b = bytes(blen)
            else:
                b = b"\xff" * blen
            try:
                co = ff.decode_bytes_to_polynomial_coefficients(b=b, log2_bias=secpar, modulus=Q, degree=d, norm_bound=nb, weight_bound=wb)
            except ValueError:
                co = None
            dec.append({"secpar": secpar, "weight": wb, "bound": nb, "bytes_hex": b.hex(), "coefs": co})
        dec.append({"secpar": secpar, "weight": p["omega_ch"], "bound": 1, "bytes_hex": bytes(10).hex(), "coefs": None})
    J["decoder"] = dec
    np.savez_compressed(os.path.join(OUT, "scheme_golden.npz"), **out)


VAL_RE = re.compile(r"values=\[([^\]]*)\]")


def gen_kats():
    """The reference's own committed KAT CSVs that re-derive from current code."""
    csv.field_size_limit(10**9)
    out = {}
    kd = os.path.join(REF, "KATs", "KAT_values")
    for secpar in (128, 256):
        rows = list(csv.reader(open(os.path.join(kd, f"fusion_setup_KAT_{secpar}.csv"))))
        seeds, firsts, shas = [], [], []
        for cin, cout in rows:
            sp, seed = eval(cin)
            assert sp == secpar
            seeds.append(seed)
            firsts.append([int(x) for x in VAL_RE.search(cout).group(1).split(",")])
            shas.append(sha(cout))
            assert str(ff.fusion_setup(secpar, seed)) == cout, "setup KAT does not re-derive"
        out[f"setup_seeds_{secpar}"] = np.array(seeds, dtype=np.uint64)
        out[f"setup_poly_{secpar}"] = np.array(firsts, dtype=np.int64)
        J[f"kat_setup_sha_{secpar}"] = shas
    # hash KATs, secpar 128 only (the 256 files are among the missing blobs)
    def vk_of(cell):
        m = VAL_RE.findall(cell)
        return [int(x) for x in m[-2].split(",")], [int(x) for x in m[-1].split(",")]

    rows = list(csv.reader(open(os.path.join(kd, "intermediate_hash_message_to_int_KAT_128.csv"))))
    J["kat_hash_msg_128"] = [[re.search(r", '([^']*)'\)$", cin).group(1), cout] for cin, cout in rows]
    rows = list(csv.reader(open(os.path.join(kd, "intermediate_hash_vk_and_int_to_bytes_to_int_KAT_128.csv"))))
    vls, vrs, ints, ns, outs = [], [], [], [], []
    for cin, cout in rows:
        vl, vr = vk_of(cin)
        tail = cin[cin.rfind("]]))") + 4:]
        m = re.match(r", (\d+), (\d+)\)$", tail)
        vls.append(vl); vrs.append(vr); ints.append(m.group(1)); ns.append(int(m.group(2)))
        outs.append(eval(cout).hex())
    out["kat_hvi_vk_l_128"] = np.array(vls, dtype=np.int64)
    out["kat_hvi_vk_r_128"] = np.array(vrs, dtype=np.int64)
    J["kat_hvi_128"] = {"ints": ints, "n": ns, "out_hex": outs}
    rows = list(csv.reader(open(os.path.join(kd, "intermediate_hash_ch_KAT_128.csv"))))
    vls, vrs, msgs, chs = [], [], [], []
    for cin, cout in rows:
        vl, vr = vk_of(cin)
        tail = cin[cin.rfind("]]))") + 4:]
        msgs.append(re.match(r", '([^']*)'\)$", tail).group(1))
        vls.append(vl); vrs.append(vr)
        chs.append([int(x) for x in VAL_RE.search(cout).group(1).split(",")])
    out["kat_ch_vk_l_128"] = np.array(vls, dtype=np.int64)
    out["kat_ch_vk_r_128"] = np.array(vrs, dtype=np.int64)
    out["kat_ch_c_hat_128"] = np.array(chs, dtype=np.int64)
    J["kat_ch_msgs_128"] = msgs
    np.savez_compressed(os.path.join(OUT, "kat_golden.npz"), **out)


if __name__ == "__main__":
    gen_ntt()
    print("ntt done", flush=True)
    gen_scheme()
    print("scheme done", flush=True)
    gen_kats()
    print("kats done", flush=True)
    with open(os.path.join(OUT, "golden.json"), "w") as f:
        json.dump(J, f, indent=0, sort_keys=True)
    for fn in sorted(os.listdir(OUT)):
        print(fn, os.path.getsize(os.path.join(OUT, fn)))
