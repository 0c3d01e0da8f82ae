#!/usr/bin/env python3
"""Golden vectors for rings OTHER than the scheme's two: the unmodified Python reference's cooley_tukey_ntt,
gentleman_sande_intt, ntt_poly_mult and the polynomial classes' + / * on small and large moduli and degrees 2..8192.

Run in the build container only (needs /root/reference):   python tests/golden/make_ring_golden.py
Writes tests/golden/ring_golden.npz.  Nothing here is used by the product."""
import os
import sys

import numpy as np

REF = os.environ.get("FUSION_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)

from algebra.ntt import bit_reverse_copy, cooley_tukey_ntt, gentleman_sande_intt, ntt_poly_mult  # noqa: E402
from algebra.polynomials import (PolynomialCoefficientRepresentation as Coef, PolynomialNTTRepresentation as Ntt,  # noqa: E402
                                 sample_polynomial_coefficient_representation, sample_polynomial_ntt_representation, transform)

OUT = os.path.dirname(os.path.abspath(__file__))
# (degree, modulus): the small pairs are of the kind the reference's tests/test_ntt.py:18-50 generates; the rest stretch the
# degree (up to 8192) and the modulus (up to just under 2^31, incl. the scheme's prime with a degree the scheme does not use)
RINGS = [(2, 5), (4, 17), (8, 17), (4, 97), (16, 97), (32, 193), (64, 257), (64, 7681), (32, 65537), (128, 65537), (256, 7681),
         (1024, 12289), (2048, 12289), (128, 2147465729), (8192, 2013265921), (64, 2147483137)]


def primitive_root(q, order):
    g = 2
    while True:
        x = pow(g, (q - 1) // order, q)
        if pow(x, order // 2, q) != 1:
            return x
        g += 1


def main():
    rng = np.random.default_rng(0xF0512)
    out = {"rings": np.array(RINGS, dtype=np.int64)}
    for idx, (d, q) in enumerate(RINGS):
        assert (q - 1) % (2 * d) == 0
        root = primitive_root(q, 2 * d)
        inv_root = pow(root, q - 2, q)
        tw = bit_reverse_copy([pow(root, i, q) for i in range(d)])
        itw = bit_reverse_copy([pow(inv_root, i, q) for i in range(d)])
        h = q // 2
        n = 4 if d <= 2048 else 2
        a = rng.integers(-h, h + 1, size=(n, d)).astype(np.int64)
        b = rng.integers(-h, h + 1, size=(n, d)).astype(np.int64)
        a[0, :] = h            # extremes
        b[0, :] = -h
        a[1, ::2] = 0
        fwd = np.array([cooley_tukey_ntt([int(x) for x in row], q, 2 * d, tw) for row in a], dtype=np.int64)
        inv = np.array([gentleman_sande_intt([int(x) for x in row], q, 2 * d, itw) for row in b], dtype=np.int64)
        prod = np.array([ntt_poly_mult([int(x) for x in ra], [int(x) for x in rb], q, root, inv_root, 2 * d) for ra, rb in zip(a, b)], dtype=np.int64)
        pa = [Ntt(modulus=q, degree=d, root=root, inv_root=inv_root, root_order=2 * d, values=[int(x) for x in row]) for row in a]
        pb = [Ntt(modulus=q, degree=d, root=root, inv_root=inv_root, root_order=2 * d, values=[int(x) for x in row]) for row in b]
        mul = np.array([(x * y).values for x, y in zip(pa, pb)], dtype=np.int64)
        add = np.array([(x + y).values for x, y in zip(pa, pb)], dtype=np.int64)
        # the class path: transform() of a coefficient polynomial and back (must equal the raw functions)
        ca = Coef(modulus=q, degree=d, root=root, inv_root=inv_root, root_order=2 * d, coefficients=[int(x) for x in a[-1]])
        ta = transform(ca)
        assert ta.values == [int(x) for x in fwd[-1]] and transform(ta).coefficients == [int(x) for x in a[-1]]
        if d <= 64:  # the schoolbook product of the coefficient class (polynomials.py:171-216) equals the NTT product
            cb = Coef(modulus=q, degree=d, root=root, inv_root=inv_root, root_order=2 * d, coefficients=[int(x) for x in b[-1]])
            assert (ca * cb).coefficients == [int(x) for x in prod[-1]]
        for k, v in (("root", [root, inv_root]), ("tw", tw), ("itw", itw), ("a", a), ("b", b), ("fwd", fwd), ("inv", inv), ("prod", prod), ("mul", mul),
                     ("add", add)):
            out[f"r{idx}_{k}"] = np.array(v, dtype=np.int64)
        print(f"ring d={d} q={q} root={root}: ok")
    # seeded samplers on rings other than the scheme's (the reference's tests/test_polynomials.py:883-909 draws from
    # modulus 65537, degree 1024, norm_bound 1000, weight_bound 100, seed 123456789): (d, q, norm_bound, weight_bound, seed)
    import random
    cases = [(1024, 65537, 1000, 100, 123456789), (64, 257, 1000, 64, 7), (64, 257, 5, 200, 2**70 + 3), (8, 17, 3, 0, 11), (256, 7681, 1, 17, -5),
             (32, 2147483137, 2**40, 31, 99)]
    out["sample_cases"] = np.array([[c[0], c[1], min(c[2], 2**62), c[3], 0] for c in cases], dtype=np.int64)
    out["sample_seeds"] = np.array([str(c[4]) for c in cases])
    for i, (d, q, nb, wb, seed) in enumerate(cases):
        root = primitive_root(q, 2 * d)
        inv_root = pow(root, q - 2, q)
        f = sample_polynomial_coefficient_representation(modulus=q, degree=d, root=root, inv_root=inv_root, root_order=2 * d, norm_bound=nb,
                                                         weight_bound=wb, seed=seed)
        out[f"s{i}_coef"] = np.array(f.coefficients, dtype=np.int64)
        out[f"s{i}_after_coef"] = np.array([random.getrandbits(32) for _ in range(4)], dtype=np.int64)   # the global generator's next outputs
        g = sample_polynomial_ntt_representation(modulus=q, degree=d, root=root, inv_root=inv_root, root_order=2 * d, seed=seed)
        out[f"s{i}_ntt"] = np.array(g.values, dtype=np.int64)
        out[f"s{i}_after_ntt"] = np.array([random.getrandbits(32) for _ in range(4)], dtype=np.int64)
        out[f"s{i}_root"] = np.array([root, inv_root], dtype=np.int64)
    # the decoder on rings other than the scheme's (the reference's tests/test_fusion.py:475-558 decodes for modulus 65537, degree 1024,
    # norm_bound 1000, weight_bound 100, log2_bias 256): (log2_bias, modulus, degree, norm_bound, weight_bound, byte pattern, length)
    from math import ceil, log2
    import fusion.fusion as ff

    def sizes(lb, q, d, nb, wb):
        bound = max(1, min(q // 2, nb))
        bpc, bpi = ceil((log2(bound) + 1 + lb) / 8), ceil((log2(d) + lb) / 8)
        return ceil(wb / 8) + (bpc + bpi) * wb, ceil(wb / 8) + bpc * wb + bpi * d

    dcases = []
    for lb, q, d, nb, wb in [(256, 65537, 1024, 1000, 100), (128, 257, 64, 16, 64), (17, 17, 8, 3, 0), (64, 12289, 2048, 10**6, 300), (8, 7681, 256, 1, 17),
                             (128, 2013265921, 512, 2**29, 257)]:
        need, full = sizes(lb, q, d, nb, wb)
        for pat, ln in (("zero", full), ("ones", full), ("rand", full), ("rand", need), ("rand", (need + full) // 2 + 1)):
            dcases.append((lb, q, d, nb, wb, pat, ln))
    out["decode_cases"] = np.array([[c[0], c[1], c[2], c[3], c[4], c[6]] for c in dcases], dtype=np.int64)
    for i, (lb, q, d, nb, wb, pat, ln) in enumerate(dcases):
        b = bytes(ln) if pat == "zero" else b"\xff" * ln if pat == "ones" else bytes(rng.integers(0, 256, size=ln, dtype=np.uint8))
        out[f"d{i}_bytes"] = np.frombuffer(b, dtype=np.uint8)
        out[f"d{i}_coefs"] = np.array(ff.decode_bytes_to_polynomial_coefficients(b=b, log2_bias=lb, modulus=q, degree=d, norm_bound=nb, weight_bound=wb),
                                      dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "ring_golden.npz"), **out)


if __name__ == "__main__":
    main()
