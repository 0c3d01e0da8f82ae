#!/usr/bin/env python3
"""Golden vectors for the GENERAL ell-row keygen + sign (no quirk Q1), from the UNMODIFIED Python reference.

    python tests/golden/make_rows_golden.py          (build container only: needs /root/reference)

The reference's keygen(params, seed) re-seeds the generator for every row, so all ell rows of a seeded key are equal.
Its building blocks are general, though: here every row of sk_l / sk_r is drawn by the reference's own
sample_polynomial_coefficient_representation with ITS OWN seed, transformed with the reference's transform(), multiplied with the
reference's GeneralMatrix product (algebra/matrices.py:121-131) and signed with the reference's sign() -- exactly the work
fz_keygen_rows / fz_sign_mac do on the device.  Stored in rows_golden.npz: the row seeds, vk_l, vk_r and the signature of one key per level.
"""
import os
import sys

import numpy as np

REF = os.environ.get("FUSION_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)

from algebra.matrices import GeneralMatrix  # noqa: E402
from algebra.polynomials import sample_polynomial_coefficient_representation, transform  # noqa: E402
import fusion.fusion as ff  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    out = {}
    for secpar, n in ((128, 2), (256, 1)):
        p = ff.fusion_setup(secpar, 42)
        ell = p.num_rows_sk
        rng = np.random.default_rng(7000 + secpar)
        seeds = rng.integers(1, 2**62, size=(n, 2, ell), dtype=np.uint64)
        seeds[0, 0, 0] = 5                      # a one-limb seed and a two-limb seed among the rows
        seeds[0, 1, 1] = 2**32 + 9
        vks_l, vks_r, sigs, msgs = [], [], [], []
        for i in range(n):
            sides = []
            for side in range(2):
                rows = [[transform(sample_polynomial_coefficient_representation(
                    modulus=p.modulus, degree=p.degree, root=p.root, inv_root=p.inv_root, root_order=p.root_order,
                    norm_bound=p.beta_sk, weight_bound=p.omega_sk, seed=int(seeds[i, side, k])))] for k in range(ell)]
                sides.append(GeneralMatrix(matrix=rows))
            sk = ff.OneTimeSigningKey(seed=None, left_sk_hat=sides[0], right_sk_hat=sides[1])
            vk = ff.OneTimeVerificationKey(left_vk_hat=p.public_challenge * sides[0], right_vk_hat=p.public_challenge * sides[1])
            msg = f"rows-{secpar}-{i}"
            sig = ff.sign(p, (sk, vk), msg)
            vks_l.append(vk.left_vk_hat.matrix[0][0].values)
            vks_r.append(vk.right_vk_hat.matrix[0][0].values)
            sigs.append([row[0].values for row in sig.signature_hat.matrix])
            msgs.append(msg)
            # the single-signer aggregate verifies under the reference (sign relation A.z = vk_l*c + vk_r holds for distinct rows)
            agg = ff.aggregate(p, [vk], [msg], [sig])
            assert ff.verify(p, [vk], [msg], agg) == (True, ""), "reference rejects its own general-row signature"
        out[f"row_seeds_{secpar}"] = seeds
        out[f"vk_l_{secpar}"] = np.array(vks_l, dtype=np.int32)
        out[f"vk_r_{secpar}"] = np.array(vks_r, dtype=np.int32)
        out[f"sig_{secpar}"] = np.array(sigs, dtype=np.int32)
        out[f"msgs_{secpar}"] = np.array(msgs)
    np.savez_compressed(os.path.join(OUT, "rows_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
