"""TEST INFRASTRUCTURE ONLY -- a CPU stand-in for ``fusion_cryptography_b200.native`` built on the oracle.

``tests/test_reference_suite_cpu.py`` puts this module on ``sys.path`` under the name ``native`` (in a scratch directory, next to a
link to the product's ``algebra`` package) so that the reference's OWN unit tests of its algebra layer can be run against the HOST MIRROR
(``algebra/ntt.py``, ``algebra/polynomials.py``, ``algebra/matrices.py``: classes, operators, validation, exceptions, ``str()``) in a
container that has the reference checkout but no GPU.  The arithmetic the mirror asks the device for is answered by the oracle's
general-ring functions -- the same functions the CUDA kernels are compared with bit for bit in ``tests/test_gpu_ring.py``.  Nothing in the
product imports this file; on a GPU box the test is skipped (there is no reference checkout there) and the kernels are tested directly.
"""
import os
import random
import sys

import numpy as np
import torch

sys.path.insert(0, os.environ["FUSION_B200_REPO"])
from oracle import oracle as orc  # noqa: E402

Q = 2147465729


class FusionB200Error(RuntimeError):
    pass


def _np(x):
    return (x.cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)).astype(np.int32)


class Ring:
    def __init__(self, q, d, fwd=None, inv=None):
        for t in (fwd, inv):
            if t is not None and len(t) < d:
                raise IndexError("list index out of range")   # as native.Ring does (the reference's table[m + i])
        self.q, self.d, self.fwd, self.inv, self.device = q, d, fwd, inv, torch.device("cpu")

    @staticmethod
    def _t(a, shape):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32).reshape(shape))

    def ntt_fwd(self, x, out=None):
        a = _np(x)
        return self._t(orc.ring_ntt_fwd(self.q, [t % self.q for t in self.fwd[:self.d]], a.reshape(-1, self.d)), a.shape)

    def ntt_inv(self, x, out=None):
        a = _np(x)
        return self._t(orc.ring_ntt_inv(self.q, [t % self.q for t in self.inv[:self.d]], a.reshape(-1, self.d)), a.shape)

    def pointwise_mul(self, a, b, out=None):
        a, b = _np(a), _np(b)
        return self._t(orc.ring_pointwise(self.q, a, b, True), a.shape)

    def pointwise_add(self, a, b, out=None):
        a, b = _np(a), _np(b)
        return self._t(orc.ring_pointwise(self.q, a, b, False), a.shape)

    def polymul(self, a, b, out=None):
        return self.ntt_inv(self.pointwise_mul(self.ntt_fwd(a), self.ntt_fwd(b)))

    def sample_coef(self, seeds, norm_bound=52, weight_bound=None):
        rows = []
        for s in seeds:  # algebra/polynomials.py:447-459 on CPython's own generator
            rnd = random.Random()
            rnd.seed(s)
            num = max(0, min(self.d, self.d if weight_bound is None else weight_bound))
            c = [(1 + rnd.randrange(norm_bound)) * (1 - 2 * rnd.randrange(2)) for _ in range(num)] + [0] * (self.d - num)
            if num < self.d:
                for i in range(self.d - 1, 0, -1):
                    j = rnd.randrange(i + 1)
                    c[i], c[j] = c[j], c[i]
            rows.append(c)
        return torch.tensor(rows, dtype=torch.int32)

    def sample_ntt(self, seeds):
        rows = []
        for s in seeds:
            rnd = random.Random()
            rnd.seed(s)
            rows.append([rnd.randrange(self.q) - self.q // 2 for _ in range(self.d)])
        return torch.tensor(rows, dtype=torch.int32)


    def decode_poly(self, data, blen, norm_bound, weight_bound, bytes_per_coefficient, bytes_per_index):
        """fusion/fusion.py:443-481 restated on Python integers for any degree (the length check is made by the caller, as by fz_ring_decode_poly)"""
        rows = []
        bound = max(1, min(self.q // 2, norm_bound))
        for row in data.numpy():
            b = bytes(row[:blen])
            bfs = (weight_bound + 7) // 8
            sg = int.from_bytes(b[:bfs], "big")
            rest = b[bfs:]
            c = []
            for i in range(weight_bound):
                nb, rest = rest[:bytes_per_coefficient], rest[bytes_per_coefficient:]
                c.append(((int.from_bytes(nb, "big") % bound) + 1) * (2 * ((sg >> i) & 1) - 1))
            c += [0] * (self.d - len(c))
            if max(1, min(self.d, weight_bound)) < self.d:
                for i in range(self.d - 1, weight_bound, -1):
                    nb, rest = rest[:bytes_per_index], rest[bytes_per_index:]
                    j = int.from_bytes(nb, "big") % (i + 1)
                    c[i], c[j] = c[j], c[i]
            rows.append(c)
        return torch.tensor(rows, dtype=torch.int32)


class Context(Ring):
    """the scheme-level device calls fusion/fusion.py's mirror makes, answered by the oracle's restatement of the same steps"""

    def __init__(self, secpar):
        tw, itw = orc.twiddles(secpar)
        super().__init__(Q, len(tw), [int(x) for x in tw], [int(x) for x in itw])
        self.secpar, self.ell = secpar, orc.param(secpar, "ell")

    def _cent(self, a):
        h = Q // 2
        return ((a + h) % Q - h).astype(np.int32)

    def matvec(self, A, s):
        return torch.from_numpy(orc.matvec(self.secpar, _np(A), _np(s)))

    def keygen(self, A, seeds, want_sk=True):
        sk_l, sk_r, vl, vr = orc.keygen_batch(self.secpar, _np(A), np.asarray(seeds, dtype=np.uint64))
        return torch.from_numpy(sk_l), torch.from_numpy(sk_r), None, torch.from_numpy(vl), torch.from_numpy(vr)

    def sha3_256(self, msgs):
        import hashlib

        return torch.tensor([list(hashlib.sha3_256(m).digest()) for m in msgs], dtype=torch.uint8)

    def shake256(self, msgs, n):
        import hashlib

        return torch.tensor([list(hashlib.shake_256(m).digest(n)) for m in msgs], dtype=torch.uint8)

    def hash_msg(self, msgs):
        return torch.tensor([list(orc.hash_msg(self.secpar, m)) for m in msgs], dtype=torch.uint8).reshape(len(msgs), 32)

    def _challenge(self, vl, vr, dg):
        P = lambda name: orc.param(self.secpar, name)  # noqa: E731
        n_ch = (P("omega_ch") + 7) // 8 + P("bytes_coef_ch") * P("omega_ch") + self.d * P("bytes_idx")   # fusion.py:520-524
        b = orc.hash_vk_and_int_to_bytes(self.secpar, vl, vr, bytes(dg), n_ch)
        return orc.ntt_fwd(self.secpar, orc.decode_poly(self.secpar, b, self.secpar, 1, P("omega_ch")).reshape(1, self.d))[0]

    def hash_ch(self, vl, vr, dg):
        vl, vr, dg = _np(vl).reshape(-1, self.d), _np(vr).reshape(-1, self.d), dg.numpy().astype(np.uint8)
        return torch.from_numpy(np.stack([self._challenge(vl[i], vr[i], dg[i]) for i in range(len(vl))]))

    def sign_mac(self, sk_l, sk_r, c_hat):
        sk_l, sk_r, c = _np(sk_l).astype(np.int64), _np(sk_r).astype(np.int64), _np(c_hat).astype(np.int64)
        return torch.from_numpy(self._cent(sk_l * c.reshape(-1, 1, self.d) + sk_r))

    def sort_keys(self, vl, vr, counts, sel=None):
        return torch.from_numpy(orc.sort_by_repr(self.secpar, _np(vl), _np(vr)).astype(np.int32))

    def gather_sorted(self, vl, vr, dg, perm, sel=None):
        idx = perm.long()
        return vl[idx], vr[idx], dg[idx]

    def hash_ag_fast(self, vl_s, vr_s, dg_s, c_s, counts):
        """fusion.py:573-652 for one aggregate whose signers are given in hashing order"""
        import hashlib

        vl, vr, dg, ch = _np(vl_s), _np(vr_s), dg_s.numpy().astype(np.uint8), _np(c_s)
        n, dst = len(vl), (1 if self.secpar == 128 else 3)
        pieces = ["(" + orc.format_vk(self.secpar, vl[i], vr[i]) + ", " + orc.format_prehash(bytes(dg[i])) + ", " + orc.format_ch(self.secpar, ch[i]) + ")"
                  for i in range(n)]
        msg = (bytes([dst, 2]).decode() + "," + "[" + ", ".join(pieces) + "]").encode()
        P = lambda name: orc.param(self.secpar, name)  # noqa: E731
        omega = P("omega_ag")
        n_ag = (omega + 7) // 8 + (P("bytes_coef_ag") + P("bytes_idx")) * omega   # fusion.py:583-585
        xof = hashlib.shake_256(msg).digest(n * n_ag)
        coef = np.stack([orc.decode_poly(self.secpar, xof[i * n_ag:(i + 1) * n_ag], self.secpar, 1, omega).reshape(self.d) for i in range(n)])
        return torch.from_numpy(orc.ntt_fwd(self.secpar, coef))

    def aggregate_sum(self, sigs, perm, alpha, counts):
        z, a = _np(sigs).astype(np.int64)[perm.long().numpy()], _np(alpha).astype(np.int64)
        acc = np.zeros((self.ell, self.d), dtype=np.int64)
        for i in range(len(a)):
            acc = (acc + self._cent(z[i] * a[i]).astype(np.int64)) % Q
        return torch.from_numpy(self._cent(acc).reshape(1, self.ell, self.d))

    def verify_check(self, A, vl_s, vr_s, c_s, alpha, agg, counts):
        """fusion.py:706-727: 3 = target mismatch, 4 = norm, 5 = weight, 0 = accept"""
        vl, vr, c, al = (_np(t).astype(np.int64) for t in (vl_s, vr_s, c_s, alpha))
        target = np.zeros(self.d, dtype=np.int64)
        for i in range(len(vl)):
            t = self._cent(self._cent(vl[i] * c[i]).astype(np.int64) + vr[i]).astype(np.int64)
            target = (target + self._cent(t * al[i]).astype(np.int64)) % Q
        observed = orc.matvec(self.secpar, _np(A), _np(agg).reshape(1, self.ell, self.d))[0]
        if not np.array_equal(self._cent(target), observed):
            return torch.tensor([3], dtype=torch.uint8)
        coef = orc.ntt_inv(self.secpar, _np(agg).reshape(self.ell, self.d))
        if int(np.abs(coef.astype(np.int64)).max()) > orc.param(self.secpar, "beta_vf"):
            return torch.tensor([4], dtype=torch.uint8)
        if int((coef != 0).sum(axis=1).max()) > self.d:
            return torch.tensor([5], dtype=torch.uint8)
        return torch.tensor([0], dtype=torch.uint8)

    def decode_poly(self, data, blen, log2_bias, norm_bound, weight_bound):
        from math import ceil, log2

        bound = max(1, min(Q // 2, norm_bound))
        return Ring.decode_poly(self, data, blen, norm_bound, weight_bound, ceil((log2(bound) + 1 + log2_bias) / 8), ceil((log2(self.d) + log2_bias) / 8))


def get_ring(q, d, fwd=None, inv=None, device=None):
    return Ring(q, d, fwd, inv)


_ctx = {}


def get_context(secpar, device=None):
    if secpar not in _ctx:
        _ctx[secpar] = Context(secpar)
    return _ctx[secpar]
