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


def get_ring(q, d, fwd=None, inv=None, device=None):
    return Ring(q, d, fwd, inv)


def get_context(secpar, device=None):
    tw, itw = orc.twiddles(secpar)
    r = Ring(Q, len(tw), [int(x) for x in tw], [int(x) for x in itw])
    r.ell = 195 if secpar == 128 else 83
    return r
