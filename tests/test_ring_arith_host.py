"""csrc/fz_ring_arith.cuh -- the arithmetic of the general-ring kernels (any odd prime modulus below 2^31, inputs any int32, outputs the centred
residues of the reference's cent(), algebra/ntt.py:93-123; pointwise `*` / `+` of algebra/polynomials.py:341-385, 282-318) -- compiled for the HOST
and checked against Python integers on moduli from 3 to 2^31 - 1, among them the ones the reference's own tests use and the scheme's."""
import ctypes as C

import numpy as np
import pytest

import hostlib

MODULI = [3, 5, 17, 97, 257, 7681, 12289, 65537, 786433, 2013265921, 2147465729, 2147483647, 2147483629, 1073741827]


@pytest.fixture(scope="module")
def lib():
    return hostlib.load("ring_arith_host", cuda_headers=True)


def cent(v, q):
    y = v % q
    return y - q if y > q // 2 else y


def run(lib, op, q, a, b=None):
    a = hostlib.i32(a); b = hostlib.i32(b if b is not None else np.zeros(len(a)))
    out = np.empty(len(a), dtype=np.int64)
    lib.h_ring_op(C.c_int(op), C.c_uint32(q), hostlib.ptr(a), hostlib.ptr(b), hostlib.ptr(out), C.c_size_t(len(a)))
    return [int(x) for x in out]


@pytest.mark.parametrize("q", MODULI)
def test_general_ring_arithmetic(lib, q):
    rng = np.random.default_rng(q % 1000003)
    edges = [0, 1, -1, q - 1, -(q - 1), q // 2, q // 2 + 1, -(q // 2), -(q // 2) - 1, (1 << 31) - 1, -(1 << 31), q if q < (1 << 31) else 0, -q]
    edges = [e for e in edges if -(1 << 31) <= e < (1 << 31)]
    a = edges + [int(x) for x in rng.integers(-(1 << 31), 1 << 31, size=20000)]
    b = [edges[(3 * i) % len(edges)] for i in range(len(edges))] + [int(x) for x in rng.integers(-(1 << 31), 1 << 31, size=20000)]
    for x, r in zip(a, run(lib, 0, q, a)):
        assert r == x % q, (q, x, r)                                 # Python's %
    for x, y, r in zip(a, b, run(lib, 1, q, a, b)):
        assert r == cent(x * y, q), (q, x, y, r)
    for x, y, r in zip(a, b, run(lib, 2, q, a, b)):
        assert r == cent(x + y, q), (q, x, y, r)
    for x, y, r in zip(a, b, run(lib, 3, q, a, b)):                  # Shoup product: any 32-bit operand (taken as unsigned) times w = y mod q
        assert r == ((x % (1 << 32)) * ((y % (1 << 32)) % q)) % q, (q, x, y, r)
    res = [int(x) for x in rng.integers(0, q, size=5000)] + [0, q - 1, q // 2, q // 2 + 1]
    res = [r if r < (1 << 31) else r - (1 << 32) for r in res]
    for x, r in zip(res, run(lib, 4, q, res)):
        assert r == cent(x % (1 << 32), q)
