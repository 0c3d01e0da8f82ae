"""csrc/fz_mac.cuh and the accumulator folds of csrc/fz_common.cuh under the accumulation cadences of the streaming kernels (matvec / sign_mac: two
products then a fold; aggregate: a pair of products added, then a fold), compiled for the HOST: cent() of the exact integer sum for the extreme
operands the contract allows -- the matrix side canonical (|a| <= (q-1)/2), the streamed side ANY int32 -- over ell = 83 / 195 rows and a
capacity-sized aggregate.  A signed 64-bit overflow anywhere would change the result modulo q.  (algebra/matrices.py:108-131, 74-91.)"""
import ctypes as C

import numpy as np
import pytest

import hostlib

Q = hostlib.Q
H = (Q - 1) // 2


@pytest.fixture(scope="module")
def lib():
    lib = hostlib.load("mac_host", cuda_headers=True)
    lib.h_aggregate1.restype = C.c_int32
    return lib


def cent(v):
    y = v % Q
    return y - Q if y > H else y


def patterns(rng, n):
    """(matrix side, streamed side) rows of n values: all extreme combinations, alternating signs, random."""
    lo, hi = -(1 << 31), (1 << 31) - 1
    yield np.full(n, H), np.full(n, hi)
    yield np.full(n, H), np.full(n, lo)
    yield np.full(n, -H), np.full(n, lo)          # every product +2^61: the largest positive run
    yield np.full(n, -H), np.full(n, hi)
    yield np.where(np.arange(n) % 2 == 0, H, -H), np.where(np.arange(n) % 2 == 0, lo, hi)
    yield np.where(np.arange(n) % 4 < 2, H, -H), np.full(n, lo)
    for _ in range(20):
        yield rng.integers(-H, H + 1, size=n), rng.integers(lo, hi + 1, size=n)


@pytest.mark.parametrize("ell", [1, 2, 3, 4, 5, 83, 195])
def test_matvec_cadence(lib, ell):
    rng = np.random.default_rng(ell)
    for a, s in patterns(rng, ell):
        A = np.ascontiguousarray(np.repeat(a[:, None], 4, axis=1).astype(np.int32)); S = np.ascontiguousarray(np.repeat(s[:, None], 4, axis=1).astype(np.int32))
        S[:, 1] = S[::-1, 0]; A[:, 2] = A[::-1, 0]                 # four different columns
        out = np.empty(4, dtype=np.int32)
        lib.h_matvec4(hostlib.ptr(A), hostlib.ptr(S), C.c_int(ell), hostlib.ptr(out))
        want = [cent(sum(int(x) * int(y) for x, y in zip(A[:, k], S[:, k]))) for k in range(4)]
        assert out.tolist() == want, ell


@pytest.mark.parametrize("n", [1, 2, 3, 4, 7, 1796, 2818])
def test_aggregate_cadence(lib, n):
    rng = np.random.default_rng(n)
    for a, z in patterns(rng, n):
        zz, aa = hostlib.i32(z), hostlib.i32(a)
        got = lib.h_aggregate1(hostlib.ptr(zz), hostlib.ptr(aa), C.c_int(n))
        assert got == cent(sum(int(x) * int(y) for x, y in zip(zz, aa))), n
