"""The product's arithmetic headers (csrc/fz_common.cuh: Shoup / Plantard / Montgomery products, lazy and centred reductions, the 64-bit
accumulator folds; csrc/fz_keccak.cuh: Keccak-f[1600]) compiled for the HOST and checked against Python integers and hashlib -- the identities
every kernel relies on (reference: cent, algebra/ntt.py:93-123; hashlib.sha3_256 / shake_256, fusion/fusion.py:408, 419, 591), pinned
without a GPU.  tests/host/modarith_host.cpp only supplies the device intrinsics (__umulhi, __mulhi, __funnelshift_l, min)."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

import hostlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
Q = 2147465729
HALF = (Q - 1) // 2
CUDA_INC = "/usr/local/cuda/include"


def cent(v: int) -> int:
    y = v % Q
    return y - Q if y > HALF else y


@pytest.fixture(scope="module")
def lib():
    lib = hostlib.load("modarith_host", cuda_headers=True)
    lib.h_q.restype = C.c_uint32
    assert lib.h_q() == Q
    return lib


def op32(lib, op, a, b=None, c=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.int64).astype(np.uint32))
    n = len(a)
    b = np.ascontiguousarray(np.asarray(b if b is not None else np.zeros(n), dtype=np.int64).astype(np.uint32))
    c = np.ascontiguousarray(np.asarray(c if c is not None else np.zeros(n), dtype=np.int64).astype(np.uint32))
    out = np.empty(n, dtype=np.int64)
    lib.h_op32(C.c_int(op), a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
               C.c_size_t(n))
    return [int(x) for x in out]


def samples(rng, n, lo, hi, edges):
    return [int(e) for e in edges if lo <= e < hi] + [int(x) for x in rng.integers(lo, hi, size=n)]


def pair_up(xs, ys):
    n = max(len(xs), len(ys))
    return [xs[i % len(xs)] for i in range(n)], [ys[(7 * i) % len(ys)] for i in range(n)]


def test_shoup_product_any_32_bit_operand(lib):
    rng = np.random.default_rng(1)
    a = samples(rng, 60000, 0, 1 << 32, [0, 1, Q - 1, Q, Q + 1, 2 * Q - 1, 2 * Q, 2 * Q + 1, (1 << 32) - 1, 1 << 31])
    w = samples(rng, 5000, 0, Q, [0, 1, 2, Q - 1, Q - 2, 3337519, 1978410468, 23584283, 540632852, HALF, HALF + 1])
    a, w = pair_up(a, w)
    wq = [(x << 32) // Q for x in w]
    got = op32(lib, 0, a, w, wq)
    for x, y, r in zip(a, w, got):
        assert 0 <= r < 2 * Q and r % Q == x * y % Q, (x, y, r)


def test_plantard_products(lib):
    rng = np.random.default_rng(2)
    a = samples(rng, 60000, 0, 1 << 32, [0, 1, Q - 1, Q, Q + 1, 2 * Q, (1 << 32) - 1, 1 << 31])
    w = samples(rng, 5000, 0, Q, [0, 1, Q - 1, 2139077191, 2113911577, HALF, HALF + 1])   # among them n^-1 of both levels
    a, w = pair_up(a, w)
    qinv64 = pow(Q, -1, 1 << 64)
    W = [(((x * ((-(1 << 64)) % Q)) % Q) * qinv64) % (1 << 64) for x in w]
    lo, hi = [x & 0xFFFFFFFF for x in W], [x >> 32 for x in W]
    for x, y, r in zip(a, w, op32(lib, 1, a, lo, hi)):
        assert 0 <= r <= Q and r % Q == x * y % Q, (x, y, r)
    for x, y, r in zip(a, w, op32(lib, 2, a, lo, hi)):
        assert r == cent(x * y), (x, y, r)


def test_reductions(lib):
    rng = np.random.default_rng(3)
    s = samples(rng, 50000, 0, 2 * Q + 1, [0, 1, HALF, HALF + 1, Q - 1, Q, Q + 1, Q + HALF, Q + HALF + 1, 2 * Q - 1, 2 * Q])
    for x, r in zip(s, op32(lib, 3, s)):
        assert 0 <= r <= Q and r % Q == x % Q, (x, r)
    for x, r in zip(s, op32(lib, 7, s)):
        assert r == cent(x), (x, r)
    # any int32 is a legal input of the transforms
    v = samples(rng, 50000, -(1 << 31), 1 << 31, [0, 1, -1, HALF, -HALF, HALF + 1, -HALF - 1, Q - 1, -(Q - 1), (1 << 31) - 1, -(1 << 31)])
    for x, r in zip(v, op32(lib, 5, v)):
        assert 0 <= r < 2 * Q and r % Q == x % Q, (x, r)
    for x, r in zip(v, op32(lib, 6, v)):
        assert 0 <= r <= Q and r % Q == x % Q, (x, r)


def test_montgomery_and_centred_products(lib):
    rng = np.random.default_rng(4)
    yr = samples(rng, 60000, 0, Q + 1, [0, 1, HALF, HALF + 1, Q - 1, Q])
    b = samples(rng, 6000, -(1 << 31), 1 << 31, [0, 1, -1, HALF, -HALF, (1 << 31) - 1, -(1 << 31), Q - 1, -(Q - 1)])
    yr, b = pair_up(yr, b)
    r32 = pow(1 << 32, -1, Q)
    for x, y, r in zip(yr, b, op32(lib, 4, yr, b)):
        assert 0 <= r <= Q and r % Q == x * y * r32 % Q, (x, y, r)
    a = samples(rng, 60000, -(1 << 31), 1 << 31, [0, 1, -1, HALF, -HALF, (1 << 31) - 1, -(1 << 31)])
    a, b2 = pair_up(a, b)
    for x, y, r in zip(a, b2, op32(lib, 8, a, b2)):
        assert r == cent(x * y), (x, y, r)


def test_accumulator_folds(lib):
    rng = np.random.default_rng(5)
    lim = (1 << 63) - 1
    acc = [0, 1, -1, Q, -Q, HALF, HALF + 1, -HALF - 1, lim, -lim, 1 << 62, -(1 << 62), (1 << 31), (1 << 31) - 1]
    acc += [int(x) for x in rng.integers(-lim, lim, size=50000)]
    a = np.asarray(acc, dtype=np.int64)
    out = np.empty_like(a)
    lib.h_op64(C.c_int(0), a.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.c_size_t(len(a)))
    for x, r in zip(acc, out):
        assert int(r) == cent(x), (x, int(r))
    lib.h_op64(C.c_int(1), a.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.c_size_t(len(a)))
    for x, r in zip(acc, out):
        assert int(r) % Q == x % Q and abs(int(r)) < 1 << 47, (x, int(r))


def test_keccak_f_against_hashlib(lib):
    rng = np.random.default_rng(6)

    def sponge(msg: bytes, suffix: int, n: int) -> bytes:
        out = C.create_string_buffer(max(n, 1))
        lib.h_sponge(msg, C.c_size_t(len(msg)), C.c_uint8(suffix), out, C.c_size_t(n))
        return out.raw[:n]

    lengths = list(range(0, 280)) + [407, 408, 409, 1000, 6439, 9513, 40000]   # around the block boundaries, and the sizes of the scheme's hash inputs
    for L in lengths:
        msg = rng.integers(0, 256, size=L, dtype=np.uint8).tobytes()
        assert sponge(msg, 0x06, 32) == hashlib.sha3_256(msg).digest(), L
        for n in (1, 135, 136, 137, 3968, 10436):       # n_ag and n_ch at secpar 256 among them
            assert sponge(msg, 0x1F, n) == hashlib.shake_256(msg).digest(n), (L, n)
