"""csrc/fz_mt.cuh -- MT19937 seeded the way CPython's `random.seed(int)` seeds it, `randrange` as Lib/random.py draws it, and the per-thread draw loops of
both sampler kernels -- compiled for the HOST and checked against CPython's `random` itself, restating algebra/polynomials.py:436-488 (the secret-key
and public-matrix samplers of the reference).  Seeds of one, two and more 32-bit limbs take the three seeding routines of the kernels.  No GPU needed."""
import ctypes as C
import os
import random

import numpy as np
import pytest

import hostlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
Q = 2147465729
CUDA_INC = "/usr/local/cuda/include"


@pytest.fixture(scope="module")
def lib():
    return hostlib.load("sampler_host", cuda_headers=True)


def limbs(seeds):
    """random.seed(n): key = little-endian 32-bit limbs of abs(n), at least one (Modules/_randommodule.c)."""
    keys = np.zeros((len(seeds), 8), dtype=np.uint32)
    klen = np.zeros(len(seeds), dtype=np.int32)
    for i, s in enumerate(seeds):
        a, k = abs(int(s)), 0
        while True:
            keys[i, k] = a & 0xFFFFFFFF
            a >>= 32
            k += 1
            if a == 0:
                break
        klen[i] = k
    return keys, klen


def ref_coef(seed, d, bound, weight):       # algebra/polynomials.py:446-459
    random.seed(seed)
    n = max(0, min(d, weight))
    c = [(1 + random.randrange(bound)) * (1 - 2 * random.randrange(2)) for _ in range(n)]
    c += [0] * (d - n)
    if n < d:
        for i in range(d - 1, 0, -1):
            j = random.randrange(i + 1)
            c[i], c[j] = c[j], c[i]
    return c


def ref_ntt(seed, d, q):                    # algebra/polynomials.py:478-480
    random.seed(seed)
    return [random.randrange(q) - q // 2 for _ in range(d)]


SEEDS = ([0, 1, 2, 42, 43, 10**6, 10**6 + 1, 2**32 - 1, 2**32, 2**32 + 1, 2**63, 2**64 - 1, -7, -(2**40)]
         + [2**64, 2**64 + 5, 2**95 + 12345, 2**200 + 99, 3**150]                       # more than two limbs: the per-thread table
         + [int(x) for x in np.random.default_rng(5).integers(0, 2**62, size=120)])


def general(lib, mode, seeds, d, q, bound, weight):
    keys, klen = limbs(seeds)
    out = np.full((len(seeds), d), 12345, dtype=np.int32)
    lib.h_sample(C.c_int(mode), keys.ctypes.data_as(C.c_void_p), C.c_int(8), klen.ctypes.data_as(C.c_void_p), C.c_size_t(len(seeds)), C.c_int(d),
                 C.c_uint32(q), C.c_uint32(bound), C.c_int(weight), out.ctypes.data_as(C.c_void_p))
    return out


def dense(lib, stage, seeds, d, bound):
    keys, klen = limbs(seeds)
    out = np.full((len(seeds), d), 12345, dtype=np.int32)       # numpy's allocations are 16-byte aligned, rows are multiples of 16 bytes
    lib.h_sample_dense(C.c_int(stage), keys.ctypes.data_as(C.c_void_p), C.c_int(8), klen.ctypes.data_as(C.c_void_p), C.c_size_t(len(seeds)), C.c_int(d),
                       C.c_uint32(bound), out.ctypes.data_as(C.c_void_p))
    return out


@pytest.mark.parametrize("d,bound", [(256, 52), (64, 52), (256, 1), (256, 127), (64, 3), (16, 100)])
def test_dense_secret_key_sampler_both_kernels(lib, d, bound):
    want = np.array([ref_coef(s, d, bound, d) for s in SEEDS], dtype=np.int64)
    assert np.array_equal(general(lib, 0, SEEDS, d, Q, bound, d), want)        # sample_kernel, dense branch
    assert np.array_equal(dense(lib, 1, SEEDS, d, bound), want)                # sample_dense_kernel<1>: byte tile
    assert np.array_equal(dense(lib, 0, SEEDS, d, bound), want)                # sample_dense_kernel<0>: 16-byte stores


@pytest.mark.parametrize("d,bound,weight", [(256, 52, 60), (64, 1, 27), (64, 52, 0), (256, 7, 255), (16, 1000, 1)])
def test_sparse_secret_key_sampler_with_the_fisher_yates_branch(lib, d, bound, weight):
    want = np.array([ref_coef(s, d, bound, weight) for s in SEEDS[:60]], dtype=np.int64)
    assert np.array_equal(general(lib, 0, SEEDS[:60], d, Q, bound, weight), want)


@pytest.mark.parametrize("d,q", [(256, Q), (64, Q), (64, 12289), (16, 7681), (8, 17)])
def test_public_matrix_sampler(lib, d, q):
    want = np.array([ref_ntt(s, d, q) for s in SEEDS], dtype=np.int64)
    assert np.array_equal(general(lib, 1, SEEDS, d, q, 0, 0), want)


def test_more_generators_than_one_cta_holds(lib):
    """Columns are reused as a CTA's threads reuse them: 200 generators over the 84 / 92 / 90 state columns of the three kernels."""
    seeds = [10**6 + 2 * i for i in range(200)]      # the seeds of BASELINE configs[2]
    want = np.array([ref_coef(s, 256, 52, 256) for s in seeds], dtype=np.int64)
    for got in (general(lib, 0, seeds, 256, Q, 52, 256), dense(lib, 1, seeds, 256, 52), dense(lib, 0, seeds, 256, 52)):
        assert np.array_equal(got, want)
