"""csrc/fz_ntt_core.cuh -- the tables, butterflies and stage functions of the scheme's NTT kernels -- compiled for the HOST: whole forward / inverse
transforms and the fused pass run through the product's own arithmetic (one emulated lane after the other, tests/host/ntt_host.cpp) and are
compared with the vectors recorded from the reference's cooley_tukey_ntt / gentleman_sande_intt / transform (algebra/ntt.py:216-377,
algebra/polynomials.py:391-433; tests/golden/make_golden.py) and with the oracle on random and arbitrary-int32 inputs.  No GPU needed."""
import ctypes as C
import os

import numpy as np
import pytest

import hostlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
Q = 2147465729
CUDA_INC = "/usr/local/cuda/include"


@pytest.fixture(scope="module")
def lib():
    return hostlib.load("ntt_host", cuda_headers=True)


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "ntt_golden.npz"))


def run(lib, secpar, mode, a, b=None):
    logd = 8 if secpar == 256 else 6
    a = np.ascontiguousarray(np.asarray(a, dtype=np.int64).astype(np.int32))
    b = np.ascontiguousarray(np.asarray(b if b is not None else np.zeros_like(a), dtype=np.int64).astype(np.int32))
    out = np.empty_like(a)
    assert lib.h_ntt(C.c_int(logd), C.c_int(mode), a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
                     C.c_size_t(a.shape[0])) == 0
    return out.astype(np.int64)


@pytest.mark.parametrize("secpar", [128, 256])
def test_tables_are_the_reference_tables_as_shoup_pairs(lib, gold, secpar):
    d = 64 if secpar == 128 else 256
    arrs = [np.empty(d, dtype=np.uint32) for _ in range(4)]
    lib.h_tables(C.c_int(6 if secpar == 128 else 8), *[x.ctypes.data_as(C.c_void_p) for x in arrs])
    fw, fwq, iw, iwq = [x.astype(np.int64) for x in arrs]
    assert np.array_equal(fw, gold[f"tw_{secpar}"]) and np.array_equal(iw, gold[f"itw_{secpar}"])       # bit_reverse_copy of the root powers, in [0, q)
    assert [int(x) for x in fwq] == [(int(w) << 32) // Q for w in fw] and [int(x) for x in iwq] == [(int(w) << 32) // Q for w in iw]


@pytest.mark.parametrize("secpar", [128, 256])
def test_transforms_equal_the_reference_vectors(lib, gold, secpar):
    a = gold[f"a_{secpar}"]          # among them: zeros, +-(q-1)/2, deltas, arbitrary int32, INT32_MAX / INT32_MIN rows
    assert np.array_equal(run(lib, secpar, 0, a), gold[f"fwd_{secpar}"])
    assert np.array_equal(run(lib, secpar, 1, a), gold[f"inv_{secpar}"])
    assert np.array_equal(run(lib, secpar, 1, gold[f"fwd_{secpar}"]), (a + Q // 2) % Q - Q // 2)    # INTT(NTT(a)) = cent(a)


@pytest.mark.parametrize("secpar", [128, 256])
def test_fused_pass_equals_the_reference_convolution(lib, gold, secpar):
    a, b = gold[f"a_{secpar}"][8:12], gold[f"b_{secpar}"][8:12]
    b_hat = run(lib, secpar, 0, b)
    assert np.array_equal(run(lib, secpar, 2, a, b_hat), gold[f"conv_{secpar}"])     # transform(transform(x) * transform(y)) of the reference


@pytest.mark.parametrize("secpar", [128, 256])
def test_random_and_arbitrary_int32_inputs_against_the_oracle(lib, secpar):
    from oracle import oracle as orc

    d = 64 if secpar == 128 else 256
    rng = np.random.default_rng(secpar)
    h = (Q - 1) // 2
    a = rng.integers(-h, h + 1, size=(300, d)).astype(np.int64)
    a[:60] = rng.integers(-(1 << 31), 1 << 31, size=(60, d))       # any int32 is a legal input
    b = rng.integers(-(1 << 31), 1 << 31, size=(300, d)).astype(np.int64)
    want_f = orc.ntt_fwd(secpar, a.astype(np.int32))
    assert np.array_equal(run(lib, secpar, 0, a), want_f)
    assert np.array_equal(run(lib, secpar, 1, a), orc.ntt_inv(secpar, a.astype(np.int32)))
    want_p = orc.ntt_inv(secpar, orc.pointwise_mul(want_f, b.astype(np.int32)))
    assert np.array_equal(run(lib, secpar, 2, a, b), want_p)
