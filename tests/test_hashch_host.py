"""csrc/fz_hashch.cuh -- the challenge hash as one sponge thread of hash_ch_stream_kernel runs it -- compiled for the HOST: the coefficients it decodes
equal the ones the unmodified reference decoded from its own XOF bytes (hash_vk_and_int_to_bytes + decode_bytes_to_polynomial_coefficients,
fusion.py:412-481; fixtures of tests/golden/make_golden.py) and the oracle's on random keys, including keys with short, negative, zero and extreme
values whose text takes a different number of blocks.  The squeeze stops at the last byte the decoder reads (n_ch_used), as the kernel does."""
import ctypes as C
import hashlib
import json
import os
from math import ceil, log2

import numpy as np
import pytest

import hostlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
Q = 2147465729
CUDA_INC = "/usr/local/cuda/include"
LEVEL = {128: (64, 27), 256: (256, 60)}      # secpar -> (degree, omega_ch)   (fusion.py:26-27, 76, 102)


@pytest.fixture(scope="module")
def lib():
    return hostlib.load("hashch_host", cuda_headers=True)


def host_hash_ch(lib, secpar, vk_l, vk_r, digests):
    d, omega = LEVEL[secpar]
    bpc = ceil((log2(1) + 1 + secpar) / 8)                       # beta_ch = 1 (fusion.py:433-435)
    bpi = ceil((log2(d) + secpar) / 8)
    n_used = ceil(omega / 8) + bpc * omega + (d - 1 - omega) * bpi     # fusion.py:436-438, 473: the bytes the decoder reads
    vl = np.ascontiguousarray(np.asarray(vk_l, dtype=np.int64).astype(np.int32))
    vr = np.ascontiguousarray(np.asarray(vk_r, dtype=np.int64).astype(np.int32))
    dg = np.ascontiguousarray(np.frombuffer(b"".join(digests), dtype=np.uint8))
    out = np.full(vl.shape, 77, dtype=np.int32)
    rc = lib.h_hash_ch(C.c_int(1 if secpar == 256 else 0), C.c_int(d), C.c_int(omega), C.c_int(bpc), C.c_int(bpi), C.c_uint32(n_used),
                       vl.ctypes.data_as(C.c_void_p), vr.ctypes.data_as(C.c_void_p), dg.ctypes.data_as(C.c_void_p), C.c_size_t(vl.shape[0]),
                       out.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return out.astype(np.int64)


@pytest.mark.parametrize("secpar", [128, 256])
def test_challenge_coefficients_of_the_reference_keys(lib, secpar):
    J = json.load(open(os.path.join(GOLD, "golden.json")))
    Z = np.load(os.path.join(GOLD, "scheme_golden.npz"))
    msgs, pre = J[f"msgs_{secpar}"], [int(x) for x in J[f"prehash_{secpar}"]]
    dst = bytes([3 if secpar == 256 else 1, 0])                  # sign_pre_hash_dst, fusion.py:405-409
    digests = [hashlib.sha3_256(dst + b"," + m.encode()).digest() for m in msgs]
    assert [int.from_bytes(x, "little") for x in digests] == pre
    got = host_hash_ch(lib, secpar, Z[f"vk_l_{secpar}"], Z[f"vk_r_{secpar}"], digests)
    assert np.array_equal(got, Z[f"c_coef_{secpar}"])
    d, omega = LEVEL[secpar]
    assert all(int((row != 0).sum()) == omega and set(np.unique(row)) <= {-1, 0, 1} for row in got)


@pytest.mark.parametrize("secpar", [128, 256])
def test_random_and_edge_keys_against_the_oracle(lib, secpar):
    from oracle import oracle as orc

    d, _ = LEVEL[secpar]
    rng = np.random.default_rng(secpar + 1)
    h = (Q - 1) // 2
    n = 48
    vk_l = rng.integers(-h, h + 1, size=(n, d)).astype(np.int64)
    vk_r = rng.integers(-h, h + 1, size=(n, d)).astype(np.int64)
    vk_l[0] = 0; vk_r[0] = 0                                   # the shortest text
    vk_l[1] = -h; vk_r[1] = -h                                 # the longest
    vk_l[2] = h; vk_r[2] = 1
    vk_l[3] = rng.integers(-9, 10, size=d); vk_r[3] = rng.integers(-1000, 1000, size=d)
    vk_l[4, ::2] = 0; vk_r[4, 1::2] = -1
    for k in range(5, 12):                                       # text lengths spread over several block counts
        lim = 10 ** (k - 3)
        vk_l[k] = rng.integers(-lim, lim, size=d); vk_r[k] = rng.integers(-lim, lim, size=d)
    msgs = [f"host-msg-{i}".encode() for i in range(n)]
    digests = [orc.hash_msg(secpar, m) for m in msgs]
    digests[6] = bytes(32)                                       # prehash 0 -> "0"
    digests[7] = b"\xff" * 32                                    # the longest decimal (78 digits)
    got = host_hash_ch(lib, secpar, vk_l, vk_r, digests)
    omega = LEVEL[secpar][1]
    n_ch = ceil(omega / 8) + ceil((1 + secpar) / 8) * omega + d * ceil((log2(d) + secpar) / 8)     # fusion.py:520-524: what the reference squeezes
    for i in range(n):
        xof = orc.hash_vk_and_int_to_bytes(secpar, vk_l[i].astype(np.int32), vk_r[i].astype(np.int32), digests[i], n_ch)
        want = orc.decode_poly(secpar, xof, secpar, 1, LEVEL[secpar][1])
        assert np.array_equal(got[i], np.asarray(want, dtype=np.int64)), i
