"""csrc/fz_decoder.cuh -- decode_bytes_to_polynomial_coefficients (fusion.py:422-481) as the byte-driven state machine one thread of decode_kernel
runs -- compiled for the HOST: the reference's decoder cases (all-0 / all-1 bytes as in the reference's tests/test_fusion.py:475-558, random bytes,
general bounds, byte strings that run out in the shuffle -- SURVEY quirk Q4) and a Python restatement of the reference's function on random
parameters.  No GPU needed."""
import ctypes as C
import json
import os
from math import ceil, log2

import numpy as np
import pytest

import hostlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
Q = 2147465729


@pytest.fixture(scope="module", params=["decoder_host", "dec_stream_host"])      # fz_decoder.cuh, and its twin of the one-shot aggregation-hash kernels
def lib(request):
    return hostlib.load(request.param, cuda_headers=request.param == "dec_stream_host")


byte_budget, host_decode = hostlib.byte_budget, hostlib.host_decode


def reference_decode(b: bytes, log2_bias, d, norm_bound, weight):
    """fusion.py:422-481 restated with the same slicing (Python slices past the end give shorter / empty byte strings)."""
    num = max(1, min(d, weight))
    bound, bpc, bpi = byte_budget(log2_bias, d, norm_bound)
    bfs = ceil(weight / 8)
    sg, rest = b[:bfs], b[bfs:]
    bits = bin(int.from_bytes(sg, "big"))[2:].zfill(8 * bfs)[::-1]
    signs = [2 * int(x) - 1 for i, x in enumerate(bits) if i < weight]
    coefs = []
    for i in range(weight):
        nb, rest = rest[:bpc], rest[bpc:]
        coefs.append(((int.from_bytes(nb, "big") % bound) + 1) * signs[i])
    coefs += [0] * (d - len(coefs))
    if num < d:
        for i in range(d - 1, weight, -1):
            nb, rest = rest[:bpi], rest[bpi:]
            j = int.from_bytes(nb, "big") % (i + 1)
            coefs[i], coefs[j] = coefs[j], coefs[i]
    return coefs


def test_reference_decoder_cases(lib):
    cases = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))["decoder"]
    seen = 0
    for case in cases:
        if case["coefs"] is None:          # "Too few bytes to decode polynomial": an argument check of the C ABI (fz_decode_poly returns -3), not the decoder
            continue
        d = 64 if case["secpar"] == 128 else 256
        b = bytes.fromhex(case["bytes_hex"])
        assert host_decode(lib, b, case["secpar"], d, case["bound"], case["weight"]) == case["coefs"], (case["secpar"], case["weight"], case["bound"])
        assert reference_decode(b, case["secpar"], d, case["bound"], case["weight"]) == case["coefs"]      # the restatement used below is the reference's
        seen += 1
    assert seen >= 10


def test_random_parameters_against_the_restated_reference(lib):
    rng = np.random.default_rng(21)
    for trial in range(400):
        d = int(rng.choice([4, 16, 64, 256]))
        log2_bias = int(rng.choice([0, 8, 128, 256]))
        norm_bound = int(rng.choice([1, 2, 3, 52, 255, 256, 65537, Q // 2, Q]))
        weight = int(rng.integers(0, d + 1)) if rng.integers(4) else d       # 0: nothing but the shuffle of an all-zero polynomial
        bound, bpc, bpi = byte_budget(log2_bias, d, norm_bound)
        need = ceil(weight / 8) + bpc * weight + max(0, d - 1 - weight) * bpi
        have = int(rng.choice([need, need + 5, max(ceil(weight / 8) + (bpc + bpi) * weight, 1), int(rng.integers(ceil(weight / 8) + bpc * weight, need + 1))]))
        kind = int(rng.integers(3))
        b = bytes(have) if kind == 0 else (b"\xff" * have if kind == 1 else rng.integers(0, 256, size=have, dtype=np.uint8).tobytes())
        assert host_decode(lib, b, log2_bias, d, norm_bound, weight) == reference_decode(b, log2_bias, d, norm_bound, weight), (trial, d, log2_bias, norm_bound, weight, have)
