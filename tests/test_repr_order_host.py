"""csrc/fz_repr_order.cuh -- the order of verification keys by Python's str() of their repr, computed from the coefficients alone -- compiled for the
HOST and compared with sorted(..., key=str) on the REAL repr strings (fusion.py:661-663, 693; the template is the reference's own str(vk) from
tests/golden).  Crafted ties: equal leading coefficients, numbers that are prefixes of one another (1 / 10 / 100: the terminator decides), signs,
the last coefficient of a polynomial (']' instead of ','), keys equal on the left half, fully equal keys (stable)."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

import hostlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
Q = 2147465729


@pytest.fixture(scope="module")
def lib():
    return hostlib.load("repr_order_host")


def template(secpar):
    vk0 = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))[f"vk_str_{secpar}_0"]
    m = re.fullmatch(r"(.*?values=\[)([-0-9, ]+)(\].*?values=\[)([-0-9, ]+)(\].*)", vk0, re.S)
    return m.group(1), m.group(3), m.group(5)


def keys_for(secpar, rng, n):
    d = 64 if secpar == 128 else 256
    h = (Q - 1) // 2
    vl = rng.integers(-h, h + 1, size=(n, d)).astype(np.int64)
    vr = rng.integers(-h, h + 1, size=(n, d)).astype(np.int64)
    small = [0, 1, -1, 10, -10, 100, 11, 19, 2, 20, 9, 99, 1000000000, 100000000, -100000000, h, -h, 107373286, 1073732864]
    for i in range(n // 2):                 # half of the keys start with a value from a small alphabet: many ties on the first coefficient
        vl[i, 0] = small[int(rng.integers(len(small)))]
    for i in range(0, n // 3):              # long common prefixes: the first difference is deep inside the left or the right half
        src = int(rng.integers(0, n))
        k = int(rng.integers(0, 2 * d))
        both = np.concatenate([vl[src], vr[src]])
        mine = np.concatenate([vl[i], vr[i]])
        mine[:k] = both[:k]
        if k < 2 * d and rng.integers(2):
            mine[k] = small[int(rng.integers(len(small)))]
        vl[i], vr[i] = mine[:d], mine[d:]
    vl[n - 1], vr[n - 1] = vl[0], vr[0]      # a fully equal pair: the sort must be stable
    vl[n - 2], vr[n - 2] = vl[1], vr[1]
    vl[n - 2, d - 1] = 7; vl[1, d - 1] = 70  # differ only in the LAST coefficient of the left polynomial: ']' is the terminator there
    vr[n - 3] = vr[2]; vl[n - 3] = vl[2]; vr[n - 3, d - 1] = -5; vr[2, d - 1] = -50
    return vl, vr


@pytest.mark.parametrize("secpar", [128, 256])
def test_rank_sort_equals_python_sorted_on_the_repr_strings(lib, secpar):
    head, mid, tail = template(secpar)
    rng = np.random.default_rng(secpar)
    for trial, n in enumerate((2, 3, 17, 120, 400)):
        vl, vr = keys_for(secpar, rng, max(n, 6)) if n >= 6 else keys_for(secpar, rng, 6)
        vl, vr = vl[-n:], vr[-n:]
        strs = [head + ", ".join(map(str, a)) + mid + ", ".join(map(str, b)) + tail for a, b in zip(vl.tolist(), vr.tolist())]
        want = sorted(range(n), key=lambda i: strs[i])           # Python's sort is stable, like the reference's
        a = np.ascontiguousarray(vl.astype(np.int32)); b = np.ascontiguousarray(vr.astype(np.int32))
        perm = np.empty(n, dtype=np.uint32)
        lib.h_rank_sort(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), C.c_int(n), C.c_int(a.shape[1]), perm.ctypes.data_as(C.c_void_p))
        assert perm.tolist() == want, (secpar, trial)
        for _ in range(200):                                      # and the pairwise comparison agrees with the strings
            i, j = (int(x) for x in rng.integers(0, n, size=2))
            c = lib.h_repr_cmp(a[i].ctypes.data_as(C.c_void_p), b[i].ctypes.data_as(C.c_void_p), a[j].ctypes.data_as(C.c_void_p),
                               b[j].ctypes.data_as(C.c_void_p), C.c_int(a.shape[1]))
            assert c == (strs[i] > strs[j]) - (strs[i] < strs[j]), (i, j)
