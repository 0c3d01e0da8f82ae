"""csrc/fz_format.cuh -- the repr text one warp of the format kernels writes -- run unmodified on an emulated warp (32 host threads, shuffles and
__syncwarp through a barrier; tests/host/format_host.cpp) and compared with the strings the unmodified reference printed: str(vk) and
str(challenge) of the golden keys (fusion.py:328-329, 388-389 over algebra/matrices.py:40-41 and algebra/polynomials.py:257-258), the same
templates on random, short, negative and extreme values, and the decimal prehash.  No GPU needed."""
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
    return hostlib.load("format_host", pthread=True)


@pytest.fixture(scope="module")
def gold():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json"))), np.load(os.path.join(ROOT, "tests", "golden", "scheme_golden.npz"))


fmt_vk, fmt_ch = hostlib.fmt_vk, hostlib.fmt_ch


@pytest.mark.parametrize("secpar", [128, 256])
def test_reference_strings_of_the_golden_keys(lib, gold, secpar):
    J, Z = gold
    assert fmt_vk(lib, secpar, Z[f"vk_l_{secpar}"][0], Z[f"vk_r_{secpar}"][0]) == J[f"vk_str_{secpar}_0"]
    assert fmt_ch(lib, secpar, Z[f"c_hat_{secpar}"][0]) == J[f"ch_str_{secpar}_0"]


@pytest.mark.parametrize("secpar", [128, 256])
def test_random_and_edge_values_in_the_reference_templates(lib, gold, secpar):
    J, _ = gold
    m = re.fullmatch(r"(.*?values=\[)([-0-9, ]+)(\].*?values=\[)([-0-9, ]+)(\].*)", J[f"vk_str_{secpar}_0"], re.S)
    head, mid, tail = m.group(1), m.group(3), m.group(5)
    mc = re.fullmatch(r"(.*?values=\[)([-0-9, ]+)(\].*)", J[f"ch_str_{secpar}_0"], re.S)
    d = 64 if secpar == 128 else 256
    h = (Q - 1) // 2
    rng = np.random.default_rng(secpar + 7)
    rows = [np.zeros(d, dtype=np.int64), np.full(d, -h), np.full(d, h), rng.integers(-9, 10, size=d), rng.integers(-h, h + 1, size=d),
            rng.integers(-(1 << 31), 1 << 31, size=d), np.where(np.arange(d) % 3 == 0, 0, rng.integers(-1000, 1000, size=d))]
    for k in range(1, 10):
        rows.append(rng.integers(-(10 ** k), 10 ** k, size=d))        # every digit count
    for i, a in enumerate(rows):
        b = rows[(i + 3) % len(rows)]
        want = head + ", ".join(map(str, a.tolist())) + mid + ", ".join(map(str, b.tolist())) + tail
        assert fmt_vk(lib, secpar, a, b) == want, i
        assert fmt_ch(lib, secpar, a) == mc.group(1) + ", ".join(map(str, a.tolist())) + mc.group(3), i
        arr = np.ascontiguousarray(a.astype(np.int32))
        assert lib.h_values_len(arr.ctypes.data_as(C.c_void_p), C.c_int(d)) == len(", ".join(map(str, a.tolist())))


def test_prehash_decimal(lib):
    rng = np.random.default_rng(9)
    buf = C.create_string_buffer(96)
    for dg in [bytes(32), b"\xff" * 32, b"\x01" + bytes(31), bytes(31) + b"\x80"] + [rng.integers(0, 256, size=32, dtype=np.uint8).tobytes() for _ in range(500)]:
        n = lib.h_format_le256(dg, buf)
        assert buf.raw[:n] == str(int.from_bytes(dg, "little")).encode()
