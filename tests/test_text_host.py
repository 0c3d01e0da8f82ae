"""csrc/fz_text.cuh -- the decimal formatter and the byte queue every sponge thread of the streaming hash kernels runs -- compiled for the HOST and
checked against Python's str(), hashlib and the XOF bytes the unmodified reference produced for hash_ch (fusion.py:412-419; fixtures of
tests/golden/make_golden.py).  tests/host/text_host.cpp drives the queue exactly like hash_ch_stream_kernel does: tokens of at most 16 bytes are
pushed while fewer than 136 bytes are queued, whole blocks are absorbed from the queue, the last one is padded."""
import ctypes as C
import hashlib
import json
import os
import re

import numpy as np
import pytest

import hostlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def lib():
    return hostlib.load("text_host")


def test_decimal_formatter_equals_str(lib):
    rng = np.random.default_rng(11)
    vals = [0, 1, -1, 9, -9, 10, -10, 99, 100, -100, 999, 1000, 9999, 10000, 99999, 100000, 999999, 1000000, 9999999, 10000000, 99999999, 100000000,
            999999999, 1000000000, -999999999, -1000000000, 1073732864, -1073732864, 2147483647, -2147483648]
    vals += [int(x) for x in rng.integers(-(1 << 31), 1 << 31, size=40000)]
    vals += [int(x) for x in rng.integers(-1000, 1000, size=4000)]
    buf = C.create_string_buffer(16)
    for v in vals:
        for sep in (0, 1):
            n = lib.h_fmt_value(C.c_int32(v), C.c_int(sep), buf)
            want = (str(v) + (", " if sep else "")).encode()
            assert n == len(want) and buf.raw[:n] == want, (v, sep, buf.raw[:n])
            assert buf.raw[n:] == bytes(16 - n), "bytes above the text must be zero (q_push ORs them into the queue)"


def test_prehash_decimal_equals_str_of_little_endian_int(lib):
    rng = np.random.default_rng(12)
    digests = [bytes(32), b"\x01" + bytes(31), bytes(31) + b"\x01", b"\xff" * 32, bytes(28) + b"\x00\xca\x9a\x3b", b"\x00\xca\x9a\x3b" + bytes(28)]
    digests += [rng.integers(0, 256, size=32, dtype=np.uint8).tobytes() for _ in range(3000)]
    digests += [rng.integers(0, 256, size=k, dtype=np.uint8).tobytes() + bytes(32 - k) for k in range(1, 32) for _ in range(8)]   # short numbers
    buf = C.create_string_buffer(96)
    for dg in digests:
        n = lib.h_fmt_le256(dg, buf)
        want = str(int.from_bytes(dg, "little")).encode()      # fusion.py:409 followed by str() in fusion.py:417
        assert n == len(want) and buf.raw[:n] == want, (dg.hex(), buf.raw[:n])


def _tokens(parts):
    """parts: list of bytes (fixed text) or list of ints (a value list rendered like Python's ", ".join(map(str, v)))."""
    kind, value, raw, raw_len = [], [], [], []
    for p in parts:
        if isinstance(p, (bytes, bytearray)):
            for o in range(0, len(p), 16):
                chunk = bytes(p[o:o + 16])
                kind.append(0); value.append(0); raw.append(chunk + bytes(16 - len(chunk))); raw_len.append(len(chunk))
        else:
            for j, v in enumerate(p):
                kind.append(1 if j + 1 < len(p) else 2); value.append(int(v)); raw.append(bytes(16)); raw_len.append(0)
    return (np.asarray(kind, dtype=np.int32), np.asarray(value, dtype=np.int32), np.frombuffer(b"".join(raw), dtype=np.uint8).copy(),
            np.asarray(raw_len, dtype=np.int32))


def _shake(lib, parts, n):
    kind, value, raw, raw_len = _tokens(parts)
    out = C.create_string_buffer(n)
    lib.h_shake_tokens(kind.ctypes.data_as(C.c_void_p), value.ctypes.data_as(C.c_void_p), raw.ctypes.data_as(C.c_void_p), raw_len.ctypes.data_as(C.c_void_p),
                       C.c_size_t(len(kind)), out, C.c_size_t(n))
    return out.raw


def test_queue_and_sponge_on_random_token_streams(lib):
    rng = np.random.default_rng(13)
    for trial in range(300):
        parts, text = [], b""
        for _ in range(int(rng.integers(0, 12))):
            if rng.integers(0, 2):
                chunk = rng.integers(0, 256, size=int(rng.integers(0, 70)), dtype=np.uint8).tobytes()
                parts.append(chunk); text += chunk
            else:
                lim = int(rng.choice([10, 1000, 1 << 31]))
                vals = [int(x) for x in rng.integers(-lim, lim, size=int(rng.integers(1, 40)))]
                parts.append(vals); text += ", ".join(map(str, vals)).encode()
        n = int(rng.choice([1, 32, 136, 137, 500]))
        assert _shake(lib, parts, n) == hashlib.shake_256(text).digest(n), trial


@pytest.mark.parametrize("secpar", [128, 256])
def test_hash_ch_xof_of_the_reference_keys(lib, secpar):
    """dst "," str(vk) "," str(prehash) through the product's formatter, queue and Keccak-f == the bytes the reference's
    hash_vk_and_int_to_bytes returned (fusion.py:412-419)."""
    J = json.load(open(os.path.join(GOLD, "golden.json")))
    Z = np.load(os.path.join(GOLD, "scheme_golden.npz"))
    vk_l, vk_r = Z[f"vk_l_{secpar}"], Z[f"vk_r_{secpar}"]
    vk0 = J[f"vk_str_{secpar}_0"]
    m = re.fullmatch(r"(.*?values=\[)([-0-9, ]+)(\].*?values=\[)([-0-9, ]+)(\].*)", vk0, re.S)
    head, mid, tail = m.group(1).encode(), m.group(3).encode(), m.group(5).encode()
    assert [int(x) for x in m.group(2).split(", ")] == [int(x) for x in vk_l[0]] and [int(x) for x in m.group(4).split(", ")] == [int(x) for x in vk_r[0]]
    dst = bytes([3 if secpar == 256 else 1, 1])                 # sign_hash_dst, fusion.py:38-55
    n_ch = J[f"n_ch_{secpar}"]
    for i, (pre, want_sha) in enumerate(zip(J[f"prehash_{secpar}"], J[f"ch_xof_sha_{secpar}"])):
        parts = [dst + b"," + head, [int(x) for x in vk_l[i]], mid, [int(x) for x in vk_r[i]], tail + b"," + pre.encode()]
        got = _shake(lib, parts, n_ch)
        assert hashlib.sha256(got).hexdigest() == want_sha, (secpar, i)
        if i == 0:
            assert got.hex() == J[f"ch_xof_hex_{secpar}_0"]
