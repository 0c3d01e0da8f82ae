"""ctypes binding of the CPU oracle (oracle/fz_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(fusion_cryptography_b200) must never import this module.

Data layout everywhere: int32 centred coefficients, row-major [batch][rows][d].
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libfz_oracle.so")
_lib = None

_i32p = C.POINTER(C.c_int32)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)
_i64p = C.POINTER(C.c_int64)
_u8p = C.POINTER(C.c_uint8)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "fz_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libfz_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.fzo_cent.restype = C.c_int64
        _lib.fzo_cent.argtypes = [C.c_int64]
        _lib.fzo_ring_cent.restype = C.c_int64
        _lib.fzo_ring_cent.argtypes = [C.c_int64, C.c_int64]
        _lib.fzo_format_vk.restype = C.c_size_t
        _lib.fzo_format_ch.restype = C.c_size_t
        _lib.fzo_format_prehash.restype = C.c_size_t
    return _lib


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(t)


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def set_threads(n: int) -> None:
    lib().fzo_set_threads(C.c_int(n))


def max_threads() -> int:
    return int(lib().fzo_get_max_threads())


def param(secpar: int, name: str) -> int:
    out = C.c_int64()
    rc = lib().fzo_param(C.c_int(secpar), name.encode(), C.byref(out))
    if rc:
        raise ValueError(f"unknown secpar/param {secpar}/{name}")
    return int(out.value)


def seed_limbs(seed: int) -> np.ndarray:
    """32-bit little-endian limbs of abs(seed), at least one (CPython random.seed)."""
    n = abs(int(seed))
    limbs = []
    while True:
        limbs.append(n & 0xFFFFFFFF)
        n >>= 32
        if not n:
            break
    return np.array(limbs, dtype=np.uint32)


def cent(v: int) -> int:
    return int(lib().fzo_cent(C.c_int64(v)))


def twiddles(secpar: int) -> Tuple[np.ndarray, np.ndarray]:
    d = param(secpar, "d")
    tw = np.zeros(d, dtype=np.int64)
    itw = np.zeros(d, dtype=np.int64)
    lib().fzo_twiddles(C.c_int(secpar), _p(tw, _i64p), _p(itw, _i64p))
    return tw, itw


def ntt_fwd(secpar: int, polys) -> np.ndarray:
    a = _i32(polys).copy()
    d = param(secpar, "d")
    assert a.size % d == 0
    lib().fzo_ntt_fwd(C.c_int(secpar), _p(a, _i32p), C.c_size_t(a.size // d))
    return a


def ntt_inv(secpar: int, polys) -> np.ndarray:
    a = _i32(polys).copy()
    d = param(secpar, "d")
    assert a.size % d == 0
    lib().fzo_ntt_inv(C.c_int(secpar), _p(a, _i32p), C.c_size_t(a.size // d))
    return a


def pointwise_mul(a, b) -> np.ndarray:
    a, b = _i32(a), _i32(b)
    out = np.empty_like(a)
    lib().fzo_pointwise_mul(_p(a, _i32p), _p(b, _i32p), _p(out, _i32p), C.c_size_t(a.size))
    return out


def pointwise_muladd(a, b, c) -> np.ndarray:
    a, b, c = _i32(a), _i32(b), _i32(c)
    out = np.empty_like(a)
    lib().fzo_pointwise_muladd(_p(a, _i32p), _p(b, _i32p), _p(c, _i32p), _p(out, _i32p), C.c_size_t(a.size))
    return out


def polymul(secpar: int, a, b) -> np.ndarray:
    a, b = _i32(a), _i32(b)
    d = param(secpar, "d")
    out = np.empty_like(a)
    lib().fzo_polymul(C.c_int(secpar), _p(a, _i32p), _p(b, _i32p), _p(out, _i32p), C.c_size_t(a.size // d))
    return out


def _i64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int64)


def ring_ntt_fwd(q: int, table, polys) -> np.ndarray:
    """cooley_tukey_ntt (algebra/ntt.py:271-291) for any odd prime q < 2^31; polys: [n, d] int32, table: d ints."""
    a = np.atleast_2d(_i32(polys)).copy()
    t = _i64(table)
    lib().fzo_ring_ntt_fwd(C.c_int64(q), C.c_int(a.shape[-1]), _p(t, _i64p), _p(a, _i32p), C.c_size_t(a.size // a.shape[-1]))
    return a


def ring_ntt_inv(q: int, table, polys) -> np.ndarray:
    """gentleman_sande_intt (algebra/ntt.py:350-377) incl. the n^-1 scaling."""
    a = np.atleast_2d(_i32(polys)).copy()
    t = _i64(table)
    lib().fzo_ring_ntt_inv(C.c_int64(q), C.c_int(a.shape[-1]), _p(t, _i64p), _p(a, _i32p), C.c_size_t(a.size // a.shape[-1]))
    return a


def ring_pointwise(q: int, a, b, mul: bool) -> np.ndarray:
    a, b = _i32(a), _i32(b)
    out = np.empty_like(a)
    lib().fzo_ring_pointwise(C.c_int64(q), C.c_int(1 if mul else 0), _p(a, _i32p), _p(b, _i32p), _p(out, _i32p), C.c_size_t(a.size))
    return out


def matvec(secpar: int, A, s) -> np.ndarray:
    A, s = _i32(A), _i32(s)
    d, ell = param(secpar, "d"), param(secpar, "ell")
    batch = s.size // (d * ell)
    out = np.empty((batch, d), dtype=np.int32)
    lib().fzo_matvec(C.c_int(secpar), _p(A, _i32p), _p(s, _i32p), _p(out, _i32p), C.c_size_t(batch))
    return out


def mt_stream(seed: int, n: int) -> np.ndarray:
    key = seed_limbs(seed)
    out = np.empty(n, dtype=np.uint32)
    lib().fzo_mt_stream(_p(key, _u32p), C.c_int(len(key)), _p(out, _u32p), C.c_int(n))
    return out


def sample_coef_poly(secpar: int, seed: int, norm_bound: int = 52, weight_bound: Optional[int] = None) -> np.ndarray:
    d = param(secpar, "d")
    if weight_bound is None:
        weight_bound = d
    key = seed_limbs(seed)
    out = np.empty(d, dtype=np.int32)
    lib().fzo_sample_coef_poly(C.c_int(secpar), _p(key, _u32p), C.c_int(len(key)), C.c_int(norm_bound), C.c_int(weight_bound), _p(out, _i32p))
    return out


def sample_ntt_poly(secpar: int, seed: int) -> np.ndarray:
    d = param(secpar, "d")
    key = seed_limbs(seed)
    out = np.empty(d, dtype=np.int32)
    lib().fzo_sample_ntt_poly(C.c_int(secpar), _p(key, _u32p), C.c_int(len(key)), _p(out, _i32p))
    return out


def sha3_256(data: bytes) -> bytes:
    out = (C.c_uint8 * 32)()
    lib().fzo_sha3_256(data, C.c_size_t(len(data)), out)
    return bytes(out)


def shake256(data: bytes, n: int) -> bytes:
    out = (C.c_uint8 * max(n, 1))()
    lib().fzo_shake256(data, C.c_size_t(len(data)), out, C.c_size_t(n))
    return bytes(out)[:n]


def format_vk(secpar: int, vk_l, vk_r) -> str:
    vl, vr = _i32(vk_l), _i32(vk_r)
    buf = C.create_string_buffer(16384)
    n = lib().fzo_format_vk(C.c_int(secpar), _p(vl, _i32p), _p(vr, _i32p), buf)
    return buf.raw[:n].decode()


def format_ch(secpar: int, c_hat) -> str:
    c = _i32(c_hat)
    buf = C.create_string_buffer(16384)
    n = lib().fzo_format_ch(C.c_int(secpar), _p(c, _i32p), buf)
    return buf.raw[:n].decode()


def format_prehash(digest: bytes) -> str:
    buf = C.create_string_buffer(128)
    n = lib().fzo_format_prehash(digest, buf)
    return buf.raw[:n].decode()


def hash_msg(secpar: int, msg: bytes) -> bytes:
    out = (C.c_uint8 * 32)()
    lib().fzo_hash_msg(C.c_int(secpar), msg, C.c_size_t(len(msg)), out)
    return bytes(out)


def decode_poly(secpar: int, b: bytes, log2_bias: int, norm_bound: int, weight_bound: int) -> np.ndarray:
    d = param(secpar, "d")
    out = np.empty(d, dtype=np.int32)
    rc = lib().fzo_decode_poly(C.c_int(secpar), b, C.c_size_t(len(b)), C.c_int(log2_bias), C.c_int(norm_bound), C.c_int(weight_bound), _p(out, _i32p))
    if rc == -1:
        raise ValueError("Too few bytes to decode polynomial.")
    return out


def hash_vk_and_int_to_bytes(secpar: int, vk_l, vk_r, digest: bytes, n: int) -> bytes:
    vl, vr = _i32(vk_l), _i32(vk_r)
    out = (C.c_uint8 * n)()
    lib().fzo_hash_vk_and_int_to_bytes(C.c_int(secpar), _p(vl, _i32p), _p(vr, _i32p), digest, out, C.c_size_t(n))
    return bytes(out)


def hash_ch(secpar: int, vk_l, vk_r, msg: bytes, want_coef: bool = False):
    vl, vr = _i32(vk_l), _i32(vk_r)
    d = param(secpar, "d")
    c_hat = np.empty(d, dtype=np.int32)
    c = np.empty(d, dtype=np.int32)
    rc = lib().fzo_hash_ch(C.c_int(secpar), _p(vl, _i32p), _p(vr, _i32p), msg, C.c_size_t(len(msg)), _p(c_hat, _i32p), _p(c, _i32p), None)
    assert rc == 0
    return (c_hat, c) if want_coef else c_hat


def setup_A(secpar: int, seed: int) -> np.ndarray:
    d, ell = param(secpar, "d"), param(secpar, "ell")
    key = seed_limbs(seed)
    A = np.empty((ell, d), dtype=np.int32)
    lib().fzo_setup_A(C.c_int(secpar), _p(key, _u32p), C.c_int(len(key)), _p(A, _i32p))
    return A


def keygen(secpar: int, A, seed: int):
    d, ell = param(secpar, "d"), param(secpar, "ell")
    A = _i32(A)
    kl, kr = seed_limbs(seed), seed_limbs(seed + 1)
    sk_l = np.empty((ell, d), dtype=np.int32)
    sk_r = np.empty((ell, d), dtype=np.int32)
    vk_l = np.empty(d, dtype=np.int32)
    vk_r = np.empty(d, dtype=np.int32)
    lib().fzo_keygen(C.c_int(secpar), _p(A, _i32p), _p(kl, _u32p), C.c_int(len(kl)), _p(kr, _u32p), C.c_int(len(kr)),
                     _p(sk_l, _i32p), _p(sk_r, _i32p), _p(vk_l, _i32p), _p(vk_r, _i32p))
    return sk_l, sk_r, vk_l, vk_r


def keygen_batch(secpar: int, A, seeds, want_sk: bool = True):
    d, ell = param(secpar, "d"), param(secpar, "ell")
    A = _i32(A)
    seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
    b = len(seeds)
    vk_l = np.empty((b, d), dtype=np.int32)
    vk_r = np.empty((b, d), dtype=np.int32)
    if want_sk:
        sk_l = np.empty((b, ell, d), dtype=np.int32)
        sk_r = np.empty((b, ell, d), dtype=np.int32)
        pl, pr = _p(sk_l, _i32p), _p(sk_r, _i32p)
    else:
        sk_l = sk_r = None
        pl = pr = None
    lib().fzo_keygen_batch(C.c_int(secpar), _p(A, _i32p), _p(seeds, _u64p), C.c_size_t(b), pl, pr, _p(vk_l, _i32p), _p(vk_r, _i32p))
    return sk_l, sk_r, vk_l, vk_r


def pack_msgs(msgs: Sequence[bytes]) -> Tuple[np.ndarray, np.ndarray]:
    offs = np.zeros(len(msgs) + 1, dtype=np.uint64)
    for i, m in enumerate(msgs):
        offs[i + 1] = offs[i] + len(m)
    buf = np.frombuffer(b"".join(msgs) + b"\0", dtype=np.uint8).copy()
    return buf, offs


def sign(secpar: int, sk_l, sk_r, vk_l, vk_r, msg: bytes) -> np.ndarray:
    d, ell = param(secpar, "d"), param(secpar, "ell")
    sk_l, sk_r, vk_l, vk_r = _i32(sk_l), _i32(sk_r), _i32(vk_l), _i32(vk_r)
    sig = np.empty((ell, d), dtype=np.int32)
    rc = lib().fzo_sign(C.c_int(secpar), _p(sk_l, _i32p), _p(sk_r, _i32p), _p(vk_l, _i32p), _p(vk_r, _i32p), msg, C.c_size_t(len(msg)), _p(sig, _i32p))
    assert rc == 0
    return sig


def sign_batch(secpar: int, sk_l, sk_r, vk_l, vk_r, msgs: Sequence[bytes]) -> np.ndarray:
    d, ell = param(secpar, "d"), param(secpar, "ell")
    sk_l, sk_r, vk_l, vk_r = _i32(sk_l), _i32(sk_r), _i32(vk_l), _i32(vk_r)
    b = len(msgs)
    buf, offs = pack_msgs(msgs)
    sigs = np.empty((b, ell, d), dtype=np.int32)
    rc = lib().fzo_sign_batch(C.c_int(secpar), _p(sk_l, _i32p), _p(sk_r, _i32p), _p(vk_l, _i32p), _p(vk_r, _i32p),
                              _p(buf, _u8p), _p(offs, _u64p), C.c_size_t(b), _p(sigs, _i32p))
    assert rc == 0
    return sigs


def sort_by_repr(secpar: int, vk_l, vk_r) -> np.ndarray:
    vk_l, vk_r = _i32(vk_l), _i32(vk_r)
    n = vk_l.shape[0]
    order = np.empty(n, dtype=np.uint32)
    lib().fzo_sort_by_repr(C.c_int(secpar), _p(vk_l, _i32p), _p(vk_r, _i32p), C.c_size_t(n), _p(order, _u32p))
    return order


def hash_ag(secpar: int, vk_l, vk_r, msgs: Sequence[bytes], want_ch: bool = False):
    """Inputs must already be in sorted order (fusion.py:665-669)."""
    d = param(secpar, "d")
    vk_l, vk_r = _i32(vk_l), _i32(vk_r)
    n = len(msgs)
    buf, offs = pack_msgs(msgs)
    alpha = np.empty((n, d), dtype=np.int32)
    ch = np.empty((n, d), dtype=np.int32)
    rc = lib().fzo_hash_ag(C.c_int(secpar), _p(vk_l, _i32p), _p(vk_r, _i32p), _p(buf, _u8p), _p(offs, _u64p), C.c_size_t(n), _p(alpha, _i32p), _p(ch, _i32p))
    assert rc == 0
    return (alpha, ch) if want_ch else alpha


def aggregate(secpar: int, vk_l, vk_r, msgs: Sequence[bytes], sigs) -> np.ndarray:
    d, ell = param(secpar, "d"), param(secpar, "ell")
    vk_l, vk_r, sigs = _i32(vk_l), _i32(vk_r), _i32(sigs)
    n = len(msgs)
    buf, offs = pack_msgs(msgs)
    out = np.empty((ell, d), dtype=np.int32)
    rc = lib().fzo_aggregate(C.c_int(secpar), _p(vk_l, _i32p), _p(vk_r, _i32p), _p(buf, _u8p), _p(offs, _u64p), _p(sigs, _i32p), C.c_size_t(n), _p(out, _i32p))
    assert rc == 0
    return out


VERIFY_REASONS = {
    0: "",
    1: "Too many keys.",
    2: "Number of keys and messages must be equal.",
    3: "Target doesn't match image of aggregate signature.",
    4: "Norm of aggregate signature too large.",
    5: "Weight of aggregate signature too large.",
}


def verify(secpar: int, A, vk_l, vk_r, msgs: Sequence[bytes], agg) -> int:
    A, vk_l, vk_r, agg = _i32(A), _i32(vk_l), _i32(vk_r), _i32(agg)
    n_keys = vk_l.shape[0]
    buf, offs = pack_msgs(msgs)
    return int(lib().fzo_verify(C.c_int(secpar), _p(A, _i32p), _p(vk_l, _i32p), _p(vk_r, _i32p), _p(buf, _u8p), _p(offs, _u64p),
                                C.c_size_t(n_keys), C.c_size_t(len(msgs)), _p(agg, _i32p)))
