"""The C ABI of include/fusion_b200.h END TO END on the CPU: every csrc/*.cu -- kernels, launch code and the fz_* entry points -- compiled by g++ on the
CUDA-on-CPU emulator (tests/host/cuda_emu.h; "device" pointers are host pointers, the one device has two SMs) and called through ctypes with numpy
arrays exactly as fusion_cryptography_b200/native.py calls the real library with device tensors.  Everything is compared bit for bit with the oracle
(and through it with the reference, tests/test_oracle_golden.py): setup, the transforms and the fused pass, keygen in its three forms, keygen + sign,
the challenge and aggregation hashes, sort / gather, the aggregate sum and verify with its reason codes.  This is the product's source on an emulated
device, not a CPU path of the product: the package cannot load this build.  No GPU needed."""
import ctypes as C

import numpy as np
import pytest

import hostlib

Q = hostlib.Q
H = (Q - 1) // 2
P = hostlib.ptr
SZ = C.c_size_t


@pytest.fixture(scope="module")
def lib():
    lib = hostlib.load_emulated_library()
    for f in ("fz_hash_ch_workspace", "fz_hash_ag_workspace", "fz_aggregate_sum_workspace", "fz_verify_check_workspace", "fz_keygen_workspace",
              "fz_keygen_rows_workspace", "fz_keygen_sign_workspace", "fz_text_stride"):
        getattr(lib, f).restype = C.c_size_t
    lib.fz_param.restype = C.c_longlong
    return lib


class Ctx:
    def __init__(self, lib, secpar):
        self.lib, self.secpar = lib, secpar
        self.h = C.c_void_p()
        assert lib.fz_init(C.c_int(0), C.c_int(secpar), C.byref(self.h)) == 0, lib.fz_last_error()
        self.d, self.ell = int(lib.fz_param(self.h, b"d")), int(lib.fz_param(self.h, b"ell"))

    def ok(self, rc):
        assert rc == 0, self.lib.fz_last_error()

    def ws(self, name, *sizes):
        n = int(getattr(self.lib, name)(self.h, *[SZ(s) for s in sizes]))
        return np.empty(max(n, 256), dtype=np.uint8), SZ(max(n, 256))

    def i32(self, *shape):
        return np.full(shape, 424242, dtype=np.int32)


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle

    return oracle


def seed_keys(seeds):
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


def pack(msgs):
    offs = np.zeros(len(msgs) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(m) for m in msgs])
    buf = np.frombuffer(b"".join(msgs) + b"\0" * 16, dtype=np.uint8).copy()
    return buf, offs


@pytest.mark.parametrize("secpar", [128, 256])
def test_transforms_and_products(lib, orc, secpar):
    c = Ctx(lib, secpar)
    rng = np.random.default_rng(secpar)
    n = 37                                               # ragged: not a multiple of the polynomials a warp or a CTA owns
    a = rng.integers(-H, H + 1, size=(n, c.d)).astype(np.int32)
    a[:5] = rng.integers(-(1 << 31), 1 << 31, size=(5, c.d))     # any int32 is a legal input
    b = rng.integers(-H, H + 1, size=(n, c.d)).astype(np.int32)
    out = c.i32(n, c.d)
    c.ok(lib.fz_ntt_fwd(c.h, P(a), P(out), SZ(n), None))
    fa = orc.ntt_fwd(secpar, a)
    assert np.array_equal(out, fa)
    c.ok(lib.fz_ntt_inv(c.h, P(fa), P(out), SZ(n), None))
    assert np.array_equal(out, orc.ntt_inv(secpar, fa))
    c.ok(lib.fz_polymul(c.h, P(a), P(b), P(out), SZ(n), None))
    assert np.array_equal(out, orc.polymul(secpar, a, b))
    c.ok(lib.fz_ntt_pass(c.h, P(a), P(b), P(out), SZ(n), None))
    assert np.array_equal(out, orc.ntt_inv(secpar, orc.pointwise_mul(fa, b)))
    c.ok(lib.fz_pointwise_mul(c.h, P(a), P(b), P(out), SZ(n), None))
    assert np.array_equal(out, orc.pointwise_mul(a, b))
    c.ok(lib.fz_pointwise_muladd(c.h, P(a), P(b), P(fa), P(out), SZ(n), None))
    assert np.array_equal(out, orc.pointwise_muladd(a, b, fa))
    A = rng.integers(-H, H + 1, size=(c.ell, c.d)).astype(np.int32)
    s = rng.integers(-(1 << 31), 1 << 31, size=(2, c.ell, c.d)).astype(np.int32)
    mv = c.i32(2, c.d)
    c.ok(lib.fz_matvec(c.h, P(A), P(s), P(mv), SZ(2), None))
    assert np.array_equal(mv, orc.matvec(secpar, A, s))
    lib.fz_destroy(c.h)


@pytest.mark.parametrize("secpar", [128, 256])
def test_scheme_end_to_end(lib, orc, secpar):
    c = Ctx(lib, secpar)
    d, ell = c.d, c.ell
    # fusion_setup(secpar, 42): every column of the public matrix is sampled with the same seed (fusion.py:273-282)
    keys, klen = seed_keys([42])
    a_row = c.i32(1, d)
    c.ok(lib.fz_sample_ntt(c.h, P(keys), C.c_int(8), P(klen), SZ(1), P(a_row), None))
    A = np.ascontiguousarray(np.repeat(a_row, ell, axis=0))
    assert np.array_equal(A, orc.setup_A(secpar, 42))

    n = 5
    seeds = (9 * 10**6 + 2 * np.arange(n)).astype(np.uint64)
    msgs = [f"abi-emu-{i}".encode() for i in range(n)]
    sk_l, sk_r, vk_l, vk_r = orc.keygen_batch(secpar, A, seeds)
    want_sigs = orc.sign_batch(secpar, sk_l, sk_r, vk_l, vk_r, msgs)

    # keygen + sign in one call (fusion.py:338-373, 534-557)
    buf, offs = pack(msgs)
    g_vl, g_vr, g_sig = c.i32(n, d), c.i32(n, d), c.i32(n, ell, d)
    ws, wsn = c.ws("fz_keygen_sign_workspace", n)
    c.ok(lib.fz_keygen_sign(c.h, P(A), P(seeds), P(buf), P(offs), SZ(n), P(g_vl), P(g_vr), P(g_sig), P(ws), wsn, None))
    assert np.array_equal(g_vl, vk_l) and np.array_equal(g_vr, vk_r) and np.array_equal(g_sig, want_sigs)

    # keygen with the secret keys materialised, and in its general ell-row form with the same seeds in every row
    k_sl, k_sr, k_hat, k_vl, k_vr = c.i32(n, ell, d), c.i32(n, ell, d), c.i32(n, 2, d), c.i32(n, d), c.i32(n, d)
    ws, wsn = c.ws("fz_keygen_workspace", n)
    c.ok(lib.fz_keygen(c.h, P(A), P(seeds), SZ(n), P(k_sl), P(k_sr), P(k_hat), P(k_vl), P(k_vr), P(ws), wsn, None))
    assert np.array_equal(k_sl, sk_l) and np.array_equal(k_sr, sk_r) and np.array_equal(k_vl, vk_l) and np.array_equal(k_vr, vk_r)
    assert np.array_equal(k_hat[:, 0], sk_l[:, 0]) and np.array_equal(k_hat[:, 1], sk_r[:, 0])
    row_seeds = np.ascontiguousarray(np.stack([np.repeat(seeds[:2, None], ell, axis=1), np.repeat(seeds[:2, None] + 1, ell, axis=1)], axis=1))
    r_sl, r_sr, r_vl, r_vr = c.i32(2, ell, d), c.i32(2, ell, d), c.i32(2, d), c.i32(2, d)
    ws, wsn = c.ws("fz_keygen_rows_workspace", 2)
    c.ok(lib.fz_keygen_rows(c.h, P(A), P(row_seeds), SZ(2), P(r_sl), P(r_sr), P(r_vl), P(r_vr), P(ws), wsn, None))
    assert np.array_equal(r_sl, sk_l[:2]) and np.array_equal(r_sr, sk_r[:2]) and np.array_equal(r_vl, vk_l[:2]) and np.array_equal(r_vr, vk_r[:2])

    # sign from its parts: prehash, challenge, multiply-add (fusion.py:405-409, 511-531, 557)
    dg = np.zeros((n, 32), dtype=np.uint8)
    c.ok(lib.fz_hash_msg(c.h, P(buf), P(offs), SZ(n), P(dg), None))
    assert [bytes(x) for x in dg] == [orc.hash_msg(secpar, m) for m in msgs]
    c_hat = c.i32(n, d)
    ws, wsn = c.ws("fz_hash_ch_workspace", n)
    c.ok(lib.fz_hash_ch(c.h, P(vk_l), P(vk_r), P(dg), SZ(n), P(c_hat), P(ws), wsn, None))
    assert np.array_equal(c_hat, np.stack([orc.hash_ch(secpar, vk_l[i], vk_r[i], msgs[i]) for i in range(n)]))
    sig2 = c.i32(n, ell, d)
    c.ok(lib.fz_sign_mac(c.h, P(sk_l), P(sk_r), P(c_hat), P(sig2), SZ(n), None))
    assert np.array_equal(sig2, want_sigs)
    c.ok(lib.fz_sign_mac_rows(c.h, P(k_hat), P(c_hat), P(sig2), SZ(n), None))
    assert np.array_equal(sig2, want_sigs)

    # aggregate (fusion.py:655-677): two aggregates of 3 and 2 signers
    counts = [3, 2]
    agg_off = np.array([0, 3, 5], dtype=np.uint64)
    perm = np.zeros(n, dtype=np.uint32)
    c.ok(lib.fz_sort_keys(c.h, P(vk_l), P(vk_r), None, P(agg_off), SZ(2), SZ(3), P(perm), None))
    want_perm = np.concatenate([orc.sort_by_repr(secpar, vk_l[:3], vk_r[:3]), 3 + orc.sort_by_repr(secpar, vk_l[3:], vk_r[3:])])
    assert perm.tolist() == [int(x) for x in want_perm]
    vl_s, vr_s, dg_s = c.i32(n, d), c.i32(n, d), np.zeros((n, 32), dtype=np.uint8)
    c.ok(lib.fz_gather_sorted(c.h, P(vk_l), P(vk_r), P(dg), None, P(perm), SZ(n), P(vl_s), P(vr_s), P(dg_s), None))
    assert np.array_equal(vl_s, vk_l[perm]) and np.array_equal(dg_s, dg[perm])
    ch_s = np.ascontiguousarray(c_hat[perm])
    alpha = c.i32(n, d)
    ws, wsn = c.ws("fz_hash_ag_workspace", n, 2, 3)
    c.ok(lib.fz_hash_ag(c.h, P(vl_s), P(vr_s), P(dg_s), P(ch_s), P(agg_off), SZ(2), SZ(n), SZ(3), P(alpha), P(ws), wsn, None))
    want_alpha = np.concatenate([orc.hash_ag(secpar, vk_l[sl][o], vk_r[sl][o], [msgs[sl.start + int(i)] for i in o])
                                 for sl, o in ((slice(0, 3), want_perm[:3]), (slice(3, 5), want_perm[3:] - 3))])
    assert np.array_equal(alpha, want_alpha)
    aggs = c.i32(2, ell, d)
    ws, wsn = c.ws("fz_aggregate_sum_workspace", 2, 3)
    c.ok(lib.fz_aggregate_sum(c.h, P(want_sigs), P(perm), P(alpha), P(agg_off), SZ(2), SZ(3), P(aggs), P(ws), wsn, None))
    want_aggs = np.stack([orc.aggregate(secpar, vk_l[:3], vk_r[:3], msgs[:3], want_sigs[:3]), orc.aggregate(secpar, vk_l[3:], vk_r[3:], msgs[3:], want_sigs[3:])])
    assert np.array_equal(aggs, want_aggs)

    # verify (fusion.py:680-728): accept, and reject a tampered aggregate with the reference's reason
    codes = np.full(2, 99, dtype=np.uint8)
    ws, wsn = c.ws("fz_verify_check_workspace", 2)
    c.ok(lib.fz_verify_check(c.h, P(A), P(vl_s), P(vr_s), P(ch_s), P(alpha), P(aggs), P(agg_off), SZ(2), P(codes), P(ws), wsn, None))
    assert codes.tolist() == [0, 0]
    assert orc.verify(secpar, A, vk_l[:3], vk_r[:3], msgs[:3], want_aggs[0]) == 0
    bad = aggs.copy()
    bad[1, 0, 0] += 1
    c.ok(lib.fz_verify_check(c.h, P(A), P(vl_s), P(vr_s), P(ch_s), P(alpha), P(bad), P(agg_off), SZ(2), P(codes), P(ws), wsn, None))
    assert codes.tolist() == [0, 3]                          # "Target doesn't match image of aggregate signature."
    lib.fz_destroy(c.h)
