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


def _root_of_unity(q, order):
    """a primitive `order`-th root of unity mod the prime q (order a power of two dividing q - 1)"""
    for g in range(2, q):
        r = pow(g, (q - 1) // order, q)
        if pow(r, order // 2, q) == q - 1:
            return r
    raise AssertionError


def _brv_table(q, d, psi):
    bits = d.bit_length() - 1
    return [pow(psi, int(format(i, f"0{bits}b")[::-1], 2) if bits else 0, q) for i in range(d)]


@pytest.mark.parametrize("q,d", [(17, 4), (257, 16), (7681, 64), (12289, 512), (2013265921, 1024), (Q, 256)])
def test_general_rings(lib, orc, q, d):
    """fz_ring_*: the transforms, pointwise operations and seeded samplers of the reference's algebra layer on rings other than the scheme's
    (the reference's own unit tests use moduli below 2^17, tests/test_ntt.py:18-50)."""
    psi = _root_of_unity(q, 2 * d)
    fwd, inv = _brv_table(q, d, psi), _brv_table(q, d, pow(psi, q - 2, q))
    tf, ti = np.array(fwd, dtype=np.uint32), np.array(inv, dtype=np.uint32)
    ring = C.c_void_p()
    assert lib.fz_ring_create(C.c_int(0), C.c_uint32(q), C.c_int(d.bit_length() - 1), P(tf), P(ti), C.byref(ring)) == 0, lib.fz_last_error()
    rng = np.random.default_rng(q % 9973 + d)
    n = 5
    a = rng.integers(-(1 << 31), 1 << 31, size=(n, d)).astype(np.int32)          # any int32
    b = rng.integers(-(q // 2), q // 2 + 1, size=(n, d)).astype(np.int32)
    out = np.full((n, d), 31337, dtype=np.int32)
    assert lib.fz_ring_ntt_fwd(ring, P(a), P(out), SZ(n), None) == 0, lib.fz_last_error()
    fa = orc.ring_ntt_fwd(q, fwd, a)
    assert np.array_equal(out, fa)
    assert lib.fz_ring_ntt_inv(ring, P(fa), P(out), SZ(n), None) == 0, lib.fz_last_error()
    assert np.array_equal(out, orc.ring_ntt_inv(q, inv, fa))
    assert np.array_equal(out.astype(np.int64), (a.astype(np.int64) + q // 2) % q - q // 2)      # INTT(NTT(a)) = cent(a)
    assert lib.fz_ring_pointwise_mul(ring, P(a), P(b), P(out), SZ(n), None) == 0
    assert np.array_equal(out, orc.ring_pointwise(q, a, b, True))
    assert lib.fz_ring_pointwise_add(ring, P(a), P(b), P(out), SZ(n), None) == 0
    assert np.array_equal(out, orc.ring_pointwise(q, a, b, False))
    import random

    seeds = [0, 7, 2**40 + 1, 2**70 + 3]
    keys, klen = seed_keys(seeds)
    smp = np.full((len(seeds), d), 31337, dtype=np.int32)
    assert lib.fz_ring_sample_ntt(ring, P(keys), C.c_int(8), P(klen), SZ(len(seeds)), P(smp), None) == 0
    for s, row in zip(seeds, smp):
        random.seed(s)
        assert row.tolist() == [random.randrange(q) - q // 2 for _ in range(d)]                  # algebra/polynomials.py:478-480
    bound, weight = min(52, q // 2), max(1, d // 3)
    assert lib.fz_ring_sample_coef(ring, P(keys), C.c_int(8), P(klen), SZ(len(seeds)), C.c_int(bound), C.c_int(weight), P(smp), None) == 0
    for s, row in zip(seeds, smp):
        random.seed(s)                                                                           # algebra/polynomials.py:446-459
        cf = [(1 + random.randrange(bound)) * (1 - 2 * random.randrange(2)) for _ in range(weight)] + [0] * (d - weight)
        for i in range(d - 1, 0, -1):
            j = random.randrange(i + 1)
            cf[i], cf[j] = cf[j], cf[i]
        assert row.tolist() == cf
    lib.fz_ring_destroy(ring)


@pytest.mark.parametrize("secpar", [128, 256])
def test_generic_hashes_text_and_decoder(lib, orc, secpar):
    import hashlib

    c = Ctx(lib, secpar)
    rng = np.random.default_rng(secpar + 3)
    msgs = [rng.integers(0, 256, size=L, dtype=np.uint8).tobytes() for L in (0, 1, 135, 136, 137, 500)]
    buf, offs = pack(msgs)
    out32 = np.zeros((len(msgs), 32), dtype=np.uint8)
    c.ok(lib.fz_sha3_256(c.h, P(buf), P(offs), SZ(len(msgs)), P(out32), None))
    assert [bytes(x) for x in out32] == [hashlib.sha3_256(m).digest() for m in msgs]
    xo = np.zeros((len(msgs), 300), dtype=np.uint8)
    c.ok(lib.fz_shake256(c.h, P(buf), P(offs), SZ(len(msgs)), P(xo), SZ(300), None))
    assert [bytes(x) for x in xo] == [hashlib.shake_256(m).digest(300) for m in msgs]
    # the materialised texts of the buffered pipeline: mode 0 = the hash_ch message, mode 1 = one element of the hash_ag list
    d = c.d
    vl = rng.integers(-H, H + 1, size=(3, d)).astype(np.int32); vr = rng.integers(-H, H + 1, size=(3, d)).astype(np.int32)
    ch = rng.integers(-H, H + 1, size=(3, d)).astype(np.int32)
    dg = np.stack([np.frombuffer(orc.hash_msg(secpar, b"t%d" % i), dtype=np.uint8) for i in range(3)]).copy()
    for mode in (0, 1):
        stride = int(lib.fz_text_stride(c.h, C.c_int(mode)))
        text = np.zeros((3, stride), dtype=np.uint8); tlen = np.zeros(3, dtype=np.uint32)
        c.ok(lib.fz_format_text(c.h, C.c_int(mode), P(vl), P(vr), P(dg), P(ch), SZ(3), P(text), P(tlen), None))
        for i in range(3):
            vk, pre = orc.format_vk(secpar, vl[i], vr[i]), orc.format_prehash(bytes(dg[i]))
            want = (bytes([3 if secpar == 256 else 1, 1]) + b"," + vk.encode() + b"," + pre.encode()) if mode == 0 else \
                ("(" + vk + ", " + pre + ", " + orc.format_ch(secpar, ch[i]) + ")").encode()
            assert bytes(text[i, :tlen[i]]) == want, (mode, i)
    # decode_bytes_to_polynomial_coefficients through the ABI, and its "Too few bytes" error (-3)
    blen = 12000 if secpar == 128 else 30000
    data = np.ascontiguousarray(rng.integers(0, 256, size=(2, blen), dtype=np.uint8))
    coef = c.i32(2, d)
    c.ok(lib.fz_decode_poly(c.h, P(data), SZ(blen), SZ(blen), SZ(2), C.c_longlong(secpar), C.c_longlong(52), C.c_int(27), P(coef), None))
    for i in range(2):
        assert coef[i].tolist() == [int(x) for x in orc.decode_poly(secpar, bytes(data[i]), secpar, 52, 27)]
    assert lib.fz_decode_poly(c.h, P(data), SZ(blen), SZ(10), SZ(2), C.c_longlong(secpar), C.c_longlong(52), C.c_int(27), P(coef), None) == -3
    lib.fz_destroy(c.h)


def test_host_pipelines_and_options(lib, orc):
    """fz_ntt_host / fz_ntt_pass_host (chunked, double-buffered host pipelines) and the alternative implementations behind fz_set_option."""
    secpar = 128
    c = Ctx(lib, secpar)
    rng = np.random.default_rng(77)
    n = 70
    a = rng.integers(-H, H + 1, size=(n, c.d)).astype(np.int32); b = rng.integers(-H, H + 1, size=(n, c.d)).astype(np.int32)
    out = c.i32(n, c.d)
    c.ok(lib.fz_ntt_host(c.h, C.c_int(0), P(a), P(out), SZ(n)))
    assert np.array_equal(out, orc.ntt_fwd(secpar, a))
    c.ok(lib.fz_ntt_pass_host(c.h, P(a), P(b), P(out), SZ(n)))
    assert np.array_equal(out, orc.ntt_inv(secpar, orc.pointwise_mul(orc.ntt_fwd(secpar, a), b)))
    # the buffered hashing pipeline and every sampler kernel give the bytes of the default ones
    A = orc.setup_A(secpar, 42)
    seeds = (3 * 10**6 + 2 * np.arange(3)).astype(np.uint64)
    _, _, vk_l, vk_r = orc.keygen_batch(secpar, A, seeds)
    msgs = [b"opt-%d" % i for i in range(3)]
    dg = np.stack([np.frombuffer(orc.hash_msg(secpar, m), dtype=np.uint8) for m in msgs]).copy()
    want_c = np.stack([orc.hash_ch(secpar, vk_l[i], vk_r[i], msgs[i]) for i in range(3)])
    order = orc.sort_by_repr(secpar, vk_l, vk_r)
    want_alpha = orc.hash_ag(secpar, vk_l[order], vk_r[order], [msgs[int(i)] for i in order])
    agg_off = np.array([0, 3], dtype=np.uint64)
    for buffered in (1, 0):
        c.ok(lib.fz_set_option(c.h, b"buffered_hash", C.c_longlong(buffered)))
        c_hat = c.i32(3, c.d)
        ws, wsn = c.ws("fz_hash_ch_workspace", 3)
        c.ok(lib.fz_hash_ch(c.h, P(vk_l), P(vk_r), P(dg), SZ(3), P(c_hat), P(ws), wsn, None))
        assert np.array_equal(c_hat, want_c), buffered
        for impl in (2, 3, 0):
            c.ok(lib.fz_set_option(c.h, b"hash_ag_impl", C.c_longlong(impl)))
            alpha = c.i32(3, c.d)
            ws, wsn = c.ws("fz_hash_ag_workspace", 3, 1, 3)
            vl_s, vr_s = np.ascontiguousarray(vk_l[order]), np.ascontiguousarray(vk_r[order])
            c.ok(lib.fz_hash_ag(c.h, P(vl_s), P(vr_s), P(np.ascontiguousarray(dg[order])), P(np.ascontiguousarray(want_c[order])), P(agg_off), SZ(1), SZ(3), SZ(3),
                                P(alpha), P(ws), wsn, None))
            assert np.array_equal(alpha, want_alpha), (buffered, impl)
    keys, klen = seed_keys([11, 2**33 + 5, 2**65 + 9])
    import random

    for impl in (0, 2, 1):
        c.ok(lib.fz_set_option(c.h, b"sample_impl", C.c_longlong(impl)))
        smp = c.i32(3, c.d)
        c.ok(lib.fz_sample_coef(c.h, P(keys), C.c_int(8), P(klen), SZ(3), C.c_int(52), C.c_int(c.d), P(smp), None))
        for s, row in zip([11, 2**33 + 5, 2**65 + 9], smp):
            random.seed(s)
            assert row.tolist() == [(1 + random.randrange(52)) * (1 - 2 * random.randrange(2)) for _ in range(c.d)], impl
    lib.fz_destroy(c.h)


@pytest.mark.parametrize("mode", ["ring-thread", "ring-warp", "ring-sliced", "streaming-windows"])
def test_windowed_aggregate_and_verify(lib, orc, mode):
    """Engine.aggregate / Engine.verify's window loop (fusion_cryptography_b200/batch.py) restated over the C ABI: the aggregation hash advanced window by
    window -- ring-fed (text / absorb_ring or the time-sliced absorb / squeeze_xof / decode) or through the streaming resumable kernels (begin / absorb /
    squeeze) --, the weighted sums and the verify target accumulated per window with fz_pointwise_add, fz_verify_finish on the complete target,
    fz_fold_partials over partial aggregates.  Three aggregates of 4, 1 and 3 signers, windows of two sorted positions (fusion.py:655-728)."""
    secpar = 128
    c = Ctx(lib, secpar)
    for f in ("fz_hash_ag_state_bytes", "fz_hash_ag_ring_bytes", "fz_hash_ag_xof_bytes", "fz_hash_ag_text_workspace", "fz_verify_finish_workspace"):
        getattr(lib, f).restype = C.c_size_t
    d, ell = c.d, c.ell
    counts, Pw = [4, 1, 3], 2
    n, n_aggs = sum(counts), len(counts)
    A = orc.setup_A(secpar, 42)
    seeds = (4 * 10**6 + 2 * np.arange(n)).astype(np.uint64)
    sk_l, sk_r, vk_l, vk_r = orc.keygen_batch(secpar, A, seeds)
    msgs = [b"win-%d" % i for i in range(n)]
    sigs = orc.sign_batch(secpar, sk_l, sk_r, vk_l, vk_r, msgs)
    agg_off = np.zeros(n_aggs + 1, dtype=np.uint64); agg_off[1:] = np.cumsum(counts)
    want_aggs, order = [], []
    for j in range(n_aggs):
        sl = slice(int(agg_off[j]), int(agg_off[j + 1]))
        want_aggs.append(orc.aggregate(secpar, vk_l[sl], vk_r[sl], msgs[sl], sigs[sl]))
        order += [sl.start + int(i) for i in orc.sort_by_repr(secpar, vk_l[sl], vk_r[sl])]
    want_aggs = np.stack(want_aggs)
    dg = np.stack([np.frombuffer(orc.hash_msg(secpar, m), dtype=np.uint8) for m in msgs]).copy()
    c_hat = np.stack([orc.hash_ch(secpar, vk_l[i], vk_r[i], msgs[i]) for i in range(n)])
    pos_sorted = np.array(order)                       # sorted position -> original position (what fz_sort_keys returns, checked in the one-shot test)

    if mode == "ring-thread":
        c.ok(lib.fz_set_option(c.h, b"hash_ag_impl", C.c_longlong(2)))
    elif mode in ("ring-warp", "ring-sliced"):
        c.ok(lib.fz_set_option(c.h, b"hash_ag_impl", C.c_longlong(3)))
    state = np.zeros(n_aggs * int(lib.fz_hash_ag_state_bytes(c.h)), dtype=np.uint8)
    c.ok(lib.fz_hash_ag_begin(c.h, P(state), P(agg_off), SZ(n_aggs), None))
    ring_cap = int(lib.fz_hash_ag_ring_bytes(c.h, SZ(3 * Pw if mode == "ring-sliced" else Pw)))
    ring = np.zeros(n_aggs * ring_cap, dtype=np.uint8)
    text_pos = np.zeros(n_aggs, dtype=np.uint64)
    wins = []
    for w in range(0, max(counts), Pw):
        cnt = np.clip(np.array(counts) - w, 0, Pw)
        win_off = np.zeros(n_aggs + 1, dtype=np.uint64); win_off[1:] = np.cumsum(cnt)
        idx = np.concatenate([pos_sorted[int(agg_off[j]) + w: int(agg_off[j]) + w + int(cnt[j])] for j in range(n_aggs)])
        agg_of = np.repeat(np.arange(n_aggs, dtype=np.uint32), cnt)
        vl_w, vr_w = np.ascontiguousarray(vk_l[idx]), np.ascontiguousarray(vk_r[idx])
        dg_w, ch_w = np.ascontiguousarray(dg[idx]), np.ascontiguousarray(c_hat[idx])
        n_win = len(idx)
        if mode == "streaming-windows":
            c.ok(lib.fz_hash_ag_absorb(c.h, P(state), P(vl_w), P(vr_w), P(dg_w), P(ch_w), P(win_off), SZ(n_aggs), None))
        else:
            wend = np.zeros(n_aggs, dtype=np.uint64)
            tws = np.zeros(int(lib.fz_hash_ag_text_workspace(c.h, SZ(n_win))) + 256, dtype=np.uint8)
            c.ok(lib.fz_hash_ag_text(c.h, P(vl_w), P(vr_w), P(dg_w), P(ch_w), P(win_off), P(agg_of), P(agg_off), SZ(n_aggs), SZ(n_win), C.c_uint32(w),
                                     P(text_pos), P(ring), SZ(ring_cap), P(state), P(wend), P(tws), SZ(len(tws)), None))
            if mode == "ring-sliced":      # two of the three chains advance per launch, the resting one rotates (fz_hash_ag_absorb_sliced)
                c.ok(lib.fz_hash_ag_absorb_sliced(c.h, P(state), P(ring), SZ(ring_cap), P(wend), SZ(n_aggs), SZ((w // Pw + 1) % n_aggs), SZ(2), C.c_uint64(60), None))
            else:
                c.ok(lib.fz_hash_ag_absorb_ring(c.h, P(state), P(ring), SZ(ring_cap), P(wend), P(win_off), SZ(n_aggs), None))
        wins.append((win_off, agg_of, idx, vl_w, vr_w, ch_w, n_win))
    if mode == "ring-sliced":              # the text is complete: one launch over every chain without a budget closes what is left
        c.ok(lib.fz_hash_ag_absorb_sliced(c.h, P(state), P(ring), SZ(ring_cap), P(wend), SZ(n_aggs), SZ(0), SZ(n_aggs), C.c_uint64(1 << 62), None))
    assert (state.view(np.int32).reshape(n_aggs, 128)[:, 107] == 1).all(), "every sponge has closed its message"

    xof_cap = int(lib.fz_hash_ag_xof_bytes(c.h, SZ(Pw)))
    acc_agg, acc_target = None, None
    for win_off, agg_of, idx, vl_w, vr_w, ch_w, n_win in wins:
        alpha_w = c.i32(n_win, d)
        if mode == "streaming-windows":
            c.ok(lib.fz_hash_ag_squeeze(c.h, P(state), P(win_off), SZ(n_aggs), SZ(n_win), P(alpha_w), None))
        else:
            xof = np.zeros(n_aggs * xof_cap, dtype=np.uint8); bpos = np.zeros(n_aggs, dtype=np.uint32)
            c.ok(lib.fz_hash_ag_squeeze_xof(c.h, P(state), P(win_off), SZ(n_aggs), P(xof), SZ(xof_cap), P(bpos), None))
            c.ok(lib.fz_hash_ag_decode(c.h, P(xof), SZ(xof_cap), P(bpos), P(win_off), P(agg_of), SZ(n_win), P(alpha_w), None))
        rows = np.ascontiguousarray(idx.astype(np.uint32))
        part = c.i32(n_aggs, ell, d)
        ws, wsn = c.ws("fz_aggregate_sum_workspace", n_aggs, Pw)
        c.ok(lib.fz_aggregate_sum(c.h, P(sigs), P(rows), P(alpha_w), P(win_off), SZ(n_aggs), SZ(Pw), P(part), P(ws), wsn, None))
        tgt = c.i32(n_aggs, d)
        c.ok(lib.fz_verify_target(c.h, P(vl_w), P(vr_w), P(ch_w), P(alpha_w), P(win_off), SZ(n_aggs), P(tgt), None))
        if acc_agg is None:
            acc_agg, acc_target, parts = part, tgt, [part.copy()]
        else:
            parts.append(part.copy())
            c.ok(lib.fz_pointwise_add(c.h, P(acc_agg), P(part), P(acc_agg), SZ(n_aggs * ell), None))
            c.ok(lib.fz_pointwise_add(c.h, P(acc_target), P(tgt), P(acc_target), SZ(n_aggs), None))
    assert np.array_equal(acc_agg, want_aggs)
    # the same sum as the multi-GPU reduce forms it: partial aggregates [n_aggs][n_parts][ell][d] folded in one call
    stacked = np.ascontiguousarray(np.stack(parts, axis=1))
    folded = c.i32(n_aggs, ell, d)
    c.ok(lib.fz_fold_partials(c.h, P(stacked), C.c_int(len(parts)), SZ(n_aggs), P(folded), None))
    assert np.array_equal(folded, want_aggs)
    codes = np.full(n_aggs, 99, dtype=np.uint8)
    ws = np.zeros(int(lib.fz_verify_finish_workspace(c.h, SZ(n_aggs))) + 256, dtype=np.uint8)
    c.ok(lib.fz_verify_finish(c.h, P(A), P(acc_target), P(acc_agg), SZ(n_aggs), P(codes), P(ws), SZ(len(ws)), None))
    assert codes.tolist() == [0, 0, 0]
    bad = acc_agg.copy(); bad[2, 5, 7] -= 1
    c.ok(lib.fz_verify_finish(c.h, P(A), P(acc_target), P(bad), SZ(n_aggs), P(codes), P(ws), SZ(len(ws)), None))
    assert codes.tolist() == [0, 0, 3]
    c.ok(lib.fz_set_option(c.h, b"hash_ag_impl", C.c_longlong(0)))
    lib.fz_destroy(c.h)
