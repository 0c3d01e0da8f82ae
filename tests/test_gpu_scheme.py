"""GPU parity tests for K7-K11: Keccak, repr formatter, decoder, MT19937 samplers, keygen, hash_ch,
sort-by-repr, hash_ag, aggregation sum and verify -- CUDA through the C ABI vs hashlib / random (the
reference's stdlib dependencies), the CPU oracle and the reference-generated goldens / KATs."""
import hashlib
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

Q = 2147465729
H = (Q - 1) // 2


@pytest.fixture(scope="module")
def ctxs():
    from fusion_cryptography_b200 import native

    return {sp: native.get_context(sp, 0) for sp in (128, 256)}


def dev(a, dtype=np.int32):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).cuda()


def host(t):
    return t.cpu().numpy()


def test_keccak_vs_hashlib(ctxs):
    ctx = ctxs[256]
    rng = random.Random(11)
    msgs = [bytes(rng.getrandbits(8) for _ in range(n)) for n in (0, 1, 3, 135, 136, 137, 271, 272, 273, 1000, 6439, 20000)]
    msgs += [bytes(rng.getrandbits(8) for _ in range(rng.randrange(0, 700))) for _ in range(300)]
    got = host(ctx.sha3_256(msgs))
    for m, g in zip(msgs, got):
        assert bytes(g) == hashlib.sha3_256(m).digest()
    for out_len in (1, 32, 136, 137, 1551, 10436):
        got = host(ctx.shake256(msgs[:40], out_len))
        for m, g in zip(msgs, got):
            assert bytes(g) == hashlib.shake_256(m).digest(out_len), (len(m), out_len)


@pytest.mark.parametrize("secpar", [128, 256])
def test_hash_msg(ctxs, oracle, golden_json, secpar):
    ctx = ctxs[secpar]
    msgs = [m.encode() for m in golden_json[f"msgs_{secpar}"]] + [b"", b"x" * 133, b"y" * 134, b"z" * 500]
    got = host(ctx.hash_msg(msgs))
    for m, g in zip(msgs, got):
        assert bytes(g) == oracle.hash_msg(secpar, m)
    for i, want in enumerate(golden_json[f"prehash_{secpar}"]):
        assert str(int.from_bytes(bytes(got[i]), "little")) == want
    if secpar == 128:  # reference KAT intermediate_hash_message_to_int_KAT_128.csv
        kat = golden_json["kat_hash_msg_128"]
        got = host(ctx.hash_msg([m.encode() for m, _ in kat]))
        for (m, want), g in zip(kat, got):
            assert str(int.from_bytes(bytes(g), "little")) == want


@pytest.mark.parametrize("secpar", [128, 256])
def test_formatter(ctxs, oracle, secpar):
    ctx = ctxs[secpar]
    d = ctx.d
    rng = np.random.default_rng(secpar)
    n = 37
    vl = rng.integers(-H, H + 1, size=(n, d)).astype(np.int32)
    vr = rng.integers(-H, H + 1, size=(n, d)).astype(np.int32)
    ch = rng.integers(-H, H + 1, size=(n, d)).astype(np.int32)
    # edge rows: extremes (longest text), zeros / small / single digits (shortest text)
    vl[0] = -H; vr[0] = -H; ch[0] = -H
    vl[1] = 0; vr[1] = 0; ch[1] = 0
    vl[2] = rng.integers(-9, 10, size=d); vr[2] = rng.integers(-120, 120, size=d)
    vl[3, :8] = [0, -1, 9, 10, -10, 99, 100, 999999999]
    dg = rng.integers(0, 256, size=(n, 32)).astype(np.uint8)
    dg[4] = 0          # prehash int 0 -> "0"
    dg[5] = 255
    dg[6, 1:] = 0      # tiny int
    dst = bytes([3 if secpar == 256 else 1, 1]) + b","
    out, lens = ctx.format_text(0, dev(vl), dev(vr), dev(dg, np.uint8))
    out, lens = host(out), host(lens)
    for i in range(n):
        want = dst + oracle.format_vk(secpar, vl[i], vr[i]).encode() + b"," + oracle.format_prehash(bytes(dg[i])).encode()
        assert bytes(out[i, :lens[i]]) == want, i
    out, lens = ctx.format_text(1, dev(vl), dev(vr), dev(dg, np.uint8), dev(ch))
    out, lens = host(out), host(lens)
    for i in range(n):
        want = ("(" + oracle.format_vk(secpar, vl[i], vr[i]) + ", " + oracle.format_prehash(bytes(dg[i])) + ", " + oracle.format_ch(secpar, ch[i]) + ")").encode()
        assert bytes(out[i, :lens[i]]) == want, i


def test_decoder_goldens(ctxs, golden_json):
    # fusion/fusion.py:422-481: all-0 / all-1 bytes (reference tests/test_fusion.py:475-558), random bytes,
    # general bounds, byte exhaustion (quirk Q4) and the too-few-bytes error
    for case in golden_json["decoder"]:
        ctx = ctxs[case["secpar"]]
        b = np.frombuffer(bytes.fromhex(case["bytes_hex"]), dtype=np.uint8)
        data = torch.from_numpy(np.stack([b, b, b]).copy()).cuda()
        if case["coefs"] is None:
            with pytest.raises(ValueError, match="Too few bytes"):
                ctx.decode_poly(data, len(b), case["secpar"], case["bound"], case["weight"])
        else:
            got = host(ctx.decode_poly(data, len(b), case["secpar"], case["bound"], case["weight"]))
            assert got[0].tolist() == case["coefs"] and (got == got[0]).all(), (case["secpar"], case["weight"], case["bound"])


@pytest.mark.parametrize("secpar", [128, 256])
def test_samplers(ctxs, oracle, scheme_golden, kat_golden, secpar):
    ctx = ctxs[secpar]
    d = ctx.d
    seeds = [0, 1, 42, 43, 2**32 - 1, 2**32, 2**40 + 12345, 2**64 - 1, 2**64, 2**100 + 7, -5] + list(range(1000, 1100))
    got = host(ctx.sample_coef(seeds))
    for s, g in zip(seeds, got):
        random.seed(s)  # algebra/polynomials.py:448-453
        assert g.tolist() == [(1 + random.randrange(52)) * (1 - 2 * random.randrange(2)) for _ in range(d)], s
    # the dense secret-key kernel seeds a warp from two registers when every seed of the warp has at most two 32-bit limbs: one- and two-limb
    # seeds mixed, more generators than one CTA holds, each of the three kernels that can serve the call (fz_set_option "sample_impl")
    seeds2 = [2**33 + 17 * i for i in range(100)] + list(range(5000, 5100)) + [2**64 - 1 - i for i in range(30)] + [2**32 - 1, 2**32, 2**32 + 1]
    want2 = []
    for s in seeds2:
        random.seed(s)
        want2.append([(1 + random.randrange(52)) * (1 - 2 * random.randrange(2)) for _ in range(d)])
    for impl in (1, 2, 0, 1):
        ctx.set_option("sample_impl", impl)
        assert host(ctx.sample_coef(seeds2)).tolist() == want2, impl
        assert host(ctx.sample_coef(seeds))[:11].tolist() == got[:11].tolist(), impl
    for bound in (1, 2, 100, 127, 128, 300):   # 128 and above fall back to the general kernel (magnitudes no longer fit the byte tile)
        g2 = host(ctx.sample_coef(seeds2[:40] + seeds[:11], norm_bound=bound))
        for s, g in zip(seeds2[:40] + seeds[:11], g2):
            random.seed(s)
            assert g.tolist() == [(1 + random.randrange(bound)) * (1 - 2 * random.randrange(2)) for _ in range(d)], (s, bound)
    got = host(ctx.sample_ntt(seeds))
    for s, g in zip(seeds, got):
        random.seed(s)  # algebra/polynomials.py:479-480
        assert g.tolist() == [random.randrange(Q) - Q // 2 for _ in range(d)], s
    # sparse branch (Fisher-Yates, :455-459)
    got = host(ctx.sample_coef([7, 8], norm_bound=3, weight_bound=5))
    for s, g in zip([7, 8], got):
        assert np.array_equal(g, oracle.sample_coef_poly(secpar, s, norm_bound=3, weight_bound=5))
    # reference setup KATs (fusion_setup_KAT_{128,256}.csv) and the goldens
    ks = [int(x) for x in kat_golden[f"setup_seeds_{secpar}"]]
    assert np.array_equal(host(ctx.sample_ntt(ks)), kat_golden[f"setup_poly_{secpar}"])
    assert np.array_equal(host(ctx.sample_ntt([42]))[0], scheme_golden[f"A_{secpar}"][0])


@pytest.mark.parametrize("secpar", [128, 256])
def test_keygen_sign_hash_ch(ctxs, oracle, scheme_golden, golden_json, kat_golden, secpar):
    ctx = ctxs[secpar]
    g, J, t = scheme_golden, golden_json, str(secpar)
    A = dev(g[f"A_{t}"])
    seeds = g[f"seeds_{t}"]
    sk_l, sk_r, sk_hat, vk_l, vk_r = ctx.keygen(A, seeds)
    assert np.array_equal(host(vk_l), g[f"vk_l_{t}"]) and np.array_equal(host(vk_r), g[f"vk_r_{t}"])
    assert np.array_equal(host(sk_hat[:, 0]), g[f"sk_l_row0_{t}"]) and np.array_equal(host(sk_hat[:, 1]), g[f"sk_r_row0_{t}"])
    assert (host(sk_l) == host(sk_hat[:, 0])[:, None, :]).all() and (host(sk_r) == host(sk_hat[:, 1])[:, None, :]).all()
    # general (distinct-row) matvec gives the same keys from the materialised sk
    assert torch.equal(ctx.matvec(A, sk_l), vk_l)
    msgs = [m.encode() for m in J[f"msgs_{t}"]]
    dg = ctx.hash_msg(msgs)
    c_hat = ctx.hash_ch(vk_l, vk_r, dg)
    assert np.array_equal(host(c_hat), g[f"c_hat_{t}"])
    assert np.array_equal(host(ctx.ntt_inv(c_hat)), g[f"c_coef_{t}"])
    sig = ctx.sign_mac(sk_l, sk_r, c_hat)
    assert np.array_equal(host(sig[:, 0]), g[f"sig_row0_{t}"]) and (host(sig) == host(sig[:, :1])).all()
    assert torch.equal(ctx.sign_mac_rows(sk_hat, c_hat), sig)
    # bigger seeded batch vs the oracle, sk not materialised
    bs = np.arange(5000, 5000 + 2 * 150, 2, dtype=np.uint64)
    _, _, _, bvl, bvr = ctx.keygen(A, bs, want_sk=False)
    _, _, ovl, ovr = oracle.keygen_batch(secpar, g[f"A_{t}"].astype(np.int32), bs, want_sk=False)
    assert np.array_equal(host(bvl), ovl) and np.array_equal(host(bvr), ovr)
    if secpar == 128:  # reference KAT intermediate_hash_ch_KAT_128.csv
        kl, kr = kat_golden["kat_ch_vk_l_128"], kat_golden["kat_ch_vk_r_128"]
        kd = ctx.hash_msg([m.encode() for m in J["kat_ch_msgs_128"]])
        assert np.array_equal(host(ctx.hash_ch(dev(kl), dev(kr), kd)), kat_golden["kat_ch_c_hat_128"])


@pytest.mark.parametrize("secpar", [128, 256])
def test_hash_ch_random_keys(ctxs, oracle, secpar):
    ctx = ctxs[secpar]
    d = ctx.d
    rng = np.random.default_rng(99)
    n = 200
    vl = rng.integers(-H, H + 1, size=(n, d)).astype(np.int32)
    vr = rng.integers(-H, H + 1, size=(n, d)).astype(np.int32)
    vl[0] = -H; vr[0] = -H       # longest message (all 11-character values)
    vl[1] = 0; vr[1] = 7         # shortest message
    msgs = [f"m{i}".encode() * (i % 7 + 1) for i in range(n)]
    got = host(ctx.hash_ch(dev(vl), dev(vr), ctx.hash_msg(msgs)))
    for i in range(n):
        assert np.array_equal(got[i], oracle.hash_ch(secpar, vl[i], vr[i], msgs[i])), i


@pytest.mark.parametrize("secpar", [128, 256])
def test_sort_keys(ctxs, oracle, secpar):
    ctx = ctxs[secpar]
    d = ctx.d
    rng = np.random.default_rng(5)
    counts = [1, 2, 17, 300]
    n = sum(counts)
    vl = rng.integers(-H, H + 1, size=(n, d)).astype(np.int32)
    vr = rng.integers(-H, H + 1, size=(n, d)).astype(np.int32)
    # engineered ties / prefix cases: equal first values, numbers that are string prefixes of each other, sign cases
    base = 20
    vl[base:base + 12, 0] = [5, 55, 5, -5, -55, 0, 50, 500, 49, 6, -6, 5]
    vl[base + 2, 1:] = vl[base, 1:]; vr[base + 2] = vr[base]; vr[base + 2, d - 1] = 7; vr[base, d - 1] = 71   # differ only in the LAST value (']' terminator)
    vl[base + 11] = vl[base]; vr[base + 11] = vr[base]                                                      # exact duplicate key (stability)
    vl[base + 12:base + 40, 0] = 123456                                                                      # many equal first values
    vl[base + 12:base + 40, 1] = rng.integers(-30, 30, size=28)
    perm = host(ctx.sort_keys(dev(vl), dev(vr), counts))
    off = 0
    for c in counts:
        want = oracle.sort_by_repr(secpar, vl[off:off + c], vr[off:off + c]) + off
        assert perm[off:off + c].tolist() == want.tolist()
        off += c
    # sel indirection (sliding windows over a key pool)
    sel = torch.from_numpy(rng.integers(0, n, size=n).astype(np.int32))
    perm = host(ctx.sort_keys(dev(vl), dev(vr), counts, sel=sel))
    off = 0
    s = sel.numpy()
    for c in counts:
        want = oracle.sort_by_repr(secpar, vl[s[off:off + c]], vr[s[off:off + c]]) + off
        assert perm[off:off + c].tolist() == want.tolist()
        off += c


@pytest.mark.parametrize("secpar", [128, 256])
def test_hash_ag_aggregate_verify(ctxs, oracle, scheme_golden, golden_json, secpar):
    ctx = ctxs[secpar]
    g, J, t = scheme_golden, golden_json, str(secpar)
    d, ell = ctx.d, ctx.ell
    A = dev(g[f"A_{t}"])
    seeds = g[f"seeds_{t}"]
    nk = len(seeds)
    msgs = [m.encode() for m in J[f"msgs_{t}"]]
    sk_l, sk_r, sk_hat, vk_l, vk_r = ctx.keygen(A, seeds)
    dg = ctx.hash_msg(msgs)
    sig = ctx.sign_mac_rows(sk_hat, ctx.hash_ch(vk_l, vk_r, dg))
    # two aggregates in one batch: all keys, and the first key alone
    sel = torch.tensor(list(range(nk)) + [0], dtype=torch.int32)
    counts = [nk, 1]
    perm = ctx.sort_keys(vk_l, vk_r, counts, sel=sel)
    assert host(perm)[:nk].tolist() == g[f"order_{t}"].tolist()
    dg_pos = dg[sel.cuda().long()]
    vl_s, vr_s, dg_s = ctx.gather_sorted(vk_l, vk_r, dg_pos, perm, sel=sel)
    c_s = ctx.hash_ch(vl_s, vr_s, dg_s)
    alpha = ctx.hash_ag(vl_s, vr_s, dg_s, c_s, counts)
    assert np.array_equal(host(alpha[:nk]), g[f"alpha_hat_sorted_{t}"])
    assert np.array_equal(host(ctx.ntt_inv(alpha[:nk])), g[f"alpha_coef_sorted_{t}"])
    sig_pos = sig[sel.cuda().long()]
    agg = ctx.aggregate_sum(sig_pos, perm, alpha, counts)
    assert np.array_equal(host(agg[0]), g[f"agg_{t}"])
    assert np.array_equal(host(agg[1][0]), g[f"agg1_row0_{t}"])
    codes = host(ctx.verify_check(A, vl_s, vr_s, c_s, alpha, agg, counts))
    assert codes.tolist() == [0, 0]
    bad = agg.clone()
    bad[0, 0, 0] += 1      # tests/test_fusion.py:861-873
    assert host(ctx.verify_check(A, vl_s, vr_s, c_s, alpha, bad, counts)).tolist() == [3, 0]
    assert J[f"verify_tamper_{t}"] == [False, "Target doesn't match image of aggregate signature."]
    # norm branch: an "aggregate" that matches the target image but is huge -- take agg + a kernel element?
    # cheaper: make target == observed trivially with all-zero vk/agg and check the norm test on its own
    z = torch.zeros_like(vl_s)
    big = torch.zeros_like(agg)
    assert host(ctx.verify_check(A, z, z, c_s, alpha, big, counts)).tolist() == [0, 0]


@pytest.mark.parametrize("secpar", [128, 256])
def test_aggregate_many_vs_oracle(ctxs, oracle, scheme_golden, secpar):
    """A few hundred signers in several aggregates (split-K path and multi-aggregate path) vs the oracle."""
    ctx = ctxs[secpar]
    d, ell = ctx.d, ctx.ell
    A_np = scheme_golden[f"A_{secpar}"].astype(np.int32)
    A = dev(A_np)
    nk = 96
    seeds = np.arange(10**6, 10**6 + 2 * nk, 2, dtype=np.uint64)
    _, _, sk_hat, vk_l, vk_r = ctx.keygen(A, seeds, want_sk=False)
    msgs = [f"msg-{i:07d}".encode() for i in range(nk)]
    dg = ctx.hash_msg(msgs)
    sig = ctx.sign_mac_rows(sk_hat, ctx.hash_ch(vk_l, vk_r, dg))
    counts = [nk, 5, 40]
    sel = torch.tensor(list(range(nk)) + list(range(10, 15)) + list(range(50, 90)), dtype=torch.int32)
    perm = ctx.sort_keys(vk_l, vk_r, counts, sel=sel)
    vl_s, vr_s, dg_s = ctx.gather_sorted(vk_l, vk_r, dg[sel.cuda().long()], perm, sel=sel)
    c_s = ctx.hash_ch(vl_s, vr_s, dg_s)
    alpha = ctx.hash_ag(vl_s, vr_s, dg_s, c_s, counts)
    agg = ctx.aggregate_sum(sig[sel.cuda().long()], perm, alpha, counts)
    hvl, hvr, hsig = host(vk_l), host(vk_r), host(sig)
    off = 0
    s = sel.numpy()
    oracle.set_threads(oracle.max_threads())
    for j, c in enumerate(counts):
        idx = s[off:off + c]
        want = oracle.aggregate(secpar, hvl[idx], hvr[idx], [msgs[i] for i in idx], hsig[idx])
        assert np.array_equal(host(agg[j]), want), j
        assert oracle.verify(secpar, A_np, hvl[idx], hvr[idx], [msgs[i] for i in idx], want) == 0
        off += c
    assert host(ctx.verify_check(A, vl_s, vr_s, c_s, alpha, agg, counts)).tolist() == [0, 0, 0]
    # fold of partial aggregates == aggregate of the union (the multi-GPU reduce)
    half = nk // 2
    perm0 = perm[:nk]
    part_counts = [half, nk - half]
    parts = ctx.aggregate_sum(sig[sel.cuda().long()], perm0, alpha[:nk], part_counts)
    folded = ctx.fold_partials(parts.reshape(1, 2, ell, d), 2, 1)
    assert torch.equal(folded[0], agg[0])


@pytest.mark.parametrize("secpar", [128, 256])
def test_buffered_and_streaming_hash_agree(ctxs, oracle, secpar):
    """Two implementations of hash_ch / hash_ag (text staged in HBM vs generated inside the sponge threads) give the
    same bytes, incl. extreme text lengths and several aggregates of different sizes (also an empty-free ragged mix)."""
    ctx = ctxs[secpar]
    d = ctx.d
    rng = np.random.default_rng(17)
    counts = [1, 3, 40, 2, 65]
    n = sum(counts)
    vl = rng.integers(-H, H + 1, size=(n, d)).astype(np.int32)
    vr = rng.integers(-H, H + 1, size=(n, d)).astype(np.int32)
    vl[0] = -H; vr[0] = -H       # longest text
    vl[1] = 0; vr[1] = 3         # shortest text
    vl[2] = rng.integers(-1000, 1000, size=d)
    dg = rng.integers(0, 256, size=(n, 32)).astype(np.uint8)
    dg[3] = 0
    dg[4, 1:] = 0
    out = {}
    try:
        for mode, impl in ((1, 0), (0, 2), (0, 3)):   # HBM-staged text / streaming thread per aggregate / streaming warp per aggregate
            ctx.set_option("buffered_hash", mode)
            ctx.set_option("hash_ag_impl", impl)
            c = ctx.hash_ch(dev(vl), dev(vr), dev(dg, np.uint8))
            a = ctx.hash_ag(dev(vl), dev(vr), dev(dg, np.uint8), c, counts)
            out[(mode, impl)] = (host(c), host(a))
    finally:
        ctx.set_option("buffered_hash", 0)
        ctx.set_option("hash_ag_impl", 0)
    # fourth implementation: the ring-fed kernels (text by a warp per signer, chains that only permute, parallel decode), with a warp
    # with a thread and with two lanes (a lane pair) per aggregate
    try:
        for impl in (3, 2, 4):
            ctx.set_option("hash_ag_impl", impl)
            c = ctx.hash_ch(dev(vl), dev(vr), dev(dg, np.uint8))
            out[("ring", impl)] = (host(c), host(ctx.hash_ag_fast(dev(vl), dev(vr), dev(dg, np.uint8), c, counts)))
    finally:
        ctx.set_option("hash_ag_impl", 0)
    ref = out[(1, 0)]
    for k, v in out.items():
        assert np.array_equal(v[0], ref[0]) and np.array_equal(v[1], ref[1]), k
    out = {0: out[(0, 3)], 1: ref}
    # and against the oracle for the first aggregates (messages are irrelevant here: digests are given)
    for i in range(4):
        xof = oracle.hash_vk_and_int_to_bytes(secpar, vl[i], vr[i], bytes(dg[i]), ctx.param("n_ch"))
        want = oracle.ntt_fwd(secpar, oracle.decode_poly(secpar, xof, secpar, 1, ctx.param("omega_ch")))
        assert np.array_equal(out[0][0][i], want)


@pytest.mark.parametrize("secpar", [128, 256])
@pytest.mark.parametrize("impl", [3, 2, 4])
def test_hash_pipeline_stays_inside_its_byte_buffers(ctxs, secpar, impl):
    """compute-sanitizer is not available on this pool, so the byte rings and XOF buffers of the ring-fed aggregation hash (fz_hash_ag_text,
    fz_hash_ag_absorb_ring, fz_hash_ag_squeeze_xof, fz_hash_ag_decode) are carved out of the middle of larger buffers whose head and tail hold
    guard bytes: the guards must survive, the state records must stay inside their array, and the coefficients must equal the one-shot hash's.
    Extreme text lengths (all coefficients -(q-1)/2: the longest repr; all zero: the shortest) sit first and last in their aggregates."""
    ctx = ctxs[secpar]
    d = ctx.d
    rng = np.random.default_rng(5 + secpar + impl)
    counts = [1, 7, 2, 19]
    n, n_aggs = sum(counts), len(counts)
    vl = rng.integers(-H, H + 1, size=(n, d)).astype(np.int32)
    vr = rng.integers(-H, H + 1, size=(n, d)).astype(np.int32)
    vl[0] = vr[0] = -H
    vl[-1] = vr[-1] = 0
    vl[1] = vr[8] = -H
    dg = rng.integers(0, 256, size=(n, 32)).astype(np.uint8)
    dg[0] = 255
    dg[-1] = 0
    tvl, tvr, tdg = dev(vl), dev(vr), dev(dg, np.uint8)
    try:
        ctx.set_option("hash_ag_impl", impl)
        c = ctx.hash_ch(tvl, tvr, tdg)
        want = ctx.hash_ag(tvl, tvr, tdg, c, counts)
        agg_off, n_total, max_keys = ctx._agg_off(counts)
        agg_of = torch.repeat_interleave(torch.arange(n_aggs, dtype=torch.int32, device="cuda"), torch.tensor(counts, device="cuda"))
        GB = 4096
        ring0, cap, text_pos = ctx.hash_ag_ring(n_aggs, max_keys)
        xof0, xcap = ctx.hash_ag_xof_buffer(n_aggs, max_keys)
        rbuf = torch.full((GB + ring0.numel() + GB,), 0xA5, dtype=torch.uint8, device="cuda")
        xbuf = torch.full((GB + xof0.numel() + GB,), 0xA5, dtype=torch.uint8, device="cuda")
        ring, xof = rbuf[GB:GB + ring0.numel()], xbuf[GB:GB + xof0.numel()]
        sb = int(ctx.lib.fz_hash_ag_state_bytes(ctx._h))
        sbuf = torch.full((GB + n_aggs * sb + GB,), 0xA5, dtype=torch.uint8, device="cuda")
        state = sbuf[GB:GB + n_aggs * sb]
        ctx._chk(ctx.lib.fz_hash_ag_begin(ctx._h, state.data_ptr(), agg_off.data_ptr(), n_aggs, ctx._stream()))
        wend = ctx.hash_ag_text(tvl, tvr, tdg, c, agg_off, agg_of, agg_off, n_aggs, n_total, 0, text_pos, ring, cap)
        ctx.hash_ag_absorb_ring(state, ring, cap, wend, agg_off, n_aggs)
        bpos = ctx.hash_ag_squeeze_xof(state, agg_off, n_aggs, xof, xcap)
        obuf = torch.full((GB + n_total * d + GB,), 0x5A5A5A5A, dtype=torch.int32, device="cuda")
        out = obuf[GB:GB + n_total * d].view(n_total, d)
        ctx._chk(ctx.lib.fz_hash_ag_decode(ctx._h, xof.data_ptr(), xcap, bpos.data_ptr(), agg_off.data_ptr(), agg_of.data_ptr(), n_total, out.data_ptr(),
                                           ctx._stream()))
        torch.cuda.synchronize()
        for name, buf, size in (("ring", rbuf, ring0.numel()), ("xof", xbuf, xof0.numel()), ("state", sbuf, n_aggs * sb)):
            assert bool((buf[:GB] == 0xA5).all()) and bool((buf[GB + size:] == 0xA5).all()), f"{name}: a kernel wrote outside its buffer"
        g = 0x5A5A5A5A
        assert bool((obuf[:GB] == g).all()) and bool((obuf[GB + n_total * d:] == g).all()), "decode wrote outside its output"
        assert torch.equal(out, want)
    finally:
        ctx.set_option("hash_ag_impl", 0)


@pytest.mark.parametrize("impl", [0, 2, 4])
def test_text_ring_overrun_is_detected_not_hashed(ctxs, impl):
    """Ring discipline of the windowed aggregation hash (advisor finding): when a window's text would overwrite bytes a sponge has not absorbed
    yet, the writer flags the aggregate (bit 62 of wend), the absorb call poisons its sponge and the host sees a sponge that never closed --
    the hash of overwritten bytes is never produced.  Provoked with a ring of the minimum size against a window of several signers."""
    ctx = ctxs[256]
    d = ctx.d
    rng = np.random.default_rng(11)
    counts = [6, 1]
    n, n_aggs = sum(counts), len(counts)
    vl, vr = dev(rng.integers(-H, H + 1, size=(n, d)).astype(np.int32)), dev(rng.integers(-H, H + 1, size=(n, d)).astype(np.int32))
    dg = dev(rng.integers(0, 256, size=(n, 32)).astype(np.uint8), np.uint8)
    c = ctx.hash_ch(vl, vr, dg)
    agg_off, n_total, max_keys = ctx._agg_off(counts)
    agg_of = torch.repeat_interleave(torch.arange(n_aggs, dtype=torch.int32, device="cuda"), torch.tensor(counts, device="cuda"))
    cap = 272 * 80                                   # 21,760 bytes: room for two signers' text (about 9.5 KB each), not for six
    ring = torch.zeros(n_aggs * cap, dtype=torch.uint8, device="cuda")
    text_pos = torch.zeros(n_aggs, dtype=torch.int64, device="cuda")
    state = ctx.hash_ag_begin(agg_off, n_aggs)
    wend = ctx.hash_ag_text(vl, vr, dg, c, agg_off, agg_of, agg_off, n_aggs, n_total, 0, text_pos, ring, cap, state=state)
    flags = (wend.cpu().numpy().astype(np.uint64) >> np.uint64(62)) & np.uint64(3)
    assert flags.tolist() == [3, 2], flags           # aggregate 0: complete + overrun; aggregate 1 (one signer): complete, fits
    try:
        ctx.set_option("hash_ag_impl", impl)      # a warp, a thread, a lane pair per aggregate: each must refuse the overrun ring
        ctx.hash_ag_absorb_ring(state, ring, cap, wend, agg_off, n_aggs)
    finally:
        ctx.set_option("hash_ag_impl", 0)
    assert ctx.hash_ag_open_sponges(state, n_aggs) == 1
    # without the state the writer cannot know (the caller's discipline, as before): no flag
    text_pos.zero_()
    wend2 = ctx.hash_ag_text(vl, vr, dg, c, agg_off, agg_of, agg_off, n_aggs, n_total, 0, text_pos, ring, cap)
    assert ((wend2.cpu().numpy().astype(np.uint64) >> np.uint64(62)) & np.uint64(1)).tolist() == [0, 0]
