"""Batched engine (fusion_cryptography_b200.batch) on the GPU: BASELINE configs [2]-[4] at test sizes,
against the CPU oracle; chunking, sliding key windows, tampered aggregates, the single-aggregate
sharded path (world size 1) and size-independent properties at larger sizes."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

Q = 2147465729
H = (Q - 1) // 2


def host(t):
    return t.cpu().numpy()


@pytest.fixture(scope="module")
def engines():
    from fusion_cryptography_b200.batch import Engine

    return {sp: Engine(sp, seed=42) for sp in (128, 256)}


@pytest.mark.parametrize("secpar", [128, 256])
def test_keygen_sign_batch_vs_oracle(engines, oracle, scheme_golden, secpar):
    eng = engines[secpar]
    A = scheme_golden[f"A_{secpar}"].astype(np.int32)
    assert np.array_equal(host(eng.A), A)
    n = 300
    seeds = (10**6 + 2 * np.arange(n)).astype(np.uint64)      # SURVEY C3 seeds (stride 2, quirk Q2)
    msgs = [f"msg-{i:07d}".encode() for i in range(n)]
    vk_l, vk_r, sigs = eng.keygen_sign(seeds, msgs, chunk=128)   # 3 chunks, last one ragged; two-stream pipeline (signature rows beside the next chunk's hash)
    # the same on one stream, chunk sizes that leave 1 / 2 / 5 chunks, and with the signatures of each chunk overwriting alternating buffers
    eng.tuning.sign_overlap = False
    try:
        v2l, v2r, s2 = eng.keygen_sign(seeds, msgs, chunk=128)
    finally:
        eng.tuning.sign_overlap = True
    assert torch.equal(v2l, vk_l) and torch.equal(v2r, vk_r) and torch.equal(s2, sigs)
    for ch in (300, 150, 64, None):
        v3l, v3r, s3 = eng.keygen_sign(seeds, msgs, chunk=ch)
        assert torch.equal(v3l, vk_l) and torch.equal(v3r, vk_r) and torch.equal(s3, sigs), ch
        v4l, v4r, none = eng.keygen_sign(seeds, msgs, chunk=ch, keep_sigs=False)
        assert none is None and torch.equal(v4l, vk_l) and torch.equal(v4r, vk_r), ch
    oracle.set_threads(oracle.max_threads())
    sk_l, sk_r, ovl, ovr = oracle.keygen_batch(secpar, A, seeds)
    assert np.array_equal(host(vk_l), ovl) and np.array_equal(host(vk_r), ovr)
    assert np.array_equal(host(sigs), oracle.sign_batch(secpar, sk_l, sk_r, ovl, ovr, msgs))
    # general kernels on materialised (distinct-row capable) keys give the same signatures
    kb = eng.keygen(seeds[:40], want_sk=True)
    assert torch.equal(eng.sign(kb, msgs[:40]), sigs[:40])
    # sign relation on everything: A . z == vk_l * c + vk_r  (reference tests/test_fusion.py:694-731)
    c = eng.challenges(vk_l, vk_r, msgs)
    lhs = eng.ctx.matvec(eng.A, sigs)
    rhs = eng.ctx.pointwise_muladd(vk_l, c, vk_r)
    assert torch.equal(lhs, rhs)


@pytest.mark.parametrize("secpar", [128, 256])
def test_aggregate_verify_batch(engines, oracle, scheme_golden, secpar):
    eng = engines[secpar]
    A = scheme_golden[f"A_{secpar}"].astype(np.int32)
    pool = 64
    seeds = (2 * 10**6 + 2 * np.arange(pool)).astype(np.uint64)
    kb = eng.keygen(seeds)
    # sliding windows over the key pool, per-aggregate messages
    counts = [1, 2, 9, 33, 17, 5, 64, 12]
    sel, msgs = [], []
    for j, c in enumerate(counts):
        start = (7 * j) % pool
        for i in range(c):
            sel.append((start + i) % pool)
            msgs.append(f"agg{j}-msg{i}".encode())
    sel_t = torch.tensor(sel, dtype=torch.int32)
    # one signature per position: key sel[p] signs msgs[p]
    from fusion_cryptography_b200.batch import KeyBatch
    idx = sel_t.cuda().long()
    kpos = KeyBatch(sk_hat=kb.sk_hat[idx].contiguous(), vk_l=kb.vk_l[idx].contiguous(), vk_r=kb.vk_r[idx].contiguous())
    sigs = eng.sign(kpos, msgs)
    # one window / windows of 5 and of 1 sorted positions per aggregate (the resumable sponge of hash_ag), with a warp
    # per aggregate (impl 3), a thread per aggregate (impl 2) and a lane pair per aggregate (impl 4): all bit-identical
    ref_aggs = None
    try:
        for impl in (3, 2, 4):
            eng.ctx.set_option("hash_ag_impl", impl)
            for max_pos in (1 << 17, 40, 8):
                aggs = eng.aggregate(kb.vk_l, kb.vk_r, msgs, sigs, counts, sel=sel_t, max_positions=max_pos)
                codes = host(eng.verify(kb.vk_l, kb.vk_r, msgs, aggs, counts, sel=sel_t, max_positions=max_pos))
                assert codes.tolist() == [0] * len(counts), (impl, max_pos)
                if ref_aggs is None:
                    ref_aggs = aggs.clone()
                assert torch.equal(aggs, ref_aggs), (impl, max_pos)
    finally:
        eng.ctx.set_option("hash_ag_impl", 0)
    # the time-sliced absorb (batches a little above 8 warp-chains per SM): pretend the GPU has room for 6 chains, so that 6 of the 8
    # aggregates advance per launch with a block budget while the resting pair rotates
    eng.tuning.agg_slots = 6
    try:
        for max_pos in (1 << 17, 40, 16):
            aggs = eng.aggregate(kb.vk_l, kb.vk_r, msgs, sigs, counts, sel=sel_t, max_positions=max_pos)
            assert torch.equal(aggs, ref_aggs), ("sliced", max_pos)
            assert host(eng.verify(kb.vk_l, kb.vk_r, msgs, aggs, counts, sel=sel_t, max_positions=max_pos)).tolist() == [0] * len(counts)
    finally:
        eng.tuning.agg_slots = None
    hvl, hvr, hs = host(kb.vk_l), host(kb.vk_r), host(sigs)
    oracle.set_threads(oracle.max_threads())
    off = 0
    for j, c in enumerate(counts):
        ks = sel[off:off + c]
        want = oracle.aggregate(secpar, hvl[ks], hvr[ks], msgs[off:off + c], hs[off:off + c])
        assert np.array_equal(host(aggs[j]), want), j
        assert oracle.verify(secpar, A, hvl[ks], hvr[ks], msgs[off:off + c], want) == 0
        off += c
    # tampered aggregates (one value + 1) are rejected with the reference's reason code, the others still pass
    bad = aggs.clone()
    bad[2, 0, 0] += 1
    bad[6, eng.ell - 1, eng.d - 1] -= 1
    codes = host(eng.verify(kb.vk_l, kb.vk_r, msgs, bad, counts, sel=sel_t)).tolist()
    assert codes == [0, 0, 3, 0, 0, 0, 3, 0]
    # wrong message for one signer
    wm = list(msgs)
    wm[3] = b"other"
    codes = host(eng.verify(kb.vk_l, kb.vk_r, wm, aggs, counts, sel=sel_t)).tolist()
    assert codes == [0, 0, 3, 0, 0, 0, 0, 0]
    with pytest.raises(IndexError):
        eng.aggregate(kb.vk_l, kb.vk_r, [], sigs, [0])


@pytest.mark.parametrize("secpar", [128, 256])
def test_pair_layout_many_ragged_aggregates(engines, secpar):
    """The lane-pair chain kernels (fz_hashstream.cu pairag::, 16 aggregates per warp, warp-uniform trip counts with per-pair snapshots): 45 aggregates
    of 1..13 signers -- three warps, the last one partly empty, every warp with pairs that finish at different blocks -- over windows of all, 64 and
    16 sorted positions, bit-identical with the warp-per-aggregate kernels; verify accepts all and rejects a tampered one."""
    eng = engines[secpar]
    pool = 64
    seeds = (3 * 10**6 + 2 * np.arange(pool)).astype(np.uint64)
    kb = eng.keygen(seeds)
    rng = np.random.default_rng(secpar)
    counts = [int(c) for c in rng.integers(1, 14, size=45)]
    sel, msgs = [], []
    for j, c in enumerate(counts):
        start = (11 * j) % pool
        for i in range(c):
            sel.append((start + i) % pool)
            msgs.append(f"a{j}-m{i}".encode())
    sel_t = torch.tensor(sel, dtype=torch.int32)
    from fusion_cryptography_b200.batch import KeyBatch
    idx = sel_t.cuda().long()
    kpos = KeyBatch(sk_hat=kb.sk_hat[idx].contiguous(), vk_l=kb.vk_l[idx].contiguous(), vk_r=kb.vk_r[idx].contiguous())
    sigs = eng.sign(kpos, msgs)
    out = {}
    try:
        for impl in (3, 4):
            eng.ctx.set_option("hash_ag_impl", impl)
            for max_pos in (1 << 17, 64, 16):
                aggs = eng.aggregate(kb.vk_l, kb.vk_r, msgs, sigs, counts, sel=sel_t, max_positions=max_pos)
                codes = host(eng.verify(kb.vk_l, kb.vk_r, msgs, aggs, counts, sel=sel_t, max_positions=max_pos))
                assert codes.tolist() == [0] * len(counts), (impl, max_pos)
                out[(impl, max_pos)] = aggs.clone()
        bad = out[(4, 16)].clone()
        bad[17, 0, 3] += 1
        codes = host(eng.verify(kb.vk_l, kb.vk_r, msgs, bad, counts, sel=sel_t, max_positions=16)).tolist()
        assert codes == [3 if j == 17 else 0 for j in range(len(counts))]
    finally:
        eng.ctx.set_option("hash_ag_impl", 0)
    ref = out[(3, 1 << 17)]
    for k, v in out.items():
        assert torch.equal(v, ref), k


def test_chain_layout_is_chosen_by_the_number_of_aggregates(engines):
    """No option set: 1,800 aggregates take the lane-pair chain kernels, 41,000 the thread-per-aggregate ones (fz_hashstream.cu chain_layout); both must
    give what the warp-per-aggregate kernels give.  Aggregates of one and two signers keep the batches small."""
    eng = engines[128]
    pool = 64
    kb = eng.keygen((4 * 10**6 + 2 * np.arange(pool)).astype(np.uint64))
    for n_aggs, per in ((1800, 2), (41000, 1)):
        counts = [per] * n_aggs
        sel = torch.tensor([(5 * j + i) % pool for j in range(n_aggs) for i in range(per)], dtype=torch.int32)
        msgs = [b"m%d" % (p % 97) for p in range(n_aggs * per)]
        from fusion_cryptography_b200.batch import KeyBatch
        idx = sel.cuda().long()
        sigs = eng.sign(KeyBatch(sk_hat=kb.sk_hat[idx].contiguous(), vk_l=kb.vk_l[idx].contiguous(), vk_r=kb.vk_r[idx].contiguous()), msgs)
        auto = eng.aggregate(kb.vk_l, kb.vk_r, msgs, sigs, counts, sel=sel)
        assert bool((eng.verify(kb.vk_l, kb.vk_r, msgs, auto, counts, sel=sel) == 0).all())
        try:
            eng.ctx.set_option("hash_ag_impl", 3)
            ref = eng.aggregate(kb.vk_l, kb.vk_r, msgs, sigs, counts, sel=sel)
        finally:
            eng.ctx.set_option("hash_ag_impl", 0)
        assert torch.equal(auto, ref), n_aggs


def test_single_aggregate_sharded_world1(engines, oracle):
    from fusion_cryptography_b200 import parallel

    eng = engines[256]
    n = 48
    seeds = (3 * 10**6 + 2 * np.arange(n)).astype(np.uint64)
    msgs = [f"m{i}".encode() for i in range(n)]
    vk_l, vk_r, sigs = eng.keygen_sign(seeds, msgs)
    agg = parallel.aggregate_one_sharded(eng, vk_l, vk_r, msgs, sigs)
    want = oracle.aggregate(256, host(vk_l), host(vk_r), msgs, host(sigs))
    assert np.array_equal(host(agg[0]), want)
    # the sharded verification of one aggregate (world size 1 here; world size 2 runs on gloo in tests/test_parallel_cpu.py,
    # world sizes 2-8 on NCCL in bench.py / tools/multi_gpu_check.py)
    assert host(parallel.verify_one_sharded(eng, vk_l, vk_r, msgs, agg)).tolist() == [0]
    assert host(parallel.verify_one_sharded(eng, vk_l, vk_r, msgs, agg)).tolist() == host(eng.verify(vk_l, vk_r, msgs, agg, [n])).tolist()
    bad = agg.clone()
    bad[0, 3, 9] -= 1
    assert host(parallel.verify_one_sharded(eng, vk_l, vk_r, msgs, bad)).tolist() == [3]
    with pytest.raises(ValueError, match="equal"):
        parallel.verify_one_sharded(eng, vk_l, vk_r, msgs[:-1], agg)


@pytest.mark.parametrize("secpar", [256, 128])
def test_capacity_aggregate_properties(oracle, secpar):
    """BASELINE config[3] at full size (N = 2818 @256, N = 1796 @128 -- at 128 the coefficient decoder of fusion.py:594-629 never runs out of
    bytes, the other branch of quirk Q4): aggregate and verify one capacity-sized aggregate, bit-exact against the oracle; properties:
    verify accepts, a tampered copy is rejected, and aggregation is the alpha-weighted sum (linearity in the signatures)."""
    from fusion_cryptography_b200.batch import Engine

    eng = Engine(secpar, seed=42)
    n = eng.ctx.capacity
    assert n == {256: 2818, 128: 1796}[secpar]
    seeds = (10**6 + 2 * np.arange(n)).astype(np.uint64)
    msgs = [f"msg-{i:07d}".encode() for i in range(n)]
    vk_l, vk_r, sigs = eng.keygen_sign(seeds, msgs, chunk=1024)
    agg = eng.aggregate(vk_l, vk_r, msgs, sigs, [n])
    assert host(eng.verify(vk_l, vk_r, msgs, agg, [n])).tolist() == [0]
    # bit-exact against the oracle at full capacity (one SHAKE256 over 26.8 MB of repr text, 11.2 MB squeezed), and the oracle's verify
    oracle.set_threads(oracle.max_threads())
    want = oracle.aggregate(secpar, host(vk_l), host(vk_r), msgs, host(sigs))
    assert np.array_equal(host(agg[0]), want)
    assert oracle.verify(secpar, host(eng.A), host(vk_l), host(vk_r), msgs, want) == 0
    bad = agg.clone()
    bad[0, 5, 7] += 1
    assert host(eng.verify(vk_l, vk_r, msgs, bad, [n])).tolist() == [3]
    # linearity: aggregate(2 * sigs) == 2 * aggregate(sigs)  (mod q)
    two = ((sigs.to(torch.int64) * 2 + H) % Q - H).to(torch.int32)
    agg2 = eng.aggregate(vk_l, vk_r, msgs, two, [n])
    assert torch.equal(agg2, ((agg.to(torch.int64) * 2 + H) % Q - H).to(torch.int32))
    # one more key than the capacity is an argument error (fusion.py:686-687)
    with pytest.raises(ValueError, match="Too many keys"):
        eng.verify(vk_l, vk_r, msgs + [b"x"], agg, [n + 1])


@pytest.mark.parametrize("secpar", [128, 256])
def test_norm_bound_rejection_in_a_batch(engines, oracle, secpar):
    """fusion.py:722-725 through Engine.verify on a batch: an aggregate whose image equals the target but whose coefficient norm exceeds
    beta_vf is rejected with code 4 ("Norm of aggregate signature too large."), its neighbours in the batch are unaffected.
    A kernel vector of A (A . k = 0) with a huge coefficient is added to a valid aggregate: with all columns of A equal (quirk Q1),
    k = (+v, -v, 0, ...) for any polynomial v is one."""
    eng = engines[secpar]
    ctx, d, ell = eng.ctx, eng.d, eng.ell
    n, na = 6, 3
    seeds = (5 * 10**6 + 2 * np.arange(n)).astype(np.uint64)
    msgs = [f"nb-{i}".encode() for i in range(n)] * na
    vk_l, vk_r, sigs = eng.keygen_sign(seeds, msgs[:n])
    sel = torch.arange(n, dtype=torch.int32).repeat(na)
    aggs = eng.aggregate(vk_l, vk_r, msgs, sigs, [n] * na, sel=sel, sig_sel=sel)
    assert host(eng.verify(vk_l, vk_r, msgs, aggs, [n] * na, sel=sel)).tolist() == [0, 0, 0]
    # v: coefficient polynomial with one entry just above beta_vf / 2 in two rows -> after adding, row 0 has a coefficient above beta_vf
    beta = oracle.param(secpar, "beta_vf")
    v = np.zeros((1, d), dtype=np.int32)
    v[0, 0] = H - 7                                     # far above beta_vf (= q / 4)
    v_hat = ctx.ntt_fwd(torch.from_numpy(v).cuda())[0]
    bad = aggs.clone()
    bad[1, 0] = ctx.pointwise_add(bad[1, 0].contiguous(), v_hat)
    neg = ((-(v_hat.to(torch.int64)) + H) % Q - H).to(torch.int32)
    bad[1, 1] = ctx.pointwise_add(bad[1, 1].contiguous(), neg)
    codes = host(eng.verify(vk_l, vk_r, msgs, bad, [n] * na, sel=sel)).tolist()
    assert codes == [0, 4, 0], codes
    hv, hr = host(vk_l), host(vk_r)
    assert oracle.verify(secpar, host(eng.A), hv, hr, msgs[:n], host(bad[1])) == 4
    coef = oracle.ntt_inv(secpar, host(bad[1]))
    assert int(np.abs(coef.astype(np.int64)).max()) > beta


@pytest.mark.parametrize("secpar", [128, 256])
def test_general_row_keygen_sign(engines, oracle, secpar):
    """fz_keygen_rows / Engine.keygen_sign_rows: the general ell-row keygen + sign (2*ell generators, 2*ell transforms, ell-term products; nothing
    of quirk Q1): (i) against vectors the unmodified reference produced row by row (tests/golden/make_rows_golden.py), (ii) against the oracle on
    random row seeds, chunked with a ragged last chunk, (iii) with row seeds (s, s + 1) it IS keygen(params, s), (iv) the sign relation."""
    import os

    eng = engines[secpar]
    ctx, ell, d = eng.ctx, eng.ell, eng.d
    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rows_golden.npz")))
    seeds = g[f"row_seeds_{secpar}"]
    msgs = [str(m).encode() for m in g[f"msgs_{secpar}"]]
    vk_l, vk_r, sigs = eng.keygen_sign_rows(seeds, msgs)
    assert np.array_equal(host(vk_l), g[f"vk_l_{secpar}"]) and np.array_equal(host(vk_r), g[f"vk_r_{secpar}"])
    assert np.array_equal(host(sigs), g[f"sig_{secpar}"])
    # (ii) random row seeds, 7 keys in chunks of 3
    rng = np.random.default_rng(99 + secpar)
    n = 7
    rs = rng.integers(0, 2**63, size=(n, 2, ell), dtype=np.uint64)
    rs[1, 0, :3] = [0, 1, 2**32 - 1]
    m2 = [f"general-{i}".encode() for i in range(n)]
    vl2, vr2, s2 = eng.keygen_sign_rows(rs, m2, chunk=3)
    A = host(eng.A)
    oracle.set_threads(oracle.max_threads())
    for i in range(n):
        sk = [oracle.ntt_fwd(secpar, np.stack([oracle.sample_coef_poly(secpar, int(rs[i, side, k])) for k in range(ell)])) for side in range(2)]
        ovl, ovr = oracle.matvec(secpar, A, sk[0][None])[0], oracle.matvec(secpar, A, sk[1][None])[0]
        assert np.array_equal(host(vl2[i]), ovl) and np.array_equal(host(vr2[i]), ovr), i
        assert np.array_equal(host(s2[i]), oracle.sign(secpar, sk[0], sk[1], ovl, ovr, m2[i])), i
    kb = eng.keygen_rows(rs[:2])
    assert kb.sk_l.shape == (2, ell, d) and not torch.equal(kb.sk_l[0, 0], kb.sk_l[0, 1])      # rows really differ
    # (iii) all rows of side 0 seeded s, of side 1 seeded s + 1: the reference's keygen(params, s), the long way round
    s = np.array([10**6, 10**6 + 2], dtype=np.uint64)
    rs3 = np.stack([np.repeat(s[:, None], ell, 1), np.repeat(s[:, None] + 1, ell, 1)], axis=1)
    k3 = eng.keygen_rows(rs3)
    kq = eng.keygen(s, want_sk=True)
    assert torch.equal(k3.vk_l, kq.vk_l) and torch.equal(k3.vk_r, kq.vk_r) and torch.equal(k3.sk_l, kq.sk_l) and torch.equal(k3.sk_r, kq.sk_r)
    # (iv) A . z == vk_l * c + vk_r  (reference tests/test_fusion.py:694-731)
    c = eng.challenges(vl2, vr2, m2)
    assert torch.equal(ctx.matvec(eng.A, s2), ctx.pointwise_muladd(vl2, c, vr2))
    # general-row signatures aggregate and verify like any others
    agg = eng.aggregate(vl2, vr2, m2, s2, [n])
    assert host(eng.verify(vl2, vr2, m2, agg, [n])).tolist() == [0]
    with pytest.raises(ValueError):
        ctx.keygen_rows(eng.A, rs[:, :, :-1])
