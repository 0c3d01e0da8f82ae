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
    vk_l, vk_r, sigs = eng.keygen_sign(seeds, msgs, chunk=128)   # 3 chunks, last one ragged
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
    # per aggregate (impl 3) and a thread per aggregate (impl 2): all bit-identical
    ref_aggs = None
    try:
        for impl in (3, 2):
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
    os.environ["FZ_AGG_SLOTS"] = "6"
    try:
        for max_pos in (1 << 17, 40, 16):
            aggs = eng.aggregate(kb.vk_l, kb.vk_r, msgs, sigs, counts, sel=sel_t, max_positions=max_pos)
            assert torch.equal(aggs, ref_aggs), ("sliced", max_pos)
            assert host(eng.verify(kb.vk_l, kb.vk_r, msgs, aggs, counts, sel=sel_t, max_positions=max_pos)).tolist() == [0] * len(counts)
    finally:
        del os.environ["FZ_AGG_SLOTS"]
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


def test_capacity_aggregate_properties(oracle):
    """BASELINE config[3] at full size (N = 2818 @256): aggregate and verify one capacity-sized aggregate; properties:
    verify accepts, a tampered copy is rejected, and aggregation is the alpha-weighted sum (linearity in the signatures)."""
    from fusion_cryptography_b200.batch import Engine

    eng = Engine(256, seed=42)
    n = eng.ctx.capacity
    seeds = (10**6 + 2 * np.arange(n)).astype(np.uint64)
    msgs = [f"msg-{i:07d}".encode() for i in range(n)]
    vk_l, vk_r, sigs = eng.keygen_sign(seeds, msgs, chunk=1024)
    agg = eng.aggregate(vk_l, vk_r, msgs, sigs, [n])
    assert host(eng.verify(vk_l, vk_r, msgs, agg, [n])).tolist() == [0]
    # bit-exact against the oracle at full capacity (one SHAKE256 over 26.8 MB of repr text, 11.2 MB squeezed), and the oracle's verify
    oracle.set_threads(oracle.max_threads())
    want = oracle.aggregate(256, host(vk_l), host(vk_r), msgs, host(sigs))
    assert np.array_equal(host(agg[0]), want)
    assert oracle.verify(256, host(eng.A), host(vk_l), host(vk_r), msgs, want) == 0
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
