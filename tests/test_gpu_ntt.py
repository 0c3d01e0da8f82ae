"""GPU parity tests for K1-K5 (NTT / INTT / pointwise / fused polymul / matvec / sign MAC):
CUDA path through the C ABI vs the CPU oracle and the reference-generated goldens."""
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


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).cuda()


def host(t):
    return t.cpu().numpy()


@pytest.mark.parametrize("secpar", [128, 256])
def test_ntt_golden(ctxs, ntt_golden, secpar):
    ctx = ctxs[secpar]
    a = ntt_golden[f"a_{secpar}"].astype(np.int32)  # includes non-canonical int32 rows
    assert np.array_equal(host(ctx.ntt_fwd(dev(a))), ntt_golden[f"fwd_{secpar}"])
    assert np.array_equal(host(ctx.ntt_inv(dev(a))), ntt_golden[f"inv_{secpar}"])
    b = ntt_golden[f"b_{secpar}"].astype(np.int32)
    assert np.array_equal(host(ctx.pointwise_mul(dev(a), dev(b))),
                          np.array([[((int(x) * int(y) + H) % Q) - H for x, y in zip(r, s)] for r, s in zip(a, b)]))
    ac = np.array([[((int(x) + H) % Q) - H for x in row] for row in a], dtype=np.int32)
    assert np.array_equal(host(ctx.pointwise_mul(dev(ac), dev(b))), ntt_golden[f"mul_{secpar}"])
    fwd = ntt_golden[f"fwd_{secpar}"].astype(np.int32)
    assert np.array_equal(host(ctx.pointwise_muladd(dev(ac), dev(b), dev(fwd))), ntt_golden[f"muladd_{secpar}"])
    assert np.array_equal(host(ctx.polymul(dev(ac[8:12]), dev(b[8:12]))), ntt_golden[f"conv_{secpar}"])


@pytest.mark.parametrize("secpar", [128, 256])
@pytest.mark.parametrize("n", [0, 1, 2, 7, 33, 1000, 4099])
def test_ntt_vs_oracle_ragged(ctxs, oracle, secpar, n):
    ctx = ctxs[secpar]
    d = ctx.d
    rng = np.random.default_rng(1000 + n)
    a = rng.integers(-H, H + 1, size=(n, d)).astype(np.int32)
    b = rng.integers(-(2**31), 2**31, size=(n, d)).astype(np.int32)  # any int32
    if n == 0:
        assert ctx.ntt_fwd(dev(a)).numel() == 0
        return
    for x in (a, b):
        assert np.array_equal(host(ctx.ntt_fwd(dev(x))), oracle.ntt_fwd(secpar, x))
        assert np.array_equal(host(ctx.ntt_inv(dev(x))), oracle.ntt_inv(secpar, x))
    assert np.array_equal(host(ctx.polymul(dev(a), dev(b))), oracle.polymul(secpar, a, b))
    assert np.array_equal(host(ctx.pointwise_mul(dev(a), dev(b))), oracle.pointwise_mul(a, b))
    # in place
    t = dev(a)
    ctx.ntt_fwd(t, out=t)
    assert np.array_equal(host(t), oracle.ntt_fwd(secpar, a))
    ctx.ntt_inv(t, out=t)
    assert np.array_equal(host(t), a)  # INTT(NTT(f)) == f for canonical f


def test_ntt_full_size_properties(ctxs, oracle):
    """BASELINE config[1]: 2^20 secpar-256 polynomials.  Full comparison against the oracle
    (the C oracle handles 2^20 polynomials in seconds) + round trip + linearity."""
    ctx = ctxs[256]
    n, d = 1 << 20, 256
    rng = np.random.default_rng(0xF0510)
    a = rng.integers(-H, H + 1, size=(n, d), dtype=np.int64).astype(np.int32)
    ta = dev(a)
    fa = ctx.ntt_fwd(ta)
    oracle.set_threads(oracle.max_threads())
    want = oracle.ntt_fwd(256, a)
    assert np.array_equal(host(fa), want)
    back = ctx.ntt_inv(fa)
    assert torch.equal(back, ta)
    del want, back
    b = np.random.default_rng(0xF0511).integers(-H, H + 1, size=(n, d), dtype=np.int64).astype(np.int32)
    tb = dev(b)
    fb = ctx.ntt_fwd(tb)
    # linearity: NTT(a + b) == NTT(a) + NTT(b)  (mod q, canonical)
    s = ((ta.to(torch.int64) + tb.to(torch.int64) + H) % Q - H).to(torch.int32)
    fs = ((fa.to(torch.int64) + fb.to(torch.int64) + H) % Q - H).to(torch.int32)
    assert torch.equal(ctx.ntt_fwd(s), fs)
    # fused polymul == unfused chain, and vs oracle on a slice
    pm = ctx.polymul(ta, tb)
    assert torch.equal(pm, ctx.ntt_inv(ctx.pointwise_mul(fa, fb)))
    assert np.array_equal(host(pm[:4096]), oracle.polymul(256, a[:4096], b[:4096]))
    assert np.array_equal(host(pm[-4096:]), oracle.polymul(256, a[-4096:], b[-4096:]))


@pytest.mark.parametrize("secpar", [128, 256])
@pytest.mark.parametrize("batch", [1, 3, 130])
def test_matvec_and_sign_mac(ctxs, oracle, secpar, batch):
    ctx = ctxs[secpar]
    d, ell = ctx.d, ctx.ell
    rng = np.random.default_rng(7 + batch)
    A = rng.integers(-H, H + 1, size=(ell, d)).astype(np.int32)
    s = rng.integers(-H, H + 1, size=(batch, ell, d)).astype(np.int32)
    s[0, :, :4] = H
    A[:, :2] = -H
    assert np.array_equal(host(ctx.matvec(dev(A), dev(s))), oracle.matvec(secpar, A, s))
    s2 = rng.integers(-(2**31), 2**31, size=(batch, ell, d)).astype(np.int32)  # streamed operand: any int32
    assert np.array_equal(host(ctx.matvec(dev(A), dev(s2))), oracle.matvec(secpar, A, s2))
    sk_r = rng.integers(-H, H + 1, size=(batch, ell, d)).astype(np.int32)
    c = rng.integers(-H, H + 1, size=(batch, d)).astype(np.int32)
    want = oracle.pointwise_muladd(s, np.broadcast_to(c[:, None, :], s.shape), sk_r)
    assert np.array_equal(host(ctx.sign_mac(dev(s), dev(sk_r), dev(c))), want)


def test_ntt_host_path(ctxs, oracle):
    ctx = ctxs[256]
    n = 70000  # > one 32 MiB chunk
    rng = np.random.default_rng(3)
    a = rng.integers(-H, H + 1, size=(n, 256)).astype(np.int32)
    hin = torch.from_numpy(a).pin_memory()
    hout = torch.empty_like(hin).pin_memory()
    ctx.ntt_host(hin, hout)
    assert np.array_equal(hout.numpy(), oracle.ntt_fwd(256, a))
    ctx.ntt_host(hout, hin, inverse=True)
    assert np.array_equal(hin.numpy(), a)


@pytest.mark.gpu
@pytest.mark.parametrize("secpar", [128, 256])
def test_fused_pass_equals_three_kernels(secpar):
    """fz_ntt_pass = INTT(NTT(a) o b_hat) in one kernel: the same bits as forward transform, pointwise product, inverse transform
    (algebra/polynomials.py:391-433, :341-385), on a ragged count, out of place and in place, and against the oracle."""
    import torch
    from fusion_cryptography_b200 import native
    from oracle import oracle as orc

    ctx = native.get_context(secpar)
    d = ctx.d
    h = (native.Q - 1) // 2
    rng = np.random.default_rng(secpar + 7)
    n = 4099
    a = rng.integers(-h, h + 1, size=(n, d)).astype(np.int32)
    b = rng.integers(-h, h + 1, size=(n, d)).astype(np.int32)
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    want = ctx.ntt_inv(ctx.pointwise_mul(ctx.ntt_fwd(ta), tb))
    got = ctx.ntt_pass(ta, tb)
    assert torch.equal(got, want)
    inplace = ta.clone()
    ctx.ntt_pass(inplace, tb, out=inplace)
    assert torch.equal(inplace, want)
    k = 64
    assert np.array_equal(got[:k].cpu().numpy(), orc.ntt_inv(secpar, orc.pointwise_mul(orc.ntt_fwd(secpar, a[:k]), b[:k])))


@pytest.mark.gpu
def test_fused_pass_full_size_vs_oracle(ctxs, oracle):
    """The kernel bench.py times, at bench.py's size: fz_ntt_pass over 2^20 secpar-256 polynomials, every coefficient against the oracle's
    cooley_tukey_ntt -> cent(x * y) -> gentleman_sande_intt (algebra/ntt.py:271-291, polynomials.py:341-385, ntt.py:350-377), bench.py's seeds."""
    ctx = ctxs[256]
    n, d = 1 << 20, 256
    a = np.random.default_rng(0xF0510).integers(-H, H + 1, size=(n, d), dtype=np.int64).astype(np.int32)
    b = np.random.default_rng(0xF0511).integers(-H, H + 1, size=(n, d), dtype=np.int64).astype(np.int32)
    got = ctx.ntt_pass(dev(a), dev(b))
    oracle.set_threads(oracle.max_threads())
    want = oracle.ntt_inv(256, oracle.pointwise_mul(oracle.ntt_fwd(256, a), b))
    assert np.array_equal(host(got), want)
    # the host-buffer entry point of the end-to-end measurement gives the same bits
    hin, hout = torch.from_numpy(a).pin_memory(), torch.empty((n, d), dtype=torch.int32).pin_memory()
    ctx.pass_host(hin, dev(b), hout)
    assert np.array_equal(hout.numpy(), want)


@pytest.mark.gpu
@pytest.mark.parametrize("secpar", [128, 256])
@pytest.mark.parametrize("n", [1, 2, 3, 37, 1025])
def test_scheme_ntt_kernels_stay_inside_their_buffers(ctxs, secpar, n):
    """compute-sanitizer is not available on this pool, so: every scheme NTT-family kernel writes into the middle of a larger buffer whose
    head and tail hold guard words; the guards must survive and the payload must equal the result computed in a separate buffer
    (warp-private tiles, ragged last warp / last CTA of the persistent grids)."""
    ctx = ctxs[secpar]
    d, ell = ctx.d, ctx.ell
    rng = np.random.default_rng(n + secpar)
    a = dev(rng.integers(-H, H + 1, size=(n, d)).astype(np.int32))
    b = dev(rng.integers(-H, H + 1, size=(n, d)).astype(np.int32))
    c = dev(rng.integers(-H, H + 1, size=(n, d)).astype(np.int32))
    G, GUARD = 4 * d, 0x5A5A5A5A

    def guarded(fn, rows=n):
        buf = torch.full((G + rows * d + G,), GUARD, dtype=torch.int32, device="cuda")
        out = buf[G:G + rows * d].view(rows, d)
        fn(out)
        assert bool((buf[:G] == GUARD).all()) and bool((buf[G + rows * d:] == GUARD).all()), "a kernel wrote outside its output"
        return out.clone()

    assert torch.equal(guarded(lambda o: ctx.ntt_fwd(a, out=o)), ctx.ntt_fwd(a))
    assert torch.equal(guarded(lambda o: ctx.ntt_inv(a, out=o)), ctx.ntt_inv(a))
    assert torch.equal(guarded(lambda o: ctx.ntt_pass(a, b, out=o)), ctx.ntt_pass(a, b))
    assert torch.equal(guarded(lambda o: ctx.polymul(a, b, out=o)), ctx.polymul(a, b))
    assert torch.equal(guarded(lambda o: ctx.pointwise_mul(a, b, out=o)), ctx.pointwise_mul(a, b))
    assert torch.equal(guarded(lambda o: ctx.pointwise_add(a, b, out=o)), ctx.pointwise_add(a, b))
    assert torch.equal(guarded(lambda o: ctx.pointwise_muladd(a, b, c, out=o)), ctx.pointwise_muladd(a, b, c))
    A = dev(rng.integers(-H, H + 1, size=(ell, d)).astype(np.int32))
    s = dev(rng.integers(-H, H + 1, size=(n, ell, d)).astype(np.int32))
    r = dev(rng.integers(-H, H + 1, size=(n, ell, d)).astype(np.int32))
    assert torch.equal(guarded(lambda o: ctx.matvec(A, s, out=o)), ctx.matvec(A, s))
    assert torch.equal(guarded(lambda o: ctx.sign_mac(s, r, c, out=o.view(n, ell, d)), rows=n * ell).view(n, ell, d), ctx.sign_mac(s, r, c))
