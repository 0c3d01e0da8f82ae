"""The general ring kernels (fz_ring_*: any odd prime modulus below 2^31, degree 2..16384, caller-supplied tables) and the host
mirror of algebra.ntt / algebra.polynomials on rings other than the scheme's: against vectors recorded from the unmodified
reference (tests/golden/make_ring_golden.py), against the CPU oracle on larger ragged batches, and through the properties the
reference's own tests check (tests/test_ntt.py, tests/test_polynomials.py: round trips, the convolution theorem, str formats,
the exceptions)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).cuda()


def host(t):
    return t.cpu().numpy()


def rings(G):
    return [(i, int(d), int(q)) for i, (d, q) in enumerate(G["rings"])]


def test_ring_kernels_match_reference_vectors(ring_golden):
    from fusion_cryptography_b200 import native

    G = ring_golden
    for idx, d, q in rings(G):
        tw, itw = [int(x) for x in G[f"r{idx}_tw"]], [int(x) for x in G[f"r{idx}_itw"]]
        r = native.get_ring(q, d, tw, itw)
        a, b = dev(G[f"r{idx}_a"]), dev(G[f"r{idx}_b"])
        assert np.array_equal(host(r.ntt_fwd(a)), G[f"r{idx}_fwd"]), (d, q)
        assert np.array_equal(host(r.ntt_inv(b)), G[f"r{idx}_inv"]), (d, q)
        assert np.array_equal(host(r.pointwise_mul(a, b)), G[f"r{idx}_mul"]), (d, q)
        assert np.array_equal(host(r.pointwise_add(a, b)), G[f"r{idx}_add"]), (d, q)
        assert np.array_equal(host(r.polymul(a, b)), G[f"r{idx}_prod"]), (d, q)
        assert np.array_equal(host(r.ntt_inv(r.ntt_fwd(a))), G[f"r{idx}_a"]), (d, q)
        x = a.clone()
        r.ntt_fwd(x, out=x)   # in place
        assert np.array_equal(host(x), G[f"r{idx}_fwd"]), (d, q)


@pytest.mark.parametrize("d,q,n", [(2, 5, 1000), (4, 17, 4099), (64, 257, 3001), (64, 2147483137, 2050), (128, 2147465729, 777), (256, 7681, 1025),
                                   (1024, 12289, 131), (4096, 2013265921, 37), (8192, 2013265921, 5), (16384, 2013265921, 3)])
def test_ring_kernels_match_oracle_on_ragged_batches(oracle, d, q, n):
    """tile edges: n is never a multiple of the polynomials a CTA tile holds; any int32 input (not only canonical residues)"""
    from fusion_cryptography_b200 import native

    def root_of(order):
        g = 2
        while True:
            x = pow(g, (q - 1) // order, q)
            if pow(x, order // 2, q) != 1:
                return x
            g += 1

    root = root_of(2 * d)
    inv_root = pow(root, q - 2, q)

    def table(w):
        bits = d.bit_length() - 1
        p = [pow(w, i, q) for i in range(d)]
        return [p[int(format(i, f"0{bits}b")[::-1], 2)] for i in range(d)]

    tw, itw = table(root), table(inv_root)
    r = native.get_ring(q, d, tw, itw)
    rng = np.random.default_rng(d * 31 + n)
    a = rng.integers(-(2**31), 2**31, size=(n, d)).astype(np.int32)
    b = rng.integers(-(q // 2), q // 2 + 1, size=(n, d)).astype(np.int32)
    fa = host(r.ntt_fwd(dev(a)))
    assert np.array_equal(fa, oracle.ring_ntt_fwd(q, tw, a))
    assert np.array_equal(host(r.ntt_inv(dev(a))), oracle.ring_ntt_inv(q, itw, a))
    assert np.array_equal(host(r.pointwise_mul(dev(a), dev(b))), oracle.ring_pointwise(q, a, b, True))
    assert np.array_equal(host(r.pointwise_add(dev(a), dev(b))), oracle.ring_pointwise(q, a, b, False))
    # operands that do not start on a 16-byte boundary take the scalar load / store path
    buf = torch.zeros(2 * n * d + 3, dtype=torch.int32, device="cuda")
    ua, ub = buf[1:1 + n * d].view(n, d), buf[n * d + 2:2 * n * d + 2].view(n, d)
    ua.copy_(dev(a)), ub.copy_(dev(b))
    assert ua.data_ptr() % 16 != 0
    assert np.array_equal(host(r.ntt_fwd(ua)), fa)
    assert np.array_equal(host(r.pointwise_mul(ua, ub)), oracle.ring_pointwise(q, a, b, True))
    r.ntt_inv(ua, out=ua)
    assert np.array_equal(host(ua), oracle.ring_ntt_inv(q, itw, a))
    # INTT(NTT(a)) is the canonical centred form of a
    h = q // 2
    assert np.array_equal(host(r.ntt_inv(dev(fa))), ((a.astype(np.int64) + h) % q - h).astype(np.int32))
    # convolution theorem against the negacyclic schoolbook product on one row (the reference's tests/test_ntt.py:397-429)
    if d <= 256:
        x, y = [int(v) for v in b[0]], [int(v) for v in b[n - 1]]
        want = [0] * d
        for i in range(d):
            for j in range(d):
                k = i + j
                want[k % d] += x[i] * y[j] * (-1 if k >= d else 1)
        want = [((v + h) % q) - h for v in want]
        assert host(r.polymul(dev(b[0:1]), dev(b[n - 1:n])))[0].tolist() == want


def test_scheme_ring_through_general_kernels_equals_scheme_kernels(ntt_golden):
    """the scheme's own (q, d, table) through fz_ring_* gives the bits of the compile-time kernels"""
    from fusion_cryptography_b200 import native

    for secpar in (128, 256):
        ctx = native.get_context(secpar)
        tw, itw = [int(x) for x in ntt_golden[f"tw_{secpar}"]], [int(x) for x in ntt_golden[f"itw_{secpar}"]]
        r = native.Ring(native.Q, ctx.d, tw, itw)
        a = dev(ntt_golden[f"a_{secpar}"].astype(np.int32))
        assert torch.equal(r.ntt_fwd(a), ctx.ntt_fwd(a)) and torch.equal(r.ntt_inv(a), ctx.ntt_inv(a))
        assert np.array_equal(host(r.ntt_fwd(a)), ntt_golden[f"fwd_{secpar}"])
        r.close()


def test_ring_argument_errors():
    from fusion_cryptography_b200 import native

    with pytest.raises(NotImplementedError):
        native.Ring(2**31 + 11, 4, [1, 1, 1, 1], None)      # modulus does not fit the 32-bit lazy arithmetic
    with pytest.raises(NotImplementedError):
        native.Ring(17, 6, [1] * 6, None)                   # degree is no power of two
    with pytest.raises(NotImplementedError):
        native.Ring(2013265921, 32768, [1] * 32768, None)   # above the largest shared-memory tile
    with pytest.raises(IndexError):
        native.Ring(17, 8, [1, 2, 3], None)                 # the reference's table[m + i] would raise the same
    r = native.Ring(17, 4, [1, 13, 9, 15], None)
    with pytest.raises(native.FusionB200Error):
        r.ntt_inv(torch.zeros(4, dtype=torch.int32))        # created without an inverse table
    with pytest.raises(ValueError):
        r.ntt_fwd(torch.zeros(6, dtype=torch.int32))        # not a whole number of polynomials


# ---------------------------------------------------------------------------------------------- host mirror
def test_algebra_ntt_functions_on_general_rings(ring_golden):
    """cooley_tukey_ntt / gentleman_sande_intt / ntt_poly_mult with the reference's signatures: in place, return the same list,
    f and g of ntt_poly_mult left in canonical centred form (algebra/ntt.py:291, :377, :472-483)."""
    from fusion_cryptography_b200.algebra import ntt

    G = ring_golden
    for idx, d, q in rings(G):
        if d > 2048:
            continue
        root, inv_root = (int(x) for x in G[f"r{idx}_root"])
        tw, itw = [int(x) for x in G[f"r{idx}_tw"]], [int(x) for x in G[f"r{idx}_itw"]]
        row = G[f"r{idx}_a"].shape[0] - 1
        val = [int(x) for x in G[f"r{idx}_a"][row]]
        big = [v + q * (7 * i - 3) for i, v in enumerate(val)]     # Python ints far outside int32: only the residue matters
        ret = ntt.cooley_tukey_ntt(big, q, 2 * d, tw)
        assert ret is big and big == [int(x) for x in G[f"r{idx}_fwd"][row]]
        assert ntt.gentleman_sande_intt(big, q, 2 * d, itw) == val
        f, g = [int(x) + q for x in G[f"r{idx}_a"][row]], [int(x) for x in G[f"r{idx}_b"][row]]
        assert ntt.ntt_poly_mult(f, g, q, root, inv_root, 2 * d) == [int(x) for x in G[f"r{idx}_prod"][row]]
        assert f == [int(x) for x in G[f"r{idx}_a"][row]]
    with pytest.raises(NotImplementedError):
        ntt.cooley_tukey_ntt([1, 2, 3, 4], 17, 4, [1, 4, 13, 16])      # root_order == degree (algebra/ntt.py:267-270)
    with pytest.raises(ValueError):
        ntt.cooley_tukey_ntt([1, 2, 3, 4], 15, 8, [1, 4, 13, 16])      # composite modulus
    with pytest.raises(ValueError):
        ntt.cooley_tukey_ntt([1, 2, 3], 17, 8, [1, 4, 13, 16])         # length is no power of two
    with pytest.raises(ValueError):
        ntt.gentleman_sande_intt([1, 2, 3, 4], 17, 16, [1, 4, 13, 16])  # root_order neither d nor 2d
    with pytest.raises(TypeError):
        ntt.cooley_tukey_ntt([1, 2, 3, 4.0], 17, 8, [1, 4, 13, 16])


def test_polynomial_classes_on_general_rings(ring_golden):
    """transform / + / * / == / str of the polynomial classes on small rings (the reference's tests/test_polynomials.py run on
    moduli below 2^17): values recorded from the reference, formats as its tests spell them (:229, :250)."""
    from fusion_cryptography_b200.algebra.polynomials import (PolynomialCoefficientRepresentation as Coef,
                                                             PolynomialNTTRepresentation as Ntt, transform)

    G = ring_golden
    for idx, d, q in rings(G):
        if d > 256:
            continue
        root, inv_root = (int(x) for x in G[f"r{idx}_root"])
        kw = dict(modulus=q, degree=d, root=root, inv_root=inv_root, root_order=2 * d)
        row = G[f"r{idx}_a"].shape[0] - 1
        a, b = [int(x) for x in G[f"r{idx}_a"][row]], [int(x) for x in G[f"r{idx}_b"][row]]
        ca, cb = Coef(coefficients=list(a), **kw), Coef(coefficients=list(b), **kw)
        ta = transform(ca)
        assert isinstance(ta, Ntt) and ta.values == [int(x) for x in G[f"r{idx}_fwd"][row]]
        assert transform(ta) == ca and transform(ta).coefficients == a
        assert str(ta) == (f"PolynomialNTTRepresentation(modulus={q}, degree={d}, root={root}, inv_root={inv_root}, root_order={2 * d}, "
                           f"values={[int(x) for x in G[f'r{idx}_fwd'][row]]})")
        assert (ca * cb).coefficients == [int(x) for x in G[f"r{idx}_prod"][row]]
        assert transform(transform(ca) * transform(cb)) == ca * cb                      # the convolution theorem, tests/test_polynomials.py:792-880
        na, nb = Ntt(values=list(a), **kw), Ntt(values=list(b), **kw)
        assert (na * nb).values == [int(x) for x in G[f"r{idx}_mul"][row]]
        assert (na + nb).values == [int(x) for x in G[f"r{idx}_add"][row]]
        assert (ca + cb).coefficients == [int(x) for x in G[f"r{idx}_add"][row]]
        assert ca.coefficients == a and na.values == a                                   # operands untouched
    with pytest.raises(NotImplementedError):
        transform(Coef(modulus=17, degree=4, root=4, inv_root=13, root_order=4, coefficients=[1, 2, 3, 4]))   # root_order == degree


def _py_sample_coef(d, q, norm_bound, weight_bound, seed):
    """algebra/polynomials.py:447-459 restated on CPython's own `random` (the reference's only third-party arithmetic here)"""
    import random

    rnd = random.Random()
    rnd.seed(seed)
    num, bound = max(0, min(d, weight_bound)), max(0, min(q // 2, norm_bound))
    c = [(1 + rnd.randrange(bound)) * (1 - 2 * rnd.randrange(2)) for _ in range(num)] + [0] * (d - num)
    if num < d:
        for i in range(d - 1, 0, -1):
            j = rnd.randrange(i + 1)
            c[i], c[j] = c[j], c[i]
    return c


def test_ring_samplers(ring_golden):
    """Seeded samplers on rings other than the scheme's: vectors recorded from the reference (incl. the case its own
    tests/test_polynomials.py:883-909 draws), the state CPython's global generator is left in, and a sweep against
    CPython's `random` for degrees 2..2048, sparse and dense weights, multi-limb and negative seeds."""
    import random

    from fusion_cryptography_b200 import native
    from fusion_cryptography_b200.algebra.polynomials import (sample_polynomial_coefficient_representation as sample_coef,
                                                             sample_polynomial_ntt_representation as sample_ntt)

    G = ring_golden
    for i, (d, q, nb, wb, _) in enumerate(G["sample_cases"]):
        d, q, nb, wb, seed = int(d), int(q), int(nb), int(wb), int(G["sample_seeds"][i])
        root, inv_root = (int(x) for x in G[f"s{i}_root"])
        kw = dict(modulus=q, degree=d, root=root, inv_root=inv_root, root_order=2 * d)
        f = sample_coef(norm_bound=nb, weight_bound=wb, seed=seed, **kw)
        after = [random.getrandbits(32) for _ in range(4)]
        assert f.coefficients == [int(x) for x in G[f"s{i}_coef"]], (d, q)
        assert after == [int(x) for x in G[f"s{i}_after_coef"]]
        g = sample_ntt(seed=seed, **kw)
        after = [random.getrandbits(32) for _ in range(4)]
        assert g.values == [int(x) for x in G[f"s{i}_ntt"]], (d, q)
        assert after == [int(x) for x in G[f"s{i}_after_ntt"]]
        assert f.norm(p="infty") <= nb and f.weight() <= wb
    rng = random.Random(5)
    for d, q in [(2, 5), (16, 97), (64, 257), (128, 65537), (2048, 12289), (512, 2013265921)]:
        r = native.get_ring(q, d)
        seeds = [rng.randrange(1 << rng.choice([8, 31, 32, 33, 64, 200])) * rng.choice([1, -1]) for _ in range(7)]
        for nb, wb in [(1, d), (q, d), (3, d // 2), (1000, 1), (2, 0), (52, d + 5)]:
            got = r.sample_coef(seeds, norm_bound=max(1, min(q // 2, nb)), weight_bound=wb).cpu().tolist()
            assert got == [_py_sample_coef(d, q, nb, wb, s) for s in seeds], (d, q, nb, wb)
        got = r.sample_ntt(seeds).cpu().tolist()
        want = []
        for s in seeds:
            rnd = random.Random()
            rnd.seed(s)
            want.append([rnd.randrange(q) - q // 2 for _ in range(d)])
        assert got == want, (d, q)


def test_general_matrix_on_small_rings():
    """GeneralMatrix + / * over polynomials of small rings, as the reference's tests/test_matrices.py:21-260 exercises them (2 x 2
    matrices of coefficient-representation polynomials on moduli below 2^17), against a plain schoolbook negacyclic product."""
    import random

    from fusion_cryptography_b200.algebra.matrices import GeneralMatrix
    from fusion_cryptography_b200.algebra.ntt import find_primitive_root
    from fusion_cryptography_b200.algebra.polynomials import (PolynomialCoefficientRepresentation as Poly,
                                                             PolynomialNTTRepresentation as PolyNTT, transform)

    rng = random.Random(11)

    def school(x, y, d, q):
        out = [0] * d
        for i, xi in enumerate(x):
            for j, yj in enumerate(y):
                k = i + j
                out[k % d] += xi * yj * (-1 if k >= d else 1)
        return [v % q for v in out]

    for d, q in [(4, 17), (8, 97), (16, 193), (64, 257)]:
        root = find_primitive_root(modulus=q, root_order=2 * d)
        inv_root = pow(root, q - 2, q)
        kw = dict(modulus=q, degree=d, root=root, inv_root=inv_root, root_order=2 * d)
        L = [[[rng.randrange(q) for _ in range(d)] for _ in range(2)] for _ in range(2)]
        R = [[[rng.randrange(q) for _ in range(d)] for _ in range(2)] for _ in range(2)]
        ml = GeneralMatrix(matrix=[[Poly(coefficients=list(c), **kw) for c in row] for row in L])
        mr = GeneralMatrix(matrix=[[Poly(coefficients=list(c), **kw) for c in row] for row in R])
        prod, ssum = ml * mr, ml + mr
        for i in range(2):
            for j in range(2):
                want = [(a + b) % q for a, b in zip(school(L[i][0], R[0][j], d, q), school(L[i][1], R[1][j], d, q))]
                assert [c % q for c in prod.matrix[i][j].coefficients] == want, (d, q, i, j)
                assert [c % q for c in ssum.matrix[i][j].coefficients] == [(a + b) % q for a, b in zip(L[i][j], R[i][j])]
                assert all(-(q // 2) <= c <= q // 2 for c in prod.matrix[i][j].coefficients)
        # the same product through the NTT representation (entries transformed one by one), then back
        nl = GeneralMatrix(matrix=[[transform(p) for p in row] for row in ml.matrix])
        nr = GeneralMatrix(matrix=[[transform(p) for p in row] for row in mr.matrix])
        nprod = nl * nr
        assert nprod.elem_class is PolyNTT
        for i in range(2):
            for j in range(2):
                assert transform(nprod.matrix[i][j]) == prod.matrix[i][j]
        # matrix * element
        e = transform(Poly(coefficients=[1] + [0] * (d - 1), **kw))   # the constant polynomial 1
        assert (nl * e) == nl


def test_decoder_on_general_rings(ring_golden):
    """decode_bytes_to_polynomial_coefficients (fusion/fusion.py:422-481) for degrees and weights beyond the scheme's: vectors recorded from the
    reference for all-zero, all-one and random byte strings, full-length and cut short inside the shuffle -- among them the
    case of the reference's own tests/test_fusion.py:475-558 -- and its ValueError for strings below the length check."""
    from fusion_cryptography_b200.fusion import fusion as ff

    G = ring_golden
    for i, (lb, q, d, nb, wb, ln) in enumerate(G["decode_cases"]):
        b = G[f"d{i}_bytes"].tobytes()
        assert len(b) == int(ln)
        got = ff.decode_bytes_to_polynomial_coefficients(b=b, log2_bias=int(lb), modulus=int(q), degree=int(d), norm_bound=int(nb), weight_bound=int(wb))
        assert got == [int(x) for x in G[f"d{i}_coefs"]], (i, int(lb), int(q), int(d), int(nb), int(wb), int(ln))
    with pytest.raises(ValueError, match="Too few bytes to decode polynomial. Expected 6813 but got 10"):
        ff.decode_bytes_to_polynomial_coefficients(b=bytes(10), log2_bias=256, modulus=65537, degree=1024, norm_bound=1000, weight_bound=100)


@pytest.mark.parametrize("d,q,n", [(2, 5, 1027), (8, 17, 333), (64, 257, 1001), (256, 7681, 67), (1024, 12289, 9), (8192, 2013265921, 3)])
def test_ring_kernels_stay_inside_their_buffers(d, q, n):
    """Guard words before and after every output (and a poisoned input copy) survive the transforms, the pointwise kernels, the samplers and
    the decoder on ragged batch sizes: nothing is written outside [n, d] and the inputs are not modified."""
    from fusion_cryptography_b200 import native

    g = 4096
    r = native.get_ring(q, d, [1] * d, [1] * d)       # any table will do: only the addresses matter here
    rng = np.random.default_rng(d + n)
    a = dev(rng.integers(-(q // 2), q // 2 + 1, size=(n, d)))
    b = dev(rng.integers(-(q // 2), q // 2 + 1, size=(n, d)))
    a0, b0 = a.clone(), b.clone()

    def guarded(fn):
        buf = torch.full((g + n * d + g,), 0x5A5A5A5A, dtype=torch.int32, device="cuda")
        fn(buf[g:g + n * d].view(n, d))
        torch.cuda.synchronize()
        assert bool((buf[:g] == 0x5A5A5A5A).all()) and bool((buf[g + n * d:] == 0x5A5A5A5A).all())
        assert bool((buf[g:g + n * d] != 0x5A5A5A5A).any()) or q < 1000

    guarded(lambda o: r.ntt_fwd(a, out=o))
    guarded(lambda o: r.ntt_inv(a, out=o))
    guarded(lambda o: r.pointwise_mul(a, b, out=o))
    guarded(lambda o: r.pointwise_add(a, b, out=o))
    assert torch.equal(a, a0) and torch.equal(b, b0)
    seeds = list(range(100, 100 + min(n, 40)))
    for sample in (lambda: r.sample_coef(seeds, norm_bound=max(1, min(q // 2, 7)), weight_bound=d // 2), lambda: r.sample_ntt(seeds)):
        out = sample()
        assert out.shape == (len(seeds), d) and int(out.abs().max()) <= q // 2


def test_empty_batches():
    """zero polynomials: every entry point returns an empty result and launches nothing"""
    from fusion_cryptography_b200 import native

    r = native.get_ring(257, 64, [1] * 64, [1] * 64)
    e = torch.empty((0, 64), dtype=torch.int32, device="cuda")
    for out in (r.ntt_fwd(e), r.ntt_inv(e), r.pointwise_mul(e, e), r.pointwise_add(e, e), r.polymul(e, e), r.sample_coef([]), r.sample_ntt([])):
        assert out.shape == (0, 64)
    for secpar in (128, 256):
        ctx = native.get_context(secpar)
        z = torch.empty((0, ctx.d), dtype=torch.int32, device="cuda")
        for out in (ctx.ntt_fwd(z), ctx.ntt_inv(z), ctx.pointwise_mul(z, z), ctx.polymul(z, z), ctx.ntt_pass(z, z)):
            assert out.shape == (0, ctx.d)
