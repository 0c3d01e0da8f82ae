"""A battery of API calls -- ordinary, edge and erroneous -- whose outcomes (sha-256 of ``str(result)``, or the exception's type and text) are printed
as one JSON object.  tests/test_reference_suite_cpu.py runs this file twice, with ``algebra`` / ``fusion`` resolving to the unmodified reference and
to the host mirror (device calls served by the oracle stand-in), and compares the two outputs.  Imports only the reference's public names."""
import hashlib
import json
import random

from algebra.matrices import GeneralMatrix
from algebra.ntt import (bit_reverse_copy, cent, cooley_tukey_ntt, find_primitive_root, gentleman_sande_intt, has_primitive_root_of_unity,
                         is_odd_prime, is_pow_two_geq_two, is_primitive_root, is_root_of_unity, ntt_poly_mult)
from algebra.polynomials import (PolynomialCoefficientRepresentation as Coef, PolynomialNTTRepresentation as Ntt,
                                 sample_polynomial_coefficient_representation, sample_polynomial_ntt_representation, transform)
import fusion.fusion as ff

OUT = {}


def case(name, fn):
    try:
        r = fn()
        OUT[name] = "ok:" + hashlib.sha256(str(r).encode()).hexdigest()
    except Exception as e:  # the reference's behaviour on bad input is part of the interface
        OUT[name] = f"{type(e).__name__}:{e}"


q, d = 257, 8
root = find_primitive_root(modulus=q, root_order=2 * d)
inv_root = pow(root, q - 2, q)
kw = dict(modulus=q, degree=d, root=root, inv_root=inv_root, root_order=2 * d)
tw = bit_reverse_copy([pow(root, i, q) for i in range(d)])
itw = bit_reverse_copy([pow(inv_root, i, q) for i in range(d)])
a, b = [1, -2, 3, 400, 5 * q + 1, -q, 0, 7], [9, 8, 7, 6, 5, 4, 3, -200]

# ---- algebra.ntt
case("cent", lambda: [cent(v, q, q // 2, 9) for v in (-1000, -129, -128, 0, 128, 129, 256, 257, 10**12)])
case("cent_type", lambda: cent(1.5, q, q // 2, 9))
case("cent_modulus", lambda: cent(1, 1, 1, 1))
case("prime", lambda: [is_odd_prime(v) for v in (-3, 0, 1, 2, 3, 9, 257, 2147465729)])
case("pow2", lambda: [is_pow_two_geq_two(v) for v in (0, 1, 2, 3, 4, 1024, 1025)])
case("has_root", lambda: [has_primitive_root_of_unity(257, o) for o in (1, 2, 16, 256, 512, 3)])
case("is_root", lambda: [is_root_of_unity(root, q, 16), is_root_of_unity(root, q, 8), is_primitive_root(root, q, 16), is_primitive_root(root * root % q, q, 16)])
case("find_root", lambda: [find_primitive_root(257, o) for o in (2, 4, 16, 256)] + [find_primitive_root(17, 32)])
case("brv", lambda: bit_reverse_copy(list(range(16))))
case("brv_type", lambda: bit_reverse_copy((1, 2)))
case("ct", lambda: cooley_tukey_ntt(list(a), q, 2 * d, tw))
case("gs", lambda: gentleman_sande_intt(list(b), q, 2 * d, itw))
case("ct_roundtrip", lambda: gentleman_sande_intt(cooley_tukey_ntt(list(a), q, 2 * d, tw), q, 2 * d, itw))
case("ct_long_table", lambda: cooley_tukey_ntt(list(a), q, 2 * d, tw + [5, 6, 7]))
case("ct_short_table", lambda: cooley_tukey_ntt(list(a), q, 2 * d, tw[:5]))
case("ct_not_list", lambda: cooley_tukey_ntt(tuple(a), q, 2 * d, tw))
case("ct_float", lambda: cooley_tukey_ntt([1.0] * d, q, 2 * d, tw))
case("ct_composite", lambda: cooley_tukey_ntt(list(a), 255, 2 * d, tw))
case("ct_no_root", lambda: cooley_tukey_ntt(list(a), 19, 2 * d, tw))
case("ct_len", lambda: cooley_tukey_ntt(list(a)[:6], q, 2 * d, tw))
case("ct_order", lambda: cooley_tukey_ntt(list(a), q, 4 * d, tw))
case("ct_order_eq_len", lambda: cooley_tukey_ntt(list(a), q, d, tw))
case("gs_table_type", lambda: gentleman_sande_intt(list(a), q, 2 * d, [1.0] * d))
case("polymult", lambda: ntt_poly_mult(list(a), list(b), q, root, inv_root, 2 * d))
case("polymult_side_effect", lambda: (lambda f, g: (ntt_poly_mult(f, g, q, root, inv_root, 2 * d), f, g))(list(a), list(b)))
case("polymult_bad_root", lambda: ntt_poly_mult(list(a), list(b), q, root * root % q, inv_root, 2 * d))
case("polymult_bad_inv", lambda: ntt_poly_mult(list(a), list(b), q, root, inv_root + 1, 2 * d))
case("polymult_len", lambda: ntt_poly_mult(list(a), list(b)[:4], q, root, inv_root, 2 * d))

# ---- algebra.polynomials
ca, cb, na, nb = Coef(coefficients=list(a), **kw), Coef(coefficients=list(b), **kw), Ntt(values=list(a), **kw), Ntt(values=list(b), **kw)
case("coef_str", lambda: ca)
case("ntt_str", lambda: na)
case("transform", lambda: transform(ca))
case("transform_back", lambda: transform(transform(ca)))
case("transform_type", lambda: transform([1, 2]))
case("coef_add", lambda: ca + cb)
case("coef_sub", lambda: ca - cb)
case("coef_neg", lambda: -ca)
case("coef_mul", lambda: ca * cb)
case("coef_mul_one", lambda: ca * 1)
case("coef_mul_zero", lambda: ca * 0)
case("coef_radd_zero", lambda: 0 + ca)
case("coef_eq", lambda: [ca == ca, ca == cb, ca == Coef(coefficients=[x + q for x in a], **kw), ca == na, ca == 5])
case("coef_norm_weight", lambda: [ca.norm(p="infty"), ca.weight(), cb.norm(p="infty"), Coef(coefficients=[0] * d, **kw).weight()])
case("coef_norm_p", lambda: ca.norm(p=2))
case("coef_add_mixed", lambda: ca + na)
case("ntt_add", lambda: na + nb)
case("ntt_sub", lambda: na - nb)
case("ntt_neg", lambda: -na)
case("ntt_mul", lambda: na * nb)
case("ntt_rmul_one", lambda: 1 * na)
case("ntt_mul_zero_poly", lambda: na * Ntt(values=[0, q, -q, 0, 0, 0, 0, 2 * q], **kw))
case("ntt_eq", lambda: [na == na, na == nb, na == Ntt(values=[x - q for x in a], **kw), Ntt(values=[0] * d, **kw) == 0, na == 0])
case("ntt_add_other_ring", lambda: na + Ntt(modulus=17, degree=8, root=3, inv_root=6, root_order=16, values=[1] * 8))
case("ntt_mul_other_degree", lambda: na * Ntt(modulus=257, degree=4, root=find_primitive_root(257, 8), inv_root=pow(find_primitive_root(257, 8), 255, 257), root_order=8, values=[1] * 4))
case("ctor_len", lambda: Coef(coefficients=[1, 2, 3], **kw))
case("ctor_type", lambda: Ntt(values=(1,) * d, **kw))
case("ctor_elem_type", lambda: Ntt(values=[1.0] * d, **kw))
case("ctor_root", lambda: Coef(modulus=q, degree=d, root=root * root % q, inv_root=inv_root, root_order=2 * d, coefficients=list(a)))
case("ctor_inv_root", lambda: Coef(modulus=q, degree=d, root=root, inv_root=inv_root + 1, root_order=2 * d, coefficients=list(a)))
case("ctor_order", lambda: Coef(modulus=q, degree=d, root=root, inv_root=inv_root, root_order=24, coefficients=list(a)))
case("ctor_modulus_type", lambda: Coef(modulus=257.0, degree=d, root=root, inv_root=inv_root, root_order=2 * d, coefficients=list(a)))
case("sample_coef_seeded", lambda: sample_polynomial_coefficient_representation(norm_bound=5, weight_bound=3, seed=99, **kw))
case("sample_coef_state", lambda: random.getrandbits(64))
case("sample_coef_dense", lambda: sample_polynomial_coefficient_representation(norm_bound=10**9, weight_bound=100, seed=-7, **kw))
case("sample_coef_zero_bound", lambda: sample_polynomial_coefficient_representation(norm_bound=0, weight_bound=3, seed=1, **kw))
case("sample_coef_zero_weight", lambda: sample_polynomial_coefficient_representation(norm_bound=3, weight_bound=0, seed=1, **kw))
case("sample_ntt_seeded", lambda: sample_polynomial_ntt_representation(seed=2**70, **kw))
random.seed(4)
case("sample_coef_unseeded", lambda: sample_polynomial_coefficient_representation(norm_bound=5, weight_bound=3, seed=None, **kw))
case("sample_ntt_unseeded", lambda: sample_polynomial_ntt_representation(seed=None, **kw))

# ---- larger rings (the general-ring kernels on the device)
for dd, qq in ((64, 7681), (1024, 12289), (2048, 65537 * 0 + 12289)):
    rr = find_primitive_root(modulus=qq, root_order=2 * dd)
    kk = dict(modulus=qq, degree=dd, root=rr, inv_root=pow(rr, qq - 2, qq), root_order=2 * dd)
    f = sample_polynomial_coefficient_representation(norm_bound=qq // 2, weight_bound=dd, seed=1000 + dd, **kk)
    g = sample_polynomial_coefficient_representation(norm_bound=3, weight_bound=dd // 3, seed=2000 + dd, **kk)
    h = sample_polynomial_ntt_representation(seed=3000 + dd, **kk)
    case(f"big_sample_{dd}", lambda: (f, g, h))
    case(f"big_transform_{dd}", lambda: (transform(f), transform(h)))
    case(f"big_roundtrip_{dd}", lambda: transform(transform(f)) == f)
    case(f"big_ntt_ops_{dd}", lambda: (transform(f) * transform(g) + h, transform(f) - h))
    if dd <= 1024:
        case(f"big_coef_mul_{dd}", lambda: f * g)
        case(f"big_convolution_{dd}", lambda: transform(transform(f) * transform(g)) == f * g)

# ---- algebra.matrices
M1 = GeneralMatrix(matrix=[[ca, cb], [cb, ca]])
M2 = GeneralMatrix(matrix=[[cb], [ca]])
case("mat_str", lambda: M1)
case("mat_mul", lambda: M1 * M2)
case("mat_add", lambda: M1 + M1)
case("mat_sub", lambda: M1 - M1)
case("mat_mul_elem", lambda: M1 * ca)
case("mat_dim", lambda: M2 * M1)
case("mat_add_dim", lambda: M1 + M2)
case("mat_eq", lambda: [M1 == M1, M1 == M2, M1 == 0, (M1 - M1) == 0])
case("mat_norm_weight", lambda: [M1.norm(p="infty"), M1.weight()])
case("mat_len_iter", lambda: [len(M1), [len(r) for r in M1], M1[0][1] == cb])
case("mat_ragged", lambda: GeneralMatrix(matrix=[[ca, cb], [ca]]))
case("mat_mixed", lambda: GeneralMatrix(matrix=[[ca, na]]))
case("mat_not_list", lambda: GeneralMatrix(matrix=((ca,),)))
case("mat_mul_other_class", lambda: M1 * GeneralMatrix(matrix=[[na], [nb]]))
case("mat_mod", lambda: GeneralMatrix(matrix=[[7, 9], [11, 13]]) % 5)
case("mat_mod_bad", lambda: GeneralMatrix(matrix=[[7]]) % 1)
case("mat_int_mul", lambda: GeneralMatrix(matrix=[[1, 2], [3, 4]]) * GeneralMatrix(matrix=[[5], [6]]))

# ---- fusion.fusion
for secpar in (128, 256):
    t = str(secpar)
    case("setup_" + t, lambda: ff.fusion_setup(secpar, 11))
    p = ff.fusion_setup(secpar, 11)
    keys = [ff.keygen(p, s) for s in (5, 6, 2**64 + 3, -9)]
    case("keygen_" + t, lambda: keys)
    case("keygen_state_" + t, lambda: random.getrandbits(64))
    case("keygen_none_" + t, lambda: ff.keygen(p, None))
    msgs = ["", "m", "é" * 40, "x" * 1000]
    case("prehash_" + t, lambda: [ff.hash_message_to_int(p, m) for m in msgs])
    sigs = [ff.sign(p, k, m) for k, m in zip(keys, msgs)]
    case("sign_" + t, lambda: sigs)
    vks = [k[1] for k in keys]
    case("hash_ch_" + t, lambda: [ff.hash_ch(p, v, m) for v, m in zip(vks, msgs)])
    case("hash_ag_" + t, lambda: ff.hash_ag(p, vks, msgs))
    case("hash_ag_short_" + t, lambda: ff.hash_ag(p, vks, msgs[:2]))
    case("hash_ag_empty_" + t, lambda: ff.hash_ag(p, [], []))
    agg = ff.aggregate(p, vks, msgs, sigs)
    case("aggregate_" + t, lambda: agg)
    case("aggregate_order_" + t, lambda: ff.aggregate(p, vks[::-1], msgs[::-1], sigs[::-1]))
    case("aggregate_short_" + t, lambda: ff.aggregate(p, vks, msgs, sigs[:3]))
    case("aggregate_empty_" + t, lambda: ff.aggregate(p, [], [], []))
    case("aggregate_dup_" + t, lambda: ff.aggregate(p, [vks[0], vks[0]], ["m", "m"], [sigs[0], sigs[0]]))
    case("verify_" + t, lambda: ff.verify(p, vks, msgs, agg))
    case("verify_order_" + t, lambda: ff.verify(p, vks[::-1], msgs[::-1], agg))
    case("verify_wrong_msg_" + t, lambda: ff.verify(p, vks, msgs[::-1], agg))
    case("verify_count_" + t, lambda: ff.verify(p, vks, msgs[:3], agg))
    case("verify_too_many_" + t, lambda: ff.verify(p, vks * 1000, msgs * 1000, agg))
    case("verify_empty_" + t, lambda: ff.verify(p, [], [], agg))
    case("verify_single_" + t, lambda: ff.verify(p, vks[:1], msgs[:1], sigs[0]))
    case("verify_not_agg_" + t, lambda: ff.verify(p, vks[:2], msgs[:2], sigs[0]))
    case("decode_short_" + t, lambda: ff.decode_bytes_to_polynomial_coefficients(b=bytes(10), log2_bias=secpar, modulus=p.modulus, degree=p.degree,
                                                                                 norm_bound=1, weight_bound=p.omega_ch))
    case("parse_short_" + t, lambda: ff.parse_challenge(p, bytes(100)))
    case("params_eq_" + t, lambda: [p == ff.fusion_setup(secpar, 11), p == ff.fusion_setup(secpar, 12)])
p128 = ff.fusion_setup(128, 11)
k128 = ff.keygen(p128, 5)
case("sign_bytes_message", lambda: ff.sign(p128, k128, b"bytes"))
case("prehash_int_message", lambda: ff.hash_message_to_int(p128, 5))
case("keygen_float_seed", lambda: ff.keygen(p128, 1.5))
case("keygen_str_seed", lambda: ff.keygen(p128, "7"))
case("keygen_str_seed_state", lambda: random.getrandbits(64))
case("keygen_bytes_seed", lambda: ff.keygen(p128, b"7"))
case("setup_str_secpar", lambda: sorted(vars(ff.fusion_setup("128", 1))))
case("tamper_then_verify", lambda: (lambda s: (s.signature_hat.matrix[0][0].values.__setitem__(0, s.signature_hat.matrix[0][0].values[0] + 1),
                                               ff.verify(p128, [k128[1]], ["m"], s))[1])(ff.sign(p128, k128, "m")))
case("tamper_seed_attr", lambda: (lambda k: (setattr(k[0], "seed", 77), str(k[0])[:60])[1])(ff.keygen(p128, 5)))
case("challenge_eq", lambda: [ff.hash_ch(p128, k128[1], "m") == ff.hash_ch(p128, k128[1], "m"), ff.hash_ch(p128, k128[1], "m") == ff.hash_ch(p128, k128[1], "n")])
case("xof_zero_bytes", lambda: ff.hash_vk_and_int_to_bytes(p128, k128[1], 5, 0))
case("xof_arbitrary_key", lambda: ff.hash_vk_and_int_to_bytes(p128, "not a key", 5, 40))
case("sample_matrix", lambda: ff.sample_coefficient_matrix(seed=3, modulus=257, degree=8, root_order=16, root=root, inv_root=inv_root, num_rows=2, num_cols=3,
                                                            norm_bound=4, weight_bound=5))
case("sample_ntt_matrix", lambda: ff.sample_ntt_matrix(seed=3, modulus=257, degree=8, root_order=16, root=root, inv_root=inv_root, num_rows=2, num_cols=2))


def big_key(p, seed):
    """a key pair built by hand from secret polynomials with coefficients up to q/2: consistent (vk = A * sk), far outside the scheme's norm bound"""
    mk = lambda sd: ff.sample_coefficient_matrix(seed=sd, modulus=p.modulus, degree=p.degree, root_order=p.root_order, root=p.root, inv_root=p.inv_root,  # noqa: E731
                                                 num_rows=p.num_rows_sk, num_cols=p.num_cols_sk, norm_bound=p.modulus // 2, weight_bound=p.degree)
    hat = lambda m: GeneralMatrix(matrix=[[transform(e) for e in row] for row in m.matrix])  # noqa: E731
    sl, sr = hat(mk(seed)), hat(mk(seed + 1))
    return (ff.OneTimeSigningKey(seed=seed, left_sk_hat=sl, right_sk_hat=sr),
            ff.OneTimeVerificationKey(left_vk_hat=p.public_challenge * sl, right_vk_hat=p.public_challenge * sr))


bk = big_key(p128, 21)
case("handmade_key", lambda: bk[1])
bsig = ff.sign(p128, bk, "big")
case("handmade_sign", lambda: bsig)
bagg = ff.aggregate(p128, [bk[1]], ["big"], [bsig])
case("verify_norm_too_large", lambda: ff.verify(p128, [bk[1]], ["big"], bagg))
case("verify_mixed_norm", lambda: ff.verify(p128, [bk[1], k128[1]], ["big", "m"], ff.aggregate(p128, [bk[1], k128[1]], ["big", "m"], [bsig, ff.sign(p128, k128, "m")])))


def shifted_key(p, vk):
    """the same key with non-canonical representatives: some values + q, some - q, one just above q // 2 (str(key), which the reference sorts and
    hashes, differs from the canonical key's; the values agree modulo q)"""
    def shift(poly):
        v = list(poly.values)
        v[0], v[1], v[2] = v[0] + p.modulus, v[1] - p.modulus, (p.modulus // 2 + 1 if v[2] == -(p.modulus // 2) else v[2])
        return Ntt(modulus=poly.modulus, degree=poly.degree, root=poly.root, inv_root=poly.inv_root, root_order=poly.root_order, values=v)
    return ff.OneTimeVerificationKey(left_vk_hat=GeneralMatrix(matrix=[[shift(vk.left_vk_hat.matrix[0][0])]]),
                                     right_vk_hat=GeneralMatrix(matrix=[[shift(vk.right_vk_hat.matrix[0][0])]]))


sk128 = shifted_key(p128, k128[1])
case("noncanonical_key_str", lambda: sk128)
case("noncanonical_hash_ch", lambda: ff.hash_ch(p128, sk128, "m"))
nsig = ff.sign(p128, (k128[0], sk128), "m")
case("noncanonical_sign", lambda: nsig)
case("noncanonical_hash_ag", lambda: ff.hash_ag(p128, [sk128, k128[1]], ["m", "n"]))
nsig2 = ff.sign(p128, k128, "n")
nagg = ff.aggregate(p128, [sk128, k128[1]], ["m", "n"], [nsig, nsig2])
case("noncanonical_aggregate", lambda: nagg)
case("noncanonical_verify", lambda: ff.verify(p128, [sk128, k128[1]], ["m", "n"], nagg))
case("noncanonical_verify_canonical_keys", lambda: ff.verify(p128, [k128[1], k128[1]], ["m", "n"], nagg))
case("setup_unknown", lambda: sorted(vars(ff.fusion_setup(64, 1))))
case("setup_none", lambda: len(str(ff.fusion_setup(128, None))) > 1000)

print(json.dumps(OUT))
