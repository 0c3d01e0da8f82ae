"""Drop-in for the reference's ``fusion.fusion`` module (fusion/fusion.py): the same public names,
signatures, return types, ``str()`` formats and error behaviour, with the lattice arithmetic, the
samplers, Keccak, the repr formatter and the decoders executed by hand-written sm_100a kernels
(libfusion_b200.so, C ABI in include/fusion_b200.h).  There is no CPU fallback.

    fusion_setup(secpar, seed) -> Params                     reference :294-295
    keygen(params, seed) -> (OneTimeSigningKey, OneTimeVerificationKey)   :338-373
    sign(params, key, message) -> Signature                  :534-557
    aggregate(params, keys, messages, signatures) -> Signature            :655-677
    verify(params, keys, messages, aggregate_signature) -> (bool, str)    :680-728

``sha3_256`` / ``shake_256`` / ``decode_bytes_to_polynomial_coefficients`` are module attributes with
hashlib's / the reference's call shape, backed by the GPU; the single-item helpers
(hash_message_to_int, hash_vk_and_int_to_bytes, parse_challenge, hash_ch) route through them exactly as
the reference does, so the reference's mock-based tests keep working.  sign / aggregate / verify use
the fused batched device pipeline (format -> SHAKE -> decode -> NTT on the device) as long as those
attributes are not patched.
"""
from __future__ import annotations

import random as _random
from math import ceil, log2
from typing import List, Optional, Tuple

try:
    from .. import native as _native
    from ..algebra.matrices import GeneralMatrix
    from ..algebra.polynomials import (
        PolynomialCoefficientRepresentation,
        PolynomialNTTRepresentation,
        sample_polynomial_coefficient_representation,
        sample_polynomial_ntt_representation,
        transform,
    )
except ImportError:  # drop-in layout: `fusion_cryptography_b200/` itself is on sys.path
    import native as _native
    from algebra.matrices import GeneralMatrix
    from algebra.polynomials import (
        PolynomialCoefficientRepresentation,
        PolynomialNTTRepresentation,
        sample_polynomial_coefficient_representation,
        sample_polynomial_ntt_representation,
        transform,
    )

# This implementation is not secure for prod, toy prototype only (as the reference says of itself).

# ------------------------------------------------------------------ parameters (reference :16-141)
PREFIX_PARAMETERS: dict = {}
PRIME: int = 2147465729
_LEVELS = {
    128: dict(degree=2**6, rank=195, capacity=1796, ch_weight=27, ag_weight=35, ch_bd=3, ag_bd=2, root=23584283, dst=1),
    256: dict(degree=2**8, rank=83, capacity=2818, ch_weight=60, ag_weight=60, ch_bd=1, ag_bd=1, root=3337519, dst=3),
}
SK_BD = 52
for _secpar, _L in _LEVELS.items():
    _vf_inter = SK_BD * (1 + min(_L["degree"], _L["ch_weight"]) * _L["ch_bd"])
    _p = {
        "capacity": _L["capacity"],
        "modulus": PRIME,
        "degree": _L["degree"],
        "root_order": 2 * _L["degree"],
        "root": _L["root"],
        "inv_root": pow(_L["root"], PRIME - 2, PRIME),
        "num_rows_pub_challenge": 1,
        "num_rows_sk": _L["rank"],
        "num_rows_vk": 1,
        "num_cols_pub_challenge": _L["rank"],
        "num_cols_sk": 1,
        "num_cols_vk": 1,
        "sign_pre_hash_dst": bytes([_L["dst"], 0]),
        "sign_hash_dst": bytes([_L["dst"], 1]),
        "agg_xof_dst": bytes([_L["dst"], 2]),
        "beta_sk": SK_BD,
        "beta_ch": 1,
        "beta_ag": 1,
        "omega_sk": _L["degree"],
        "omega_ch": _L["ch_weight"],
        "omega_ag": _L["ag_weight"],
        "beta_vf": _L["capacity"] * min(_L["degree"], _L["ag_weight"]) * _L["ag_bd"] * _vf_inter,
        "omega_vf": _L["degree"],
    }
    _p["bytes_for_one_coef_bdd_by_beta_ch"] = ceil((ceil(log2(2 * _p["beta_ch"] + 1) / 8) + _secpar / 8))
    _p["bytes_for_one_coef_bdd_by_beta_ag"] = ceil((ceil(log2(2 * _p["beta_ag"] + 1) / 8) + _secpar / 8))
    _p["bytes_for_poly_shuffle"] = _p["degree"] * ceil((ceil(log2(_p["degree"]) / 8) + _secpar / 8))
    PREFIX_PARAMETERS[_secpar] = _p
    # the reference's per-level module constants (fusion/fusion.py:18-68), under their public names
    globals().update({
        f"DEGREE_{_secpar}": _L["degree"], f"ROOT_ORDER_{_secpar}": 2 * _L["degree"], f"RANK_{_secpar}": _L["rank"], f"CAPACITY_{_secpar}": _L["capacity"],
        f"CH_WEIGHT_{_secpar}": _L["ch_weight"], f"AG_WEIGHT_{_secpar}": _L["ag_weight"], f"SK_BD_{_secpar}": SK_BD, f"CH_BD_{_secpar}": _L["ch_bd"],
        f"AG_BD_{_secpar}": _L["ag_bd"], f"ROOT_{_secpar}": _L["root"], f"SIGN_PRE_HASH_DST_{_secpar}": _p["sign_pre_hash_dst"],
        f"SIGN_HASH_DST_{_secpar}": _p["sign_hash_dst"], f"AGG_XOF_DST_{_secpar}": _p["agg_xof_dst"], f"VF_BD_INTERMEDIATE_{_secpar}": _vf_inter,
        f"VF_BD_{_secpar}": _p["beta_vf"],
    })


# ------------------------------------------------------------------ GPU-backed hashlib look-alikes
class _GpuHash:
    _suffix_shake = False

    def __init__(self, data: bytes = b""):
        self._data = bytes(data)

    def update(self, data: bytes) -> None:
        self._data += bytes(data)

    def _ctx(self):
        return _native.get_context(256)


class sha3_256(_GpuHash):
    """hashlib.sha3_256 call shape; the digest is computed by the Keccak kernel (fz_sha3_256)."""

    def digest(self) -> bytes:
        return bytes(self._ctx().sha3_256([self._data])[0].cpu().numpy().tobytes())

    def hexdigest(self) -> str:
        return self.digest().hex()


class shake_256(_GpuHash):
    """hashlib.shake_256 call shape; the XOF output is computed by the Keccak kernel (fz_shake256)."""

    def digest(self, length: int) -> bytes:
        if length == 0:
            return b""
        return bytes(self._ctx().shake256([self._data], length)[0].cpu().numpy().tobytes())

    def hexdigest(self, length: int) -> str:
        return self.digest(length).hex()


_PRISTINE = {"sha3_256": sha3_256, "shake_256": shake_256}


# ------------------------------------------------------------------ samplers (reference :144-201)
def _ring_of(modulus, degree, root_order, root, inv_root):
    return (modulus, degree, root, inv_root, root_order)


def sample_coefficient_matrix(seed: Optional[int], modulus: int, degree: int, root_order: int, root: int, inv_root: int, num_rows: int,
                              num_cols: int, norm_bound: int, weight_bound: int) -> GeneralMatrix:
    """Every entry is sampled with the SAME seed (reference :159-168): with a seed the matrix holds
    num_rows x num_cols copies of one polynomial, sampled once by the device MT19937 kernel."""
    if seed is None or num_rows * num_cols == 0:
        return GeneralMatrix(matrix=[[sample_polynomial_coefficient_representation(
            modulus=modulus, degree=degree, root_order=root_order, root=root, inv_root=inv_root, norm_bound=norm_bound,
            weight_bound=weight_bound, seed=seed) for _ in range(num_cols)] for _ in range(num_rows)])
    one = sample_polynomial_coefficient_representation(modulus=modulus, degree=degree, root_order=root_order, root=root, inv_root=inv_root,
                                                       norm_bound=norm_bound, weight_bound=weight_bound, seed=seed)
    t = one._tensor().reshape(1, 1, degree).expand(num_rows, num_cols, degree).contiguous()
    return GeneralMatrix._from_tensor(one._ring(), PolynomialCoefficientRepresentation, t)


def sample_ntt_matrix(seed: Optional[int], modulus: int, degree: int, root_order: int, root: int, inv_root: int, num_rows: int,
                      num_cols: int) -> GeneralMatrix:
    if seed is None or num_rows * num_cols == 0:
        return GeneralMatrix(matrix=[[sample_polynomial_ntt_representation(
            modulus=modulus, degree=degree, root_order=root_order, root=root, inv_root=inv_root, seed=seed) for _ in range(num_cols)]
            for _ in range(num_rows)])
    one = sample_polynomial_ntt_representation(modulus=modulus, degree=degree, root_order=root_order, root=root, inv_root=inv_root, seed=seed)
    t = one._tensor().reshape(1, 1, degree).expand(num_rows, num_cols, degree).contiguous()
    return GeneralMatrix._from_tensor(one._ring(), PolynomialNTTRepresentation, t)


# ------------------------------------------------------------------ Params (reference :204-295)
class Params(object):
    def __init__(self, secpar: int, seed: Optional[int]):
        if secpar in PREFIX_PARAMETERS:
            P = PREFIX_PARAMETERS[secpar]
            self.secpar = secpar
            self.capacity = P["capacity"]
            self.modulus = P["modulus"]
            self.degree = P["degree"]
            self.root_order = P["root_order"]
            self.root = P["root"]
            self.inv_root = P["inv_root"]
            self.num_rows_pub_challenge = P["num_rows_pub_challenge"]
            self.num_rows_sk = P["num_rows_sk"]
            self.num_rows_vk = P["num_rows_vk"]
            self.num_cols_pub_challenge = P["num_cols_pub_challenge"]
            self.num_cols_sk = P["num_cols_sk"]
            self.num_cols_vk = P["num_cols_vk"]
            self.sign_pre_hash_dst = P["sign_pre_hash_dst"]
            self.sign_hash_dst = P["sign_hash_dst"]
            self.agg_xof_dst = P["agg_xof_dst"]
            self.bytes_for_one_coef_bdd_by_beta_ch = P["bytes_for_one_coef_bdd_by_beta_ch"]
            self.bytes_for_one_coef_bdd_by_beta_ag = P["bytes_for_one_coef_bdd_by_beta_ag"]
            self.bytes_for_poly_shuffle = P["bytes_for_poly_shuffle"]
            self.beta_sk = P["beta_sk"]
            self.beta_ch = P["beta_ch"]
            self.beta_ag = P["beta_ag"]
            self.beta_vf = P["beta_vf"]
            self.omega_sk = P["omega_sk"]
            self.omega_ch = P["omega_ch"]
            self.omega_ag = P["omega_ag"]
            self.omega_vf = P["omega_vf"]
            self.public_challenge = sample_ntt_matrix(seed=seed, modulus=self.modulus, degree=self.degree, root_order=self.root_order,
                                                      root=self.root, inv_root=self.inv_root, num_rows=self.num_rows_pub_challenge,
                                                      num_cols=self.num_cols_pub_challenge)

    def __str__(self) -> str:
        return (f"Params(secpar={self.secpar}, capacity={self.capacity}, modulus={self.modulus}, degree={self.degree}, "
                f"root_order={self.root_order}, root={self.root}, inv_root={self.inv_root}, "
                f"num_rows_pub_challenge={self.num_rows_pub_challenge}, num_rows_sk={self.num_rows_sk}, num_rows_vk={self.num_rows_vk}, "
                f"num_cols_pub_challenge={self.num_cols_pub_challenge}, num_cols_sk={self.num_cols_sk}, num_cols_vk={self.num_cols_vk}, "
                f"beta_sk={self.beta_sk}, beta_ch={self.beta_ch}, beta_ag={self.beta_ag}, beta_vf={self.beta_vf}, omega_sk={self.omega_sk}, "
                f"omega_ch={self.omega_ch}, omega_ag={self.omega_ag}, omega_vf={self.omega_vf}, public_challenge={str(self.public_challenge)}, "
                f"sign_pre_hash_dst={self.sign_pre_hash_dst}, sign_hash_dst={self.sign_hash_dst}, agg_xof_dst={self.agg_xof_dst}, "
                f"bytes_for_one_coef_bdd_by_beta_ch={self.bytes_for_one_coef_bdd_by_beta_ch}, "
                f"bytes_for_one_coef_bdd_by_beta_ag={self.bytes_for_one_coef_bdd_by_beta_ag}, bytes_for_poly_shuffle={self.bytes_for_poly_shuffle})")

    def __repr__(self) -> str:
        return self.__str__()

    def __eq__(self, other):
        return self.__dict__ == other.__dict__

    __hash__ = None

    # device plumbing
    def _ctx(self):
        return _native.get_context(self.secpar)

    def _ring(self):
        return (self.modulus, self.degree, self.root, self.inv_root, self.root_order)

    def _A(self):
        """[ell, d] device tensor of the public matrix."""
        return self.public_challenge._tensor().reshape(self.num_cols_pub_challenge, self.degree)


def fusion_setup(secpar: int, seed: Optional[int]) -> Params:
    return Params(secpar=secpar, seed=seed)


# ------------------------------------------------------------------ keys and signatures (reference :298-402, 560-570)
class OneTimeSigningKey(object):
    def __init__(self, seed: Optional[int], left_sk_hat: GeneralMatrix, right_sk_hat: GeneralMatrix):
        self.seed = seed
        self.left_sk_hat = left_sk_hat
        self.right_sk_hat = right_sk_hat

    def __str__(self):
        return f"OneTimeSigningKey(seed={self.seed}, left_sk_hat={str(self.left_sk_hat)}, right_sk_hat={str(self.right_sk_hat)})"

    def __repr__(self):
        return self.__str__()


class OneTimeVerificationKey(object):
    def __init__(self, left_vk_hat: GeneralMatrix, right_vk_hat: GeneralMatrix):
        self.left_vk_hat = left_vk_hat
        self.right_vk_hat = right_vk_hat

    def __str__(self):
        return f"OneTimeVerificationKey(left_vk_hat={self.left_vk_hat}, right_vk_hat={self.right_vk_hat})"

    def __repr__(self):
        return self.__str__()


OneTimeKeyTuple = Tuple[OneTimeSigningKey, OneTimeVerificationKey]


class SignatureChallenge(object):
    def __init__(self, c_hat: PolynomialNTTRepresentation):
        self.c_hat = c_hat

    def __str__(self):
        return f"SignatureChallenge(c_hat={str(self.c_hat)})"

    def __repr__(self):
        return self.__str__()

    def __eq__(self, other):
        return self.c_hat == other.c_hat

    __hash__ = None


class Signature(object):
    def __init__(self, signature_hat: GeneralMatrix):
        self.signature_hat = signature_hat

    def __str__(self):
        return f"Signature(signature_hat={str(self.signature_hat)})"

    def __repr__(self):
        return self.__str__()


class AggregationCoefficient(object):
    def __init__(self, alpha_hat: PolynomialNTTRepresentation):
        self.alpha_hat = alpha_hat

    def __str__(self):
        return f"AggregationCoefficient(alpha_hat={self.alpha_hat})"

    def __repr__(self):
        return self.__str__()


# ------------------------------------------------------------------ keygen (reference :338-373)
def keygen(params: Params, seed: Optional[int]) -> OneTimeKeyTuple:
    ctx = params._ctx()
    d, ell = params.degree, params.num_rows_sk
    ring = params._ring()
    if isinstance(seed, (str, bytes, bytearray)):
        # the reference samples the left key from such a seed -- observable only through the caller's global generator -- and then fails at
        # `seed + 1` (:352); the device sampler takes keys up to 256 bits, a str / bytes seed is 512 bits and more, so only the host side is replayed
        from random import seed as _reseed

        _reseed(seed)
        for _ in range(min(d, params.omega_sk)):
            _random.randrange(params.beta_sk)
            _random.randrange(2)
        seed + 1   # TypeError: can only concatenate str (not "int") to str
    if not isinstance(seed, int) or not (0 <= seed < 2**64 - 1):
        # seeds outside the batched kernel's u64 range (negative / 64 bits and more) take the per-polynomial device
        # sampler through the same matrix code path as the reference; seed=None raises TypeError at `seed + 1`
        # after the left key has consumed the caller's generator, exactly like the reference (:352)
        left = sample_coefficient_matrix(seed=seed, modulus=params.modulus, degree=d, root_order=params.root_order, root=params.root,
                                         inv_root=params.inv_root, num_rows=ell, num_cols=params.num_cols_sk, norm_bound=params.beta_sk,
                                         weight_bound=params.omega_sk)
        right = sample_coefficient_matrix(seed=seed + 1, modulus=params.modulus, degree=d, root_order=params.root_order,
                                          root=params.root, inv_root=params.inv_root, num_rows=ell, num_cols=params.num_cols_sk,
                                          norm_bound=params.beta_sk, weight_bound=params.omega_sk)
        sk_l = ctx.ntt_fwd(left._tensor().reshape(ell, d)).reshape(1, ell, d)
        sk_r = ctx.ntt_fwd(right._tensor().reshape(ell, d)).reshape(1, ell, d)
        A = params._A()
        vk_l, vk_r = ctx.matvec(A, sk_l), ctx.matvec(A, sk_r)
    else:
        import numpy as np

        sk_l, sk_r, _, vk_l, vk_r = ctx.keygen(params._A(), np.array([seed], dtype=np.uint64), want_sk=True)
        # leave CPython's global generator where the reference leaves it: seeded with seed + 1 and advanced by one
        # secret-key polynomial (README's demo messages depend on it)
        _random.seed(seed + 1)
        for _ in range(min(d, params.omega_sk)):
            _random.randrange(params.beta_sk)
            _random.randrange(2)
    NTT = PolynomialNTTRepresentation
    sk = OneTimeSigningKey(seed=seed, left_sk_hat=GeneralMatrix._from_tensor(ring, NTT, sk_l.reshape(ell, 1, d)),
                           right_sk_hat=GeneralMatrix._from_tensor(ring, NTT, sk_r.reshape(ell, 1, d)))
    vk = OneTimeVerificationKey(left_vk_hat=GeneralMatrix._from_tensor(ring, NTT, vk_l.reshape(1, 1, d)),
                                right_vk_hat=GeneralMatrix._from_tensor(ring, NTT, vk_r.reshape(1, 1, d)))
    return sk, vk


# ------------------------------------------------------------------ hashing helpers (reference :405-531)
def hash_message_to_int(params: Params, message: str) -> int:
    salted_message = (params.sign_pre_hash_dst.decode("utf-8") + "," + message).encode()
    pre_hashed_message = sha3_256(salted_message).digest()
    return int.from_bytes(pre_hashed_message, byteorder="little")


def hash_vk_and_int_to_bytes(params: Params, key: OneTimeVerificationKey, i: int, n: int) -> bytes:
    x: bytes = (params.sign_hash_dst.decode("utf-8") + "," + str(key) + "," + str(i)).encode("utf-8")
    return shake_256(x).digest(n)


def _decoder_sizes(log2_bias: int, modulus: int, degree: int, norm_bound: int, weight_bound: int):
    bound = max(1, min(modulus // 2, norm_bound))
    bytes_per_coefficient = ceil((log2(bound) + 1 + log2_bias) / 8)
    bytes_per_index = ceil((log2(degree) + log2_bias) / 8)
    bytes_for_signums = ceil(weight_bound / 8)
    return bytes_for_signums + (bytes_per_coefficient + bytes_per_index) * weight_bound


def decode_bytes_to_polynomial_coefficients(b: bytes, log2_bias: int, modulus: int, degree: int, norm_bound: int, weight_bound: int) -> List[int]:
    """Reference :422-481, on the device (fz_decode_poly)."""
    total_bytes = _decoder_sizes(log2_bias, modulus, degree, norm_bound, weight_bound)
    if len(b) < total_bytes:
        raise ValueError(f"Too few bytes to decode polynomial. Expected {total_bytes} but got {len(b)}")
    import torch

    data = torch.frombuffer(bytearray(b), dtype=torch.uint8).reshape(1, -1) if len(b) else torch.zeros((1, 1), dtype=torch.uint8)
    if modulus == PRIME and degree in (64, 256) and 0 <= weight_bound <= 256:
        ctx = _native.get_context(128 if degree == 64 else 256)
        return ctx.decode_poly(data, len(b), log2_bias, norm_bound, weight_bound)[0].cpu().tolist()
    # any other ring: the general decoder (fz_ring_decode_poly)
    if not 0 <= weight_bound <= degree:
        raise NotImplementedError("libfusion_b200 decodes weight bounds between 0 and the degree")
    bound = max(1, min(modulus // 2, norm_bound))
    ring = _native.get_ring(modulus, degree)
    return ring.decode_poly(data, len(b), norm_bound, weight_bound, ceil((log2(bound) + 1 + log2_bias) / 8), ceil((log2(degree) + log2_bias) / 8))[0].cpu().tolist()


_PRISTINE["decode_bytes_to_polynomial_coefficients"] = decode_bytes_to_polynomial_coefficients


def parse_challenge(params: Params, b: bytes) -> PolynomialNTTRepresentation:
    if len(b) < params.omega_ch * params.bytes_for_one_coef_bdd_by_beta_ch + params.bytes_for_poly_shuffle:
        raise ValueError("hashed_vk_and_pre_hashed_message is too short")
    c_coefs: List[int] = decode_bytes_to_polynomial_coefficients(b=b, log2_bias=params.secpar, modulus=params.modulus, degree=params.degree,
                                                                 norm_bound=params.beta_ch, weight_bound=params.omega_ch)
    c = PolynomialCoefficientRepresentation(modulus=params.modulus, degree=params.degree, root=params.root, inv_root=params.inv_root,
                                            root_order=params.root_order, coefficients=c_coefs)
    return transform(c)


def _challenge_bytes(params: Params) -> int:
    num_coefs = max(0, min(params.degree, params.omega_ch))
    bound = max(0, min(params.modulus // 2, params.beta_ch))
    bytes_per_coefficient = ceil((log2(bound) + 1 + params.secpar) / 8)
    bytes_per_index = ceil((log2(params.degree) + params.secpar) / 8)
    bytes_for_signums = ceil(params.omega_ch / 8)
    return bytes_for_signums + bytes_per_coefficient * num_coefs + params.degree * bytes_per_index


def hash_ch(params: Params, key: OneTimeVerificationKey, message: str) -> SignatureChallenge:
    pre_hashed_message: int = hash_message_to_int(params=params, message=message)
    n = _challenge_bytes(params)
    chall_as_bytes: bytes = hash_vk_and_int_to_bytes(params=params, key=key, i=pre_hashed_message, n=n)
    parsed_chall: PolynomialNTTRepresentation = parse_challenge(params=params, b=chall_as_bytes)
    return SignatureChallenge(c_hat=parsed_chall)


def _patched() -> bool:
    g = globals()
    return any(g[k] is not v for k, v in _PRISTINE.items())


def _device_hashable(params: Params, keys) -> bool:
    """False when some verification key holds a HOST list with a value outside the canonical centred range [-(q//2), q//2] (a hand-made key with
    v + q, say): the reference hashes and sorts str(key) as it stands, while the device formatter prints the canonical residue, so such keys take
    the text-level helpers -- as objects that are no verification keys do.  (Device-resident values are canonical by construction.)"""
    h = params.modulus // 2
    for k in keys:
        if not isinstance(k, OneTimeVerificationKey):
            continue
        for gm in (k.left_vk_hat, k.right_vk_hat):
            for row in getattr(gm, "matrix", ()):
                for e in row:
                    vals = getattr(e, "_vals", None)
                    if vals is not None and any(type(v) is not int or v < -h or v > h for v in vals):
                        return False
    return True


def _vk_tensors(params: Params, keys: List[OneTimeVerificationKey]):
    """([n, d], [n, d]) device tensors of the keys' left / right polynomials."""
    import torch

    d = params.degree
    if not keys:
        e = torch.empty((0, d), dtype=torch.int32, device=params._ctx().device)
        return e, e.clone()
    vl = torch.stack([k.left_vk_hat._tensor().reshape(d) for k in keys])
    vr = torch.stack([k.right_vk_hat._tensor().reshape(d) for k in keys])
    return vl, vr


def _encode_messages(params: Params, messages: List[str]) -> List[bytes]:
    for m in messages:
        if not isinstance(m, str):   # what the reference's `dst + "," + message` raises (fusion.py:406)
            raise TypeError(f'can only concatenate str (not "{type(m).__name__}") to str')
    return [m.encode() for m in messages]


# ------------------------------------------------------------------ sign (reference :534-557)
def sign(params: Params, key: OneTimeKeyTuple, message: str) -> Signature:
    sk, vk = key
    ctx = params._ctx()
    d, ell = params.degree, params.num_rows_sk
    if _patched() or not _device_hashable(params, [vk]):
        pre_hashed_message = hash_message_to_int(params=params, message=message)
        b = hash_vk_and_int_to_bytes(params=params, key=vk, i=pre_hashed_message, n=_challenge_bytes(params))
        c_hat_t = parse_challenge(params=params, b=b)._tensor().reshape(1, d)
    else:
        vl, vr = _vk_tensors(params, [vk])
        c_hat_t = ctx.hash_ch(vl, vr, ctx.hash_msg(_encode_messages(params, [message])))
    z = ctx.sign_mac(sk.left_sk_hat._tensor().reshape(1, ell, d), sk.right_sk_hat._tensor().reshape(1, ell, d), c_hat_t)
    return Signature(signature_hat=GeneralMatrix._from_tensor(params._ring(), PolynomialNTTRepresentation, z.reshape(ell, 1, d)))


# ------------------------------------------------------------------ aggregation hash (reference :573-652)
def hash_vks_and_ints_and_challs_to_bytes(params: Params, keys: List[OneTimeVerificationKey], prehashed_messages: List[int],
                                          challenges: List[SignatureChallenge]) -> bytes:
    bound = max(0, min(params.modulus // 2, params.beta_ag))
    bytes_per_coefficient = ceil((log2(bound) + 1 + params.secpar) / 8)
    bytes_per_index = ceil((log2(params.degree) + params.secpar) / 8)
    bytes_for_signums = ceil(params.omega_ag / 8)
    n = len(keys) * (bytes_for_signums + (bytes_per_coefficient + bytes_per_index) * params.omega_ag)
    salted = str.encode(params.agg_xof_dst.decode("utf-8") + "," + str(list(zip(keys, prehashed_messages, challenges))))
    return shake_256(salted).digest(n)


def decode_bytes_to_agg_coefs(params: Params, b: bytes) -> List[AggregationCoefficient]:
    n = _decoder_sizes(params.secpar, params.modulus, params.degree, params.beta_ag, params.omega_ag)
    out: List[AggregationCoefficient] = []
    for i in range(len(b) // n):
        coefs = decode_bytes_to_polynomial_coefficients(b=b[i * n:(i + 1) * n], log2_bias=params.secpar, modulus=params.modulus,
                                                        degree=params.degree, norm_bound=params.beta_ag, weight_bound=params.omega_ag)
        alpha = PolynomialCoefficientRepresentation(modulus=params.modulus, degree=params.degree, root=params.root, inv_root=params.inv_root,
                                                    root_order=params.root_order, coefficients=coefs)
        out.append(AggregationCoefficient(alpha_hat=transform(alpha)))
    return out


def _device_hash_pipeline(params: Params, keys: List[OneTimeVerificationKey], messages: List[str], sort: bool):
    """sort (optional) -> hash_ch -> hash_ag on the device for ONE aggregate.  Returns (perm, vl_s, vr_s, c_s, alpha)."""
    import torch

    ctx = params._ctx()
    n = len(keys)
    vl, vr = _vk_tensors(params, keys)
    dg = ctx.hash_msg(_encode_messages(params, messages))
    perm = ctx.sort_keys(vl, vr, [n]) if sort else torch.arange(n, dtype=torch.int32, device=ctx.device)
    vl_s, vr_s, dg_s = ctx.gather_sorted(vl, vr, dg, perm)
    c_s = ctx.hash_ch(vl_s, vr_s, dg_s)
    alpha = ctx.hash_ag_fast(vl_s, vr_s, dg_s, c_s, [n])
    return perm, vl_s, vr_s, c_s, alpha


def hash_ag(params: Params, keys: List[OneTimeVerificationKey], messages: List[str]) -> List[AggregationCoefficient]:
    """Reference :632-652 (keys are hashed in the order given; aggregate / verify sort first)."""
    # the reference only ever takes str(key) here, so ANY object works as a key (its own KAT generator passes (sk, vk) tuples,
    # KATs/generate_KAT_values.py:120-133): everything that is not a plain verification key goes through the text-level helpers
    if _patched() or len(keys) != len(messages) or not all(isinstance(k, OneTimeVerificationKey) for k in keys) or not _device_hashable(params, keys):
        # (unequal lengths: the reference zips the per-signer lists but sizes the XOF output by len(keys), :636-647)
        pre = [hash_message_to_int(params=params, message=m) for _, m in zip(keys, messages)]
        challs = [hash_ch(params=params, key=k, message=m) for k, m in zip(keys, messages)]
        b = hash_vks_and_ints_and_challs_to_bytes(params=params, keys=keys, prehashed_messages=pre, challenges=challs)
        return decode_bytes_to_agg_coefs(params=params, b=b)
    n = min(len(keys), len(messages))
    if n == 0:
        return []
    _, _, _, _, alpha = _device_hash_pipeline(params, keys[:n], messages[:n], sort=False)
    ring = params._ring()
    return [AggregationCoefficient(alpha_hat=PolynomialNTTRepresentation._from_tensor(ring, alpha[i])) for i in range(n)]


# ------------------------------------------------------------------ aggregate (reference :655-677)
def _sig_tensors(params: Params, signatures: List[Signature]):
    import torch

    d, ell = params.degree, params.num_rows_sk
    return torch.stack([s.signature_hat._tensor().reshape(ell, d) for s in signatures])


def aggregate(params: Params, keys: List[OneTimeVerificationKey], messages: List[str], signatures: List[Signature]) -> Signature:
    n = min(len(keys), len(messages), len(signatures))  # zip() semantics of the reference
    if n == 0:
        raise IndexError("list index out of range")  # the reference indexes sorted_signatures[0]
    keys, messages, signatures = keys[:n], messages[:n], signatures[:n]
    ctx = params._ctx()
    d, ell = params.degree, params.num_rows_sk
    if _patched() or not all(isinstance(k, OneTimeVerificationKey) for k in keys) or not _device_hashable(params, keys):   # like hash_ag: the reference only sorts and hashes str(key) here
        order = sorted(range(n), key=lambda i: str(keys[i]))
        alphas = hash_ag(params=params, keys=[keys[i] for i in order], messages=[messages[i] for i in order])
        import torch

        perm = torch.tensor(order, dtype=torch.int32, device=ctx.device)
        alpha = torch.stack([a.alpha_hat._tensor() for a in alphas])
    else:
        perm, _, _, _, alpha = _device_hash_pipeline(params, keys, messages, sort=True)
    out = ctx.aggregate_sum(_sig_tensors(params, signatures), perm, alpha, [n])
    return Signature(signature_hat=GeneralMatrix._from_tensor(params._ring(), PolynomialNTTRepresentation, out.reshape(ell, 1, d)))


# ------------------------------------------------------------------ verify (reference :680-728)
_REASONS = {
    3: "Target doesn't match image of aggregate signature.",
    4: "Norm of aggregate signature too large.",
    5: "Weight of aggregate signature too large.",
}


def verify(params: Params, keys: List[OneTimeVerificationKey], messages: List[str], aggregate_signature: Signature) -> Tuple[bool, str]:
    if len(keys) > params.capacity:
        return False, f"Too many keys."
    elif len(keys) != len(messages):
        return False, f"Number of keys and messages must be equal."
    n = len(keys)
    if n == 0:
        raise IndexError("list index out of range")  # the reference indexes sorted_vks[0]
    ctx = params._ctx()
    d, ell = params.degree, params.num_rows_sk
    if _patched() or not _device_hashable(params, keys):
        import torch

        order = sorted(range(n), key=lambda i: str(keys[i]))
        skeys, smsgs = [keys[i] for i in order], [messages[i] for i in order]
        challs = [hash_ch(params=params, key=k, message=m) for k, m in zip(skeys, smsgs)]
        alphas = hash_ag(params=params, keys=skeys, messages=smsgs)
        vl_s, vr_s = _vk_tensors(params, skeys)
        c_s = torch.stack([c.c_hat._tensor() for c in challs])
        alpha = torch.stack([a.alpha_hat._tensor() for a in alphas])
    else:
        _, vl_s, vr_s, c_s, alpha = _device_hash_pipeline(params, keys, messages, sort=True)
    agg = aggregate_signature.signature_hat._tensor().reshape(1, ell, d)
    code = int(ctx.verify_check(params._A(), vl_s, vr_s, c_s, alpha, agg, [n])[0].item())
    if code == 0:
        return True, ""
    return False, _REASONS[code]
