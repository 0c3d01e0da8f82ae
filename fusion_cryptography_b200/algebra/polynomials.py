"""Host mirror of the reference's ``algebra.polynomials`` (same class names, constructor validation,
``str()`` formats, operators and error behaviour), computing on the GPU.

Reference: algebra/polynomials.py.  Objects keep their data either as a device tensor view (when they
were produced by a kernel) or as a plain Python list (when the caller built them, or after the caller
looked at ``.values`` / ``.coefficients`` -- the reference exposes those lists as mutable attributes,
tests/test_fusion.py:861-867 mutates them, so once a list has been handed out it is the source of
truth and is re-uploaded on the next operation).

The class objects report ``__module__ == "algebra.polynomials"``: ``GeneralMatrix.__str__`` embeds
``repr(cls)`` in the text that the scheme hashes, so the qualified name must match
the reference's.
"""
from __future__ import annotations

import random as _random
from copy import deepcopy  # noqa: F401  (public names of the reference's module: algebra/polynomials.py:1-3)
from random import randrange, seed as random_seed  # noqa: F401
from typing import Dict, List, Optional, Tuple, Union

try:
    from .. import native as _native
    from . import ntt as _ntt
except ImportError:  # drop-in layout: imported as top-level `algebra.polynomials`
    import native as _native
    from algebra import ntt as _ntt

cent = _ntt.cent
cooley_tukey_ntt = _ntt.cooley_tukey_ntt
gentleman_sande_intt = _ntt.gentleman_sande_intt
bit_reverse_copy = _ntt.bit_reverse_copy

cached_halfmods: Dict[int, int] = {}
cached_logmods: Dict[int, int] = {}
_VALIDATED: Dict[Tuple[int, int, int, int], Optional[Exception]] = {}

SCHEME_Q = _ntt.SCHEME_Q


def _check_ring(modulus: int, root: int, inv_root: int, root_order: int) -> None:
    """The reference's constructor checks (algebra/polynomials.py:36-45); the verdict per parameter tuple is cached
    (the reference repeats up to root_order - 1 modular exponentiations for every object it creates)."""
    key = (modulus, root, inv_root, root_order)
    if key not in _VALIDATED:
        err: Optional[Exception] = None
        if (modulus - 1) % root_order != 0:
            err = ValueError("root_order must be a divisor of modulus - 1")
        elif pow(root, root_order, modulus) != 1:
            err = ValueError("root must be a root of unity of order root_order")
        else:
            acc = root % modulus
            primitive = True
            for _ in range(1, root_order):
                if acc == 1:
                    primitive = False
                    break
                acc = acc * root % modulus
            if not primitive:
                err = ValueError("root must be a primitive root of unity of order root_order")
            elif (root * inv_root) % modulus != 1:
                err = ValueError("root and inv_root must be inverses of each other")
        _VALIDATED[key] = err
    if _VALIDATED[key] is not None:
        raise type(_VALIDATED[key])(*_VALIDATED[key].args)


def _eq_zero(x) -> bool:
    """`x == 0` with the reference's semantics, without materialising device data as a list"""
    if isinstance(x, PolynomialNTTRepresentation):
        return x._is_zero()
    if isinstance(x, PolynomialCoefficientRepresentation):
        return False
    return x == 0


def _check_same_ring(x, other, cls, op: str) -> None:
    """The reference's operand checks with its exact messages (algebra/polynomials.py:118-135 / :286-303 for +, :175-194 / :346-369 for *)."""
    if not isinstance(other, cls):
        raise NotImplementedError(f"{op} for {type(x)} and {type(other)} not implemented")
    for differs, add_text, mul_text in ((x.modulus != other.modulus, "moduli", "different moduli"),
                                        (x.degree != other.degree, "degrees", "different degrees"),
                                        (x.root != other.root, "roots of unity", "different roots of unity"),
                                        (x.root_order != other.root_order, "root orders", "difference orders of root of unity")):
        if differs:
            raise NotImplementedError(f"Cannot add polynomials with different {add_text}" if op == "Addition"
                                      else f"{op} for {type(x)} with {mul_text} not implemented")


class PolynomialRepresentation(object):
    modulus: int
    degree: int
    root: int
    inv_root: int
    root_order: int

    def __init__(self, modulus: int, degree: int, root: int, inv_root: int, root_order: int):
        if not isinstance(modulus, int):
            raise TypeError("modulus must be an int")
        elif not isinstance(degree, int):
            raise TypeError("degree must be an int")
        elif not isinstance(root, int):
            raise TypeError("root must be an int")
        elif not isinstance(inv_root, int):
            raise TypeError("inv_root must be an int")
        elif not isinstance(root_order, int):
            raise TypeError("root_order must be an int")
        _check_ring(modulus, root, inv_root, root_order)
        self.modulus = modulus
        self.degree = degree
        self.root = root
        self.inv_root = inv_root
        self.root_order = root_order

    @property
    def halfmod(self) -> int:
        if self.modulus not in cached_halfmods:
            cached_halfmods[self.modulus] = self.modulus // 2
        return cached_halfmods[self.modulus]

    @property
    def logmod(self) -> int:
        if self.modulus not in cached_logmods:
            cached_logmods[self.modulus] = self.modulus.bit_length() - 1
        return cached_logmods[self.modulus]

    # ---------------------------------------------------------------- device plumbing
    def _ring(self) -> Tuple[int, int, int, int, int]:
        return (self.modulus, self.degree, self.root, self.inv_root, self.root_order)

    def _secpar(self) -> Optional[int]:
        """the scheme's security level when this is one of its two rings, else None"""
        sc = _ntt._SCHEME.get(self.degree)
        if self.modulus != SCHEME_Q or sc is None or self.root != sc[1] or self.inv_root != sc[2] or self.root_order != 2 * self.degree:
            return None
        return sc[0]

    def _ctx(self):
        """The device engine of this ring: the scheme context (compile-time kernels) for the scheme's two parameter sets, a general
        ring (fz_ring_*: any odd prime modulus below 2^31, power-of-two degree up to 16384) otherwise.  No CPU fallback."""
        secpar = self._secpar()
        if secpar is not None:
            return _native.get_context(secpar)
        return _ntt.ring_for(self.modulus, self.degree, self.root, self.inv_root)

    def _get_list(self) -> List[int]:
        if self._vals is None:
            self._vals = self._t.cpu().tolist()
            self._t = None  # the list may be mutated by the caller from now on
        return self._vals

    def _peek(self) -> List[int]:
        """the stored representatives WITHOUT handing out a mutable reference (device data stays authoritative)"""
        return self._vals if self._vals is not None else self._t.cpu().tolist()

    def _is_zero(self) -> bool:
        if self._vals is not None:
            return all(x % self.modulus == 0 for x in self._vals)
        return not bool(self._t.any())  # kernel outputs are canonical: zero residue <=> zero value

    def _tensor(self):
        """int32 device tensor [degree] of canonical centred values."""
        if self._t is not None and self._vals is None:
            return self._t
        import torch

        h, q = self.modulus // 2, self.modulus
        ctx = self._ctx()
        return torch.tensor([((v + h) % q) - h for v in self._vals], dtype=torch.int32).to(ctx.device)

    @classmethod
    def _from_tensor(cls, ring, t):
        obj = cls.__new__(cls)
        obj.modulus, obj.degree, obj.root, obj.inv_root, obj.root_order = ring
        obj._vals = None
        obj._t = t
        return obj


class PolynomialCoefficientRepresentation(PolynomialRepresentation):
    def __init__(self, modulus: int, degree: int, root: int, inv_root: int, root_order: int, coefficients: List[int]):
        super().__init__(modulus=modulus, degree=degree, root=root, inv_root=inv_root, root_order=root_order)
        if not isinstance(coefficients, list):
            raise TypeError("coefficients must be a list")
        elif not all(isinstance(x, int) for x in coefficients):
            raise TypeError("coefficients must be a list of ints")
        elif len(coefficients) != degree:
            raise ValueError("coefficients must be of length degree")
        self._vals = coefficients
        self._t = None

    @property
    def coefficients(self) -> List[int]:
        return self._get_list()

    @coefficients.setter
    def coefficients(self, v: List[int]) -> None:
        self._vals = v
        self._t = None

    def __str__(self):
        return (f"PolynomialCoefficientRepresentation(modulus={self.modulus}, degree={self.degree}, root={self.root}, "
                f"inv_root={self.inv_root}, root_order={self.root_order}, coefficients={self.coefficients})")

    def __repr__(self):
        return self.__str__()

    def __eq__(self, other):
        if not isinstance(other, PolynomialCoefficientRepresentation):
            return False
        elif self.modulus != other.modulus or self.degree != other.degree or self.root != other.root or self.root_order != other.root_order:
            return False
        return all((x - y) % self.modulus == 0 for x, y in zip(self._peek(), other._peek()))

    def _check_other(self, other, what: str):
        _check_same_ring(self, other, PolynomialCoefficientRepresentation, what)

    def __add__(self, other):
        if _eq_zero(other):
            return self
        self._check_other(other, "Addition")
        out = self._ctx().pointwise_add(self._tensor(), other._tensor())
        return PolynomialCoefficientRepresentation._from_tensor(self._ring(), out)

    def __radd__(self, other):
        if _eq_zero(other):
            return self
        return self + other

    def __neg__(self):
        # the reference returns -(x % q), which is NOT centred; reproduced as is
        return PolynomialCoefficientRepresentation(modulus=self.modulus, degree=self.degree, root=self.root, inv_root=self.inv_root,
                                                   root_order=self.root_order, coefficients=[-(x % self.modulus) for x in self.coefficients])

    def __sub__(self, other):
        return self + (-other)

    def __rsub__(self, other):
        return other + (-self)

    def __mul__(self, other):
        """Product in Z_q[X]/(X^d + 1).  The reference multiplies schoolbook and folds (:196-208); the result equals
        INTT(NTT(a) * NTT(b)) exactly, which is what the fused GPU kernel computes."""
        if _eq_zero(other):
            return 0
        elif other == 1:
            return self
        self._check_other(other, "Multiplication")
        out = self._ctx().polymul(self._tensor(), other._tensor())
        return PolynomialCoefficientRepresentation._from_tensor(self._ring(), out)

    def __rmul__(self, other):
        return self.__mul__(other=other)

    def norm(self, p: Union[int, str]) -> int:
        if p != "infty":
            raise NotImplementedError(f"norm for p={p} not implemented")
        return max(abs(x) for x in self.coefficients)

    def weight(self) -> int:
        return sum(1 if x % self.modulus != 0 else 0 for x in self.coefficients)


class PolynomialNTTRepresentation(PolynomialRepresentation):
    def __init__(self, modulus: int, degree: int, root: int, inv_root: int, root_order: int, values: List[int]):
        super().__init__(modulus=modulus, degree=degree, root=root, inv_root=inv_root, root_order=root_order)
        if not isinstance(values, list):
            raise TypeError("values must be a list")
        elif not all(isinstance(x, int) for x in values):
            raise TypeError("values must be a list of ints")
        elif len(values) != degree:
            raise ValueError("values must have length degree")
        self._vals = values
        self._t = None

    @property
    def values(self) -> List[int]:
        return self._get_list()

    @values.setter
    def values(self, v: List[int]) -> None:
        self._vals = v
        self._t = None

    def __str__(self):
        return (f"PolynomialNTTRepresentation(modulus={self.modulus}, degree={self.degree}, root={self.root}, "
                f"inv_root={self.inv_root}, root_order={self.root_order}, values={self.values})")

    def __repr__(self):
        return self.__str__()

    def __eq__(self, other):
        if _eq_zero(other):
            return self._is_zero()
        elif not isinstance(other, PolynomialNTTRepresentation):
            return False
        elif self.modulus != other.modulus or self.degree != other.degree or self.root_order != other.root_order:
            return False
        elif self.root != other.root or self.inv_root != other.inv_root:
            return False
        a, b = self._peek(), other._peek()
        if len(a) != len(b):
            return False
        return all((x - y) % self.modulus == 0 for x, y in zip(a, b))

    __hash__ = None

    def _check_other(self, other, verb: str):
        _check_same_ring(self, other, PolynomialNTTRepresentation, verb)

    def __add__(self, other):
        if _eq_zero(other):
            return self
        self._check_other(other, "Addition")
        out = self._ctx().pointwise_add(self._tensor(), other._tensor())
        return PolynomialNTTRepresentation._from_tensor(self._ring(), out)

    def __radd__(self, other):
        if _eq_zero(other):
            return self
        return self + other

    def __neg__(self):
        # -(x % q), not centred, exactly as the reference
        return PolynomialNTTRepresentation(modulus=self.modulus, degree=self.degree, root=self.root, inv_root=self.inv_root,
                                           root_order=self.root_order, values=[-(x % self.modulus) for x in self.values])

    def __sub__(self, other):
        return self + (-other)

    def __rsub__(self, other):
        return other + (-self)

    def __mul__(self, other):
        if _eq_zero(other):
            return 0  # includes the all-zero polynomial, like the reference (:342-343 via __eq__(0))
        elif other == 1:
            return self
        self._check_other(other, "Multiplication")
        out = self._ctx().pointwise_mul(self._tensor(), other._tensor())
        return PolynomialNTTRepresentation._from_tensor(self._ring(), out)

    def __rmul__(self, other):
        return self.__mul__(other=other)


PolynomialCoefficientRepresentation.__hash__ = None
for _cls in (PolynomialRepresentation, PolynomialCoefficientRepresentation, PolynomialNTTRepresentation):
    _cls.__module__ = "algebra.polynomials"  # repr(cls) is part of the hashed text


def transform(x: Union[PolynomialCoefficientRepresentation, PolynomialNTTRepresentation]):
    """Coefficient <-> NTT representation (reference :391-433); the input object is left untouched.  Rings the reference's
    cooley_tukey_ntt / gentleman_sande_intt would reject (root_order other than twice the degree, a degree that is no power of
    two, a composite modulus) are rejected with the same exceptions (algebra/ntt.py:239-270)."""
    if isinstance(x, PolynomialCoefficientRepresentation):
        if x._secpar() is None:
            _ntt._validate(x._peek(), x.modulus, x.root_order, _ntt.power_table(x.modulus, x.degree, x.root), "root_powers")
        return PolynomialNTTRepresentation._from_tensor(x._ring(), x._ctx().ntt_fwd(x._tensor()))
    elif isinstance(x, PolynomialNTTRepresentation):
        if x._secpar() is None:
            _ntt._validate(x._peek(), x.modulus, x.root_order, _ntt.power_table(x.modulus, x.degree, x.inv_root), "inv_root_powers")
        return PolynomialCoefficientRepresentation._from_tensor(x._ring(), x._ctx().ntt_inv(x._tensor()))
    else:
        raise NotImplementedError(f"Transform for {type(x)} not implemented")


def _effective_seed(seed) -> int:
    """The integer CPython's ``random.seed(seed)`` turns its argument into (Lib/random.py seed(), version 2, and Modules/_randommodule.c
    random_seed): an int is used as is (its absolute value), str / bytes / bytearray become int.from_bytes(a + sha512(a).digest()), anything
    else its hash.  The device sampler takes up to 256 bits; longer keys (every str / bytes seed) are refused there."""
    if isinstance(seed, int):
        return seed
    if isinstance(seed, str):
        seed = seed.encode()
    if isinstance(seed, (bytes, bytearray)):
        import hashlib

        return int.from_bytes(bytes(seed) + hashlib.sha512(seed).digest(), "big")
    return abs(hash(seed))


def _host_draw_coefficients(degree: int, bound: int, weight: int) -> List[int]:
    """The reference's draw order on CPython's global generator (:451-459)."""
    coefficients = [(1 + _random.randrange(bound)) * (1 - 2 * _random.randrange(2)) for _ in range(weight)]
    coefficients += [0 for _ in range(degree - len(coefficients))]
    if weight < degree:
        for i in range(degree - 1, 0, -1):
            j = _random.randrange(i + 1)
            coefficients[i], coefficients[j] = coefficients[j], coefficients[i]
    return coefficients


def sample_polynomial_coefficient_representation(modulus: int, degree: int, root: int, inv_root: int, root_order: int, norm_bound: int,
                                                 weight_bound: int, seed: Optional[int]) -> PolynomialCoefficientRepresentation:
    """Reference :436-467.  With a seed the polynomial comes from the device MT19937 sampler and CPython's global
    generator is left exactly where the reference leaves it; without a seed the values ARE the
    next draws of the caller's global generator, which lives in the interpreter."""
    num = max(0, min(degree, weight_bound))
    bound = max(0, min(modulus // 2, norm_bound))
    if seed is None:
        coefficients = _host_draw_coefficients(degree, bound, num)
        return PolynomialCoefficientRepresentation(modulus=modulus, degree=degree, root=root, inv_root=inv_root, root_order=root_order,
                                                   coefficients=coefficients)
    proto = PolynomialCoefficientRepresentation(modulus=modulus, degree=degree, root=root, inv_root=inv_root, root_order=root_order,
                                                coefficients=[0] * degree)
    _random.seed(seed)
    _host_draw_coefficients(degree, bound, num)  # state side effect only (and the reference's ValueError when bound = 0 < num: randrange(0))
    t = proto._ctx().sample_coef([_effective_seed(seed)], norm_bound=max(bound, 1), weight_bound=num)[0]
    return PolynomialCoefficientRepresentation._from_tensor(proto._ring(), t)


def sample_polynomial_ntt_representation(modulus: int, degree: int, root: int, inv_root: int, root_order: int,
                                         seed: Optional[int]) -> PolynomialNTTRepresentation:
    """Reference :470-488."""
    if seed is None:
        values = [_random.randrange(modulus) - (modulus // 2) for _ in range(degree)]
        return PolynomialNTTRepresentation(modulus=modulus, degree=degree, root=root, inv_root=inv_root, root_order=root_order, values=values)
    proto = PolynomialNTTRepresentation(modulus=modulus, degree=degree, root=root, inv_root=inv_root, root_order=root_order, values=[0] * degree)
    t = proto._ctx().sample_ntt([_effective_seed(seed)])[0]
    _random.seed(seed)
    for _ in range(degree):
        _random.randrange(modulus)  # state side effect only
    return PolynomialNTTRepresentation._from_tensor(proto._ring(), t)
