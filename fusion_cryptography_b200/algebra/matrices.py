"""Host mirror of the reference's ``algebra.matrices.GeneralMatrix`` (algebra/matrices.py): same
validation, ``str()`` format (which the scheme hashes), operators and error classes.  Products and sums
of matrices of scheme polynomials run as single kernels (fz_matvec / fz_sign_mac / pointwise); other
element classes fall back to the reference's element-wise semantics through the elements' own operators.
"""
from __future__ import annotations

from copy import deepcopy  # noqa: F401  (the reference's module exposes it, and its tests reach it through `from algebra.matrices import *`)
from typing import List, Union

try:
    from .polynomials import PolynomialCoefficientRepresentation, PolynomialNTTRepresentation, _eq_zero
except ImportError:  # drop-in layout
    from algebra.polynomials import PolynomialCoefficientRepresentation, PolynomialNTTRepresentation, _eq_zero


_ALGEBRA_DUNDERS = ("__eq__", "__add__", "__neg__", "__sub__", "__mul__")


def is_algebraic_class(cls):
    return all(hasattr(cls, m) for m in _ALGEBRA_DUNDERS)


def _validated_elem_class(matrix):
    """The reference's constructor checks (algebra/matrices.py:15-36), in the reference's order and with its messages."""
    problems = (
        (lambda: not isinstance(matrix, list), "Matrix must be a list"),
        (lambda: not matrix, "Matrix must not be empty."),
        (lambda: any(not isinstance(row, list) for row in matrix), "Matrix must be a list of lists"),
        (lambda: any(not row for row in matrix), "Matrix must not contain empty lists"),
        (lambda: any(len(row) != len(matrix[0]) for row in matrix), "All rows must have the same length"),
        (lambda: not is_algebraic_class(cls=type(matrix[0][0])), "Matrix must contain only instances of the same algebraic class"),
        (lambda: any(not isinstance(item, type(matrix[0][0])) for row in matrix for item in row),
         "Matrix must contain only instances of the same algebraic class"),
    )
    for failed, message in problems:
        if failed():
            raise ValueError(message)
    return type(matrix[0][0])


class GeneralMatrix:
    elem_class: type
    matrix: List[List[object]]

    def __init__(self, matrix: List[list]):
        self.elem_class = _validated_elem_class(matrix)
        self.matrix = matrix
        self._base = None  # device tensor [rows, cols, d] the elements are views of (set by the kernels' wrappers)

    @property
    def _shape(self):
        return len(self.matrix), len(self.matrix[0])

    def _map(self, fn):
        return GeneralMatrix(matrix=[[fn(item) for item in row] for row in self.matrix])

    def _is_zero_operand(self, other) -> bool:
        """`other == 0` in the reference's sense: an all-zero matrix or a zero scalar / polynomial (matrices.py:62-63, 75-76)."""
        return other._all_zero() if isinstance(other, GeneralMatrix) else _eq_zero(other)

    # ------------------------------------------------------------------ device plumbing
    @classmethod
    def _from_tensor(cls, ring, elem_class, t):
        """t: int32 device tensor [rows, cols, d]; elements become views of it."""
        rows, cols = t.shape[0], t.shape[1]
        m = cls.__new__(cls)
        m.elem_class = elem_class
        m.matrix = [[elem_class._from_tensor(ring, t[i, j]) for j in range(cols)] for i in range(rows)]
        m._base = t
        return m

    def _is_scheme_ntt(self) -> bool:
        return self.elem_class is PolynomialNTTRepresentation

    def _tensor(self):
        """int32 device tensor [rows, cols, d]; reuses the kernel output the elements still point into."""
        import torch

        base = getattr(self, "_base", None)
        if base is not None and len(self.matrix) == base.shape[0] and all(
                len(row) == base.shape[1] and all(getattr(e, "_vals", 0) is None and e._t is not None for e in row) for row in self.matrix):
            # the elements must still be the views handed out with the base (a caller may have replaced an entry): compare addresses by
            # arithmetic -- indexing base[i, j] would build a fresh view per entry, which dominated small aggregate / verify calls
            p0, s0, s1 = base.data_ptr(), base.stride(0) * base.element_size(), base.stride(1) * base.element_size()
            if all(e._t.data_ptr() == p0 + i * s0 + j * s1 for i, row in enumerate(self.matrix) for j, e in enumerate(row)):
                return base
        return torch.stack([torch.stack([e._tensor() for e in row]) for row in self.matrix])

    def _all_zero(self) -> bool:
        return all(all(_eq_zero(item) for item in row) for row in self.matrix)

    def _ring(self):
        return self.matrix[0][0]._ring()

    def _same_ring(self) -> bool:
        r = self._ring()
        return all(e._ring() == r for row in self.matrix for e in row)

    # ------------------------------------------------------------------ reference API
    def __str__(self):
        return f"GeneralMatrix(elem_class={self.elem_class}, matrix={self.matrix})"

    def __repr__(self):
        return self.__str__()

    def __len__(self):
        return len(self.matrix)

    def __iter__(self):
        return iter(self.matrix)

    def __getitem__(self, item):
        return self.matrix[item]

    def __setitem__(self, key, value):
        self.matrix[key] = value

    def __delitem__(self, key):
        self.matrix[key] = 0

    def __eq__(self, other):
        if self._is_zero_operand(other):
            return self._all_zero()
        if not isinstance(other, GeneralMatrix) or self.elem_class != other.elem_class or self._shape != other._shape:
            return False
        return self.matrix == other.matrix

    __hash__ = None

    def __add__(self, other):
        if self._is_zero_operand(other):
            return self
        if not isinstance(other, GeneralMatrix) or self.elem_class != other.elem_class:
            raise NotImplementedError("Can only add GeneralMatrix objects of the same algebraic class")
        if self._shape != other._shape:
            raise ValueError("Matrix dimensions must match")
        on_device = self.elem_class in (PolynomialNTTRepresentation, PolynomialCoefficientRepresentation)
        if on_device and self._same_ring() and other._same_ring() and self._ring() == other._ring():
            ctx = self.matrix[0][0]._ctx()
            return GeneralMatrix._from_tensor(self._ring(), self.elem_class, ctx.pointwise_add(self._tensor(), other._tensor()))
        return GeneralMatrix(matrix=[[x + y for x, y in zip(ra, rb)] for ra, rb in zip(self.matrix, other.matrix)])

    def __radd__(self, other):
        return self if other == 0 else self + other

    def __neg__(self):
        return self._map(lambda item: -item)

    def __sub__(self, other):
        return self + (-other)

    def __mul__(self, other):
        if isinstance(other, self.elem_class):
            # matrix * element (reference :109-114): every entry times one polynomial
            if self._is_scheme_ntt() and self._same_ring() and other._ring() == self._ring() and not _eq_zero(other):
                ctx = other._ctx()
                t = self._tensor()
                out = ctx.pointwise_mul(t, other._tensor().expand_as(t).contiguous())
                return GeneralMatrix._from_tensor(self._ring(), self.elem_class, out)
            return self._map(lambda item: item * other)
        if not isinstance(other, GeneralMatrix) or self.elem_class != other.elem_class:
            raise TypeError("Can only multiply matrices of the same algebraic class")
        (rows, inner), (inner_b, cols) = self._shape, other._shape
        if inner != inner_b:
            raise ValueError("Matrix dimension mismatch")
        if self._is_scheme_ntt() and self._same_ring() and other._same_ring() and self._ring() == other._ring():
            ctx = self.matrix[0][0]._ctx()
            if rows == 1 and cols == 1 and inner == ctx.ell:
                # (1 x ell) . (ell x 1): the scheme's A.s / A.z (reference :121-131) as one fused MAC kernel
                out = ctx.matvec(self._tensor().reshape(inner, ctx.d), other._tensor().reshape(1, inner, ctx.d))
                return GeneralMatrix._from_tensor(self._ring(), self.elem_class, out.reshape(1, 1, ctx.d))
        # general shapes: left-to-right accumulation through the elements' own operators (every + and * re-centres, so the
        # value equals the reference's for any order; the order is kept anyway)
        def cell(i, j):
            acc = self.matrix[i][0] * other.matrix[0][j]
            for k in range(1, inner):
                acc += self.matrix[i][k] * other.matrix[k][j]
            return acc

        result = GeneralMatrix.__new__(GeneralMatrix)
        result.elem_class, result._base = self.elem_class, None
        result.matrix = [[cell(i, j) for j in range(cols)] for i in range(rows)]
        return result

    def __mod__(self, other):
        if not isinstance(other, int):
            raise TypeError("Can only take the remainder of a matrix with an integer")
        if other <= 1:
            raise ValueError("Modulus must be greater than 1")
        return self._map(lambda item: item % other)

    def _entries(self):
        return [item for row in self.matrix for item in row]

    def norm(self, p: Union[int, str]):
        entries = self._entries()
        if any(not hasattr(item, "norm") for item in entries):
            raise NotImplementedError("Matrix elements must have a norm method")
        if p == "infty":
            return max(item.norm(p=p) for item in entries)

    def weight(self):
        entries = self._entries()
        if any(not hasattr(item, "weight") for item in entries):
            raise NotImplementedError("Matrix elements must have a weight method")
        return max(item.weight() for item in entries)


GeneralMatrix.__module__ = "algebra.matrices"
