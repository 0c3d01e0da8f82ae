"""Compact wire format for keys and signatures: the int32 little-endian coefficients, row-major ``[rows][d]``, preceded by a
12-byte header -- nothing the reference defines (its objects only have ``str()``), but what a batched caller ships between machines instead of
the 6.4 KB (verification key) / 1 MB (signature) of decimal text.  Decoding gives the drop-in's ordinary objects: their data stays on the device
and ``str()`` / ``.values`` are rendered on demand, byte for byte what the reference prints.

    blob = wire.signature_to_bytes(sig);   sig2 = wire.signature_from_bytes(params, blob);   str(sig2) == str(sig)
"""
from __future__ import annotations

import struct

import numpy as np
import torch

from .algebra.matrices import GeneralMatrix
from .algebra.polynomials import PolynomialNTTRepresentation
from .fusion import fusion as _ff

_MAGIC = b"FZ1"
_KINDS = {"vk": 1, "sig": 2}


def _pack(kind: str, secpar: int, t: torch.Tensor) -> bytes:
    a = t.detach().to("cpu", torch.int32).contiguous().numpy().astype("<i4", copy=False)
    return _MAGIC + struct.pack("<BHIH", _KINDS[kind], secpar, a.shape[0], a.shape[1]) + a.tobytes()


def _unpack(kind: str, params, blob: bytes) -> torch.Tensor:
    if len(blob) < 12 or blob[:3] != _MAGIC:
        raise ValueError("not a fusion_b200 wire blob")
    k, secpar, rows, d = struct.unpack("<BHIH", blob[3:12])
    if k != _KINDS[kind] or secpar != params.secpar or d != params.degree:
        raise ValueError(f"wire blob is a {k}/{secpar}/{d} object, expected {_KINDS[kind]}/{params.secpar}/{params.degree}")
    if len(blob) != 12 + 4 * rows * d:
        raise ValueError("wire blob has the wrong length")
    a = np.frombuffer(blob, dtype="<i4", offset=12).reshape(rows, d)
    h = params.modulus // 2
    if a.size and (int(a.min()) < -h or int(a.max()) > h):
        raise ValueError("wire blob holds values outside the canonical centred range")
    return torch.from_numpy(a.astype(np.int32)).to(params._ctx().device)


def vk_to_bytes(params, vk) -> bytes:
    """OneTimeVerificationKey -> 12 + 2 * d * 4 bytes (left polynomial, right polynomial)"""
    d = params.degree
    if not _ff._device_hashable(params, [vk]):   # the format holds canonical centred residues; str() of such a key would not survive the round trip
        raise ValueError("verification key holds values outside the canonical centred range; the wire format cannot represent its str()")
    return _pack("vk", params.secpar, torch.stack([vk.left_vk_hat._tensor().reshape(d), vk.right_vk_hat._tensor().reshape(d)]))


def vk_from_bytes(params, blob: bytes):
    t = _unpack("vk", params, blob)
    if t.shape[0] != 2:
        raise ValueError("a verification key has two polynomials")
    ring, d = params._ring(), params.degree
    return _ff.OneTimeVerificationKey(left_vk_hat=GeneralMatrix._from_tensor(ring, PolynomialNTTRepresentation, t[0].reshape(1, 1, d)),
                                      right_vk_hat=GeneralMatrix._from_tensor(ring, PolynomialNTTRepresentation, t[1].reshape(1, 1, d)))


def signature_to_bytes(params, sig) -> bytes:
    """Signature (also an aggregate) -> 12 + ell * d * 4 bytes"""
    return _pack("sig", params.secpar, sig.signature_hat._tensor().reshape(params.num_rows_sk, params.degree))


def signature_from_bytes(params, blob: bytes):
    t = _unpack("sig", params, blob)
    if t.shape[0] != params.num_rows_sk:
        raise ValueError(f"a signature has {params.num_rows_sk} polynomials")
    return _ff.Signature(signature_hat=GeneralMatrix._from_tensor(params._ring(), PolynomialNTTRepresentation, t.reshape(params.num_rows_sk, 1, params.degree)))
