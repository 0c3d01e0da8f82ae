"""ctypes binding of libfusion_b200.so (C ABI: include/fusion_b200.h).

torch is used only as the owner of device memory and streams: every tensor handed to this
module is a contiguous int32 CUDA tensor whose raw pointer goes through the C ABI.
No CPU fallback exists: a missing library or a missing GPU raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfusion_b200.so")
_lib = None

Q = 2147465729


class FusionB200Error(RuntimeError):
    pass


def load_library():
    """Loads the CUDA library; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FusionB200Error(
            f"{LIB_PATH} not found: build it with `python -m fusion_cryptography_b200.build` "
            "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.fz_last_error.restype = C.c_char_p
    lib.fz_param.restype = C.c_longlong
    lib.fz_param.argtypes = [C.c_void_p, C.c_char_p]
    lib.fz_init.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    lib.fz_destroy.argtypes = [C.c_void_p]
    _lib = lib
    return lib


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())


class Context:
    """One per (process, device, secpar); wraps fz_init / fz_destroy."""

    def __init__(self, secpar: int, device: int = 0):
        if not torch.cuda.is_available():
            raise FusionB200Error("fusion_cryptography_b200 needs a CUDA device (no CPU fallback)")
        self.lib = load_library()
        self.secpar = secpar
        self.device_index = device
        self.device = torch.device("cuda", device)
        torch.cuda.init()
        with torch.cuda.device(self.device):
            torch.empty(1, device=self.device)  # make sure torch's primary context exists first
            h = C.c_void_p()
            rc = self.lib.fz_init(C.c_int(device), C.c_int(secpar), C.byref(h))
            if rc:
                raise FusionB200Error(self.lib.fz_last_error().decode())
        self._h = h
        self.d = self.param("d")
        self.ell = self.param("ell")
        self.capacity = self.param("capacity")

    def close(self):
        if getattr(self, "_h", None):
            self.lib.fz_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def param(self, name: str) -> int:
        v = int(self.lib.fz_param(self._h, name.encode()))
        if v == -1:
            raise KeyError(name)
        return v

    # ------------------------------------------------------------------ helpers
    def _stream(self) -> C.c_void_p:
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _chk(self, rc: int):
        if rc:
            raise FusionB200Error(self.lib.fz_last_error().decode())

    def _dev(self, x, name="tensor") -> torch.Tensor:
        if not isinstance(x, torch.Tensor):
            x = torch.as_tensor(x, dtype=torch.int32)
        if x.dtype != torch.int32:
            raise TypeError(f"{name} must be int32")
        if x.device != self.device:
            x = x.to(self.device, non_blocking=True)
        return x.contiguous()

    def _polys(self, x: torch.Tensor, name: str) -> int:
        if x.numel() % self.d:
            raise ValueError(f"{name}: number of coefficients must be a multiple of degree {self.d}")
        return x.numel() // self.d

    # ------------------------------------------------------------------ K1-K3
    def ntt_fwd(self, x, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        x = self._dev(x, "x")
        n = self._polys(x, "x")
        out = torch.empty_like(x) if out is None else out
        self._chk(self.lib.fz_ntt_fwd(self._h, _ptr(x), _ptr(out), C.c_size_t(n), self._stream()))
        return out

    def ntt_inv(self, x, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        x = self._dev(x, "x")
        n = self._polys(x, "x")
        out = torch.empty_like(x) if out is None else out
        self._chk(self.lib.fz_ntt_inv(self._h, _ptr(x), _ptr(out), C.c_size_t(n), self._stream()))
        return out

    def pointwise_mul(self, a, b, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        a, b = self._dev(a, "a"), self._dev(b, "b")
        if a.shape != b.shape:
            raise ValueError("pointwise_mul: shape mismatch")
        n = self._polys(a, "a")
        out = torch.empty_like(a) if out is None else out
        self._chk(self.lib.fz_pointwise_mul(self._h, _ptr(a), _ptr(b), _ptr(out), C.c_size_t(n), self._stream()))
        return out

    def pointwise_muladd(self, a, b, c, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        a, b, c = self._dev(a, "a"), self._dev(b, "b"), self._dev(c, "c")
        if a.shape != b.shape or a.shape != c.shape:
            raise ValueError("pointwise_muladd: shape mismatch")
        n = self._polys(a, "a")
        out = torch.empty_like(a) if out is None else out
        self._chk(self.lib.fz_pointwise_muladd(self._h, _ptr(a), _ptr(b), _ptr(c), _ptr(out), C.c_size_t(n), self._stream()))
        return out

    def polymul(self, a, b, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        a, b = self._dev(a, "a"), self._dev(b, "b")
        if a.shape != b.shape:
            raise ValueError("polymul: shape mismatch")
        n = self._polys(a, "a")
        out = torch.empty_like(a) if out is None else out
        self._chk(self.lib.fz_polymul(self._h, _ptr(a), _ptr(b), _ptr(out), C.c_size_t(n), self._stream()))
        return out

    # ------------------------------------------------------------------ K4-K5
    def matvec(self, A, s, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """out[b] = A (1 x ell) . s[b] (ell x 1);  A: [ell, d], s: [batch, ell, d] -> [batch, d]."""
        A, s = self._dev(A, "A"), self._dev(s, "s")
        if A.numel() != self.ell * self.d or s.numel() % (self.ell * self.d):
            raise ValueError("matvec: Matrix dimension mismatch")
        batch = s.numel() // (self.ell * self.d)
        out = torch.empty((batch, self.d), dtype=torch.int32, device=self.device) if out is None else out
        self._chk(self.lib.fz_matvec(self._h, _ptr(A), _ptr(s), _ptr(out), C.c_size_t(batch), self._stream()))
        return out

    def sign_mac(self, sk_l, sk_r, c_hat, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        sk_l, sk_r, c_hat = self._dev(sk_l, "sk_l"), self._dev(sk_r, "sk_r"), self._dev(c_hat, "c_hat")
        batch = c_hat.numel() // self.d
        if sk_l.numel() != batch * self.ell * self.d or sk_r.numel() != sk_l.numel():
            raise ValueError("sign_mac: shape mismatch")
        out = torch.empty_like(sk_l) if out is None else out
        self._chk(self.lib.fz_sign_mac(self._h, _ptr(sk_l), _ptr(sk_r), _ptr(c_hat), _ptr(out), C.c_size_t(batch), self._stream()))
        return out

    # ------------------------------------------------------------------ host-buffer path (end-to-end)
    def ntt_host(self, host_in: torch.Tensor, host_out: torch.Tensor, inverse: bool = False) -> torch.Tensor:
        if host_in.device.type != "cpu" or host_out.device.type != "cpu" or host_in.dtype != torch.int32:
            raise TypeError("ntt_host takes int32 CPU tensors")
        n = self._polys(host_in, "host_in")
        self._chk(self.lib.fz_ntt_host(self._h, C.c_int(1 if inverse else 0), _ptr(host_in), _ptr(host_out), C.c_size_t(n)))
        return host_out

    def pass_host(self, host_a: torch.Tensor, dev_bhat: torch.Tensor, host_out: torch.Tensor) -> torch.Tensor:
        """host a -> NTT -> * b_hat (device resident) -> INTT -> host out, copies included."""
        if host_a.device.type != "cpu" or host_out.device.type != "cpu" or host_a.dtype != torch.int32:
            raise TypeError("pass_host takes int32 CPU tensors for a and out")
        dev_bhat = self._dev(dev_bhat, "b_hat")
        n = self._polys(host_a, "host_a")
        if dev_bhat.numel() != host_a.numel() or host_out.numel() != host_a.numel():
            raise ValueError("pass_host: shape mismatch")
        self._chk(self.lib.fz_ntt_pass_host(self._h, _ptr(host_a), _ptr(dev_bhat), _ptr(host_out), C.c_size_t(n)))
        return host_out


_contexts = {}


def get_context(secpar: int, device: Optional[int] = None) -> Context:
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    key = (secpar, device)
    if key not in _contexts:
        _contexts[key] = Context(secpar, device)
    return _contexts[key]
