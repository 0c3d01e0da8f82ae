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
LIB_PATH = os.environ.get("FUSION_B200_LIB", os.path.join(_HERE, "libfusion_b200.so"))
_lib = None

Q = 2147465729


class FusionB200Error(RuntimeError):
    pass


_CTYPES = {"int": C.c_int, "size_t": C.c_size_t, "uint32_t": C.c_uint32, "uint64_t": C.c_uint64, "int32_t": C.c_int32, "long long": C.c_longlong}
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "fusion_b200.h")


def parse_prototypes(header_text: str):
    """{name: (restype, [argtypes])} of every fz_* prototype in include/fusion_b200.h: pointers (and handles) become c_void_p, `const char*` c_char_p,
    scalars their exact C width -- so that a Python int that does not fit its C parameter raises instead of being truncated silently."""
    import re

    text = re.sub(r"/\*.*?\*/", " ", header_text, flags=re.S)
    out = {}
    for m in re.finditer(r"(?:^|\n)\s*([A-Za-z_][\w \*]*?)\b(fz_\w+)\s*\(([^;{}]*?)\)\s*;", text):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()

        def ctype(decl: str):
            decl = decl.strip()
            if decl in ("void", ""):
                return None
            if "*" in decl:
                base = decl.replace("const", "").split("*")[0].strip()
                return C.c_char_p if base == "char" else C.c_void_p
            t = decl.replace("const", "").strip()
            t = " ".join(t.split()[:-1]) if len(t.split()) > 1 and t.split()[-1] not in ("long", "int") else t   # drop the parameter name
            return _CTYPES[t]

        argtypes = [ctype(a) for a in args.split(",")] if args and args != "void" else []
        out[name] = (C.c_char_p if ("char" in ret and "*" in ret) else (C.c_void_p if "*" in ret else _CTYPES[ret]), [a for a in argtypes if a is not None])
    return out


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
    # argument and result types of every entry point, read from the header the library was built against
    if os.path.exists(HEADER_PATH):
        for name, (restype, argtypes) in parse_prototypes(open(HEADER_PATH).read()).items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = restype, argtypes
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

    def set_option(self, name: str, value: int) -> None:
        self.lib.fz_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_longlong]
        self._chk(self.lib.fz_set_option(self._h, name.encode(), value))

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

    def pointwise_add(self, a, b, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        a, b = self._dev(a, "a"), self._dev(b, "b")
        if a.shape != b.shape:
            raise ValueError("pointwise_add: shape mismatch")
        n = self._polys(a, "a")
        out = torch.empty_like(a) if out is None else out
        self._chk(self.lib.fz_pointwise_add(self._h, _ptr(a), _ptr(b), _ptr(out), C.c_size_t(n), self._stream()))
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
    def ntt_pass(self, a, b_hat, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """INTT(NTT(a) o b_hat) fused in one kernel (3 KiB of HBM traffic per polynomial instead of 7 KiB)."""
        a, b_hat = self._dev(a, "a"), self._dev(b_hat, "b_hat")
        if a.shape != b_hat.shape:
            raise ValueError("ntt_pass: shape mismatch")
        n = self._polys(a, "a")
        out = torch.empty_like(a) if out is None else out
        self._chk(self.lib.fz_ntt_pass(self._h, _ptr(a), _ptr(b_hat), _ptr(out), C.c_size_t(n), self._stream()))
        return out

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


class Ring:
    """Any ring Z_q[X]/(X^d + 1) of the reference's algebra layer (q an odd prime below 2^31, d = 2^k <= 16384) with caller-supplied
    bit-reversed power tables; wraps fz_ring_create / fz_ring_destroy.  Same method names as Context for the operations it has."""

    MAX_LOGD = 14

    def __init__(self, modulus: int, degree: int, fwd_table=None, inv_table=None, device: int = 0):
        if not torch.cuda.is_available():
            raise FusionB200Error("fusion_cryptography_b200 needs a CUDA device (no CPU fallback)")
        logd = degree.bit_length() - 1
        if degree < 2 or (1 << logd) != degree or logd > self.MAX_LOGD or modulus < 3 or modulus % 2 == 0 or modulus >= 1 << 31:
            raise NotImplementedError(
                f"libfusion_b200 transforms rings with an odd modulus below 2^31 and a power-of-two degree up to {1 << self.MAX_LOGD} "
                f"(got modulus={modulus}, degree={degree}); there is no CPU fallback")
        self.lib = load_library()
        self.q, self.d, self.logd = modulus, degree, logd
        self.device_index = device
        self.device = torch.device("cuda", device)
        tabs = []
        for t in (fwd_table, inv_table):
            if t is None:
                tabs.append(None)
                continue
            if len(t) < degree:
                raise IndexError("list index out of range")   # what the reference's table[m + i] raises (algebra/ntt.py:280, :360)
            tabs.append((C.c_uint32 * degree)(*[int(v) % modulus for v in t[:degree]]))
        self.lib.fz_ring_create.argtypes = [C.c_int, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
        self.lib.fz_ring_destroy.argtypes = [C.c_void_p]
        with torch.cuda.device(self.device):
            torch.empty(1, device=self.device)
            h = C.c_void_p()
            rc = self.lib.fz_ring_create(C.c_int(device), C.c_uint32(modulus), C.c_int(logd), tabs[0], tabs[1], C.byref(h))
            if rc:
                raise FusionB200Error(self.lib.fz_last_error().decode())
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self.lib.fz_ring_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    _stream = Context._stream
    _chk = Context._chk
    _dev = Context._dev
    _polys = Context._polys

    def _unary(self, fn, x, out):
        x = self._dev(x, "x")
        n = self._polys(x, "x")
        out = torch.empty_like(x) if out is None else out
        self._chk(fn(self._h, _ptr(x), _ptr(out), C.c_size_t(n), self._stream()))
        return out

    def _binary(self, fn, a, b, out, what):
        a, b = self._dev(a, "a"), self._dev(b, "b")
        if a.shape != b.shape:
            raise ValueError(f"{what}: shape mismatch")
        n = self._polys(a, "a")
        out = torch.empty_like(a) if out is None else out
        self._chk(fn(self._h, _ptr(a), _ptr(b), _ptr(out), C.c_size_t(n), self._stream()))
        return out

    def ntt_fwd(self, x, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        return self._unary(self.lib.fz_ring_ntt_fwd, x, out)

    def ntt_inv(self, x, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        return self._unary(self.lib.fz_ring_ntt_inv, x, out)

    def pointwise_mul(self, a, b, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        return self._binary(self.lib.fz_ring_pointwise_mul, a, b, out, "pointwise_mul")

    def pointwise_add(self, a, b, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        return self._binary(self.lib.fz_ring_pointwise_add, a, b, out, "pointwise_add")

    def polymul(self, a, b, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """INTT(NTT(a) o NTT(b)), algebra/ntt.py:450-471 (result only)."""
        fa = self.ntt_fwd(a)
        fb = self.ntt_fwd(b)
        self.pointwise_mul(fa, fb, out=fa)
        return self.ntt_inv(fa, out=out)

    def sample_coef(self, seeds, norm_bound: int = 52, weight_bound=None) -> torch.Tensor:
        """algebra/polynomials.py:436-467 with integer seeds (device MT19937, CPython's seeding and randrange) -> [n, d]."""
        keys, klen = Context._seed_keys(self, seeds)
        out = torch.empty((len(seeds), self.d), dtype=torch.int32, device=self.device)
        wb = self.d if weight_bound is None else weight_bound
        self._chk(self.lib.fz_ring_sample_coef(self._h, _ptr(keys), C.c_int(8), _ptr(klen), C.c_size_t(len(seeds)), C.c_int(norm_bound), C.c_int(wb),
                                               _ptr(out), self._stream()))
        return out

    def decode_poly(self, data: torch.Tensor, blen: int, norm_bound: int, weight_bound: int, bytes_per_coefficient: int, bytes_per_index: int) -> torch.Tensor:
        """fusion/fusion.py:422-481 on [n, stride] uint8 rows (the first blen bytes of each) -> [n, d]."""
        if data.dtype != torch.uint8:
            raise TypeError("bytes must be uint8")
        data = data.to(self.device).contiguous()
        n, stride = data.shape
        out = torch.empty((n, self.d), dtype=torch.int32, device=self.device)
        rc = self.lib.fz_ring_decode_poly(self._h, _ptr(data), C.c_size_t(stride), C.c_size_t(blen), C.c_size_t(n), C.c_longlong(min(norm_bound, 1 << 62)),
                                          C.c_int(weight_bound), C.c_int(bytes_per_coefficient), C.c_int(bytes_per_index), _ptr(out), self._stream())
        if rc == -3:
            raise ValueError(self.lib.fz_last_error().decode())
        self._chk(rc)
        return out

    def sample_ntt(self, seeds) -> torch.Tensor:
        """algebra/polynomials.py:470-488 with integer seeds -> [n, d]."""
        keys, klen = Context._seed_keys(self, seeds)
        out = torch.empty((len(seeds), self.d), dtype=torch.int32, device=self.device)
        self._chk(self.lib.fz_ring_sample_ntt(self._h, _ptr(keys), C.c_int(8), _ptr(klen), C.c_size_t(len(seeds)), _ptr(out), self._stream()))
        return out


_rings = {}
_RING_CACHE = 64


def get_ring(modulus: int, degree: int, fwd_table=None, inv_table=None, device: Optional[int] = None) -> Ring:
    """Rings are cached by (modulus, degree, tables): the reference's callers pass the same few tables over and over."""
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    key = (modulus, degree, None if fwd_table is None else tuple(int(v) % modulus for v in fwd_table[:degree]),
           None if inv_table is None else tuple(int(v) % modulus for v in inv_table[:degree]), device)
    r = _rings.get(key)
    if r is None:
        if len(_rings) >= _RING_CACHE:
            _rings.pop(next(iter(_rings)))   # dropped, not closed: a caller may still hold it; fz_ring_destroy runs when the last reference goes
        r = _rings[key] = Ring(modulus, degree, fwd_table, inv_table, device)
    return r


def get_context(secpar: int, device: Optional[int] = None) -> Context:
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    key = (secpar, device)
    if key not in _contexts:
        _contexts[key] = Context(secpar, device)
    return _contexts[key]


# ---------------------------------------------------------------------------------------------
# Hashing / samplers / scheme stages (K7-K11).  Methods are attached to Context below.
def _u8(self, x, name="bytes") -> torch.Tensor:
    if not isinstance(x, torch.Tensor):
        x = torch.frombuffer(bytearray(x), dtype=torch.uint8) if len(x) else torch.empty(0, dtype=torch.uint8)
    if x.dtype != torch.uint8:
        raise TypeError(f"{name} must be uint8")
    return x.to(self.device).contiguous()


def _pack_messages(self, msgs):
    """list of bytes -> (device uint8 buffer, device int64 offsets[n+1])"""
    offs = [0]
    for m in msgs:
        offs.append(offs[-1] + len(m))
    data = b"".join(msgs) + b"\0"
    buf = torch.frombuffer(bytearray(data), dtype=torch.uint8).to(self.device)
    return buf, torch.tensor(offs, dtype=torch.int64, device=self.device)


def _ws(self, nbytes: int) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=self.device)


def sha3_256(self, msgs):
    buf, offs = self._pack_messages(msgs)
    out = torch.empty((len(msgs), 32), dtype=torch.uint8, device=self.device)
    self._chk(self.lib.fz_sha3_256(self._h, _ptr(buf), _ptr(offs), C.c_size_t(len(msgs)), _ptr(out), self._stream()))
    return out


def shake256(self, msgs, out_len: int):
    buf, offs = self._pack_messages(msgs)
    out = torch.empty((len(msgs), max(out_len, 1)), dtype=torch.uint8, device=self.device)
    self._chk(self.lib.fz_shake256(self._h, _ptr(buf), _ptr(offs), C.c_size_t(len(msgs)), _ptr(out), C.c_size_t(out_len), self._stream()))
    return out[:, :out_len]


class PackedMessages:
    """messages resident on the device: one uint8 buffer + int64 offsets[n + 1]; slicing keeps absolute offsets"""

    def __init__(self, buf: torch.Tensor, offs: torch.Tensor):
        self.buf, self.offs = buf, offs

    def __len__(self):
        return self.offs.numel() - 1

    def __getitem__(self, sl):
        if not isinstance(sl, slice):
            raise TypeError("PackedMessages supports slices only")
        lo, hi, step = sl.indices(len(self))
        if step != 1:
            raise ValueError("PackedMessages slices must be contiguous")
        return PackedMessages(self.buf, self.offs[lo:hi + 1].contiguous())


def pack_messages(self, msgs) -> "PackedMessages":
    buf, offs = self._pack_messages(msgs)
    return PackedMessages(buf, offs)


def hash_msg(self, msgs) -> torch.Tensor:
    """fusion.py:405-409 on a list of message byte strings (or PackedMessages) -> [n, 32] digests (int = little endian)."""
    if isinstance(msgs, PackedMessages):
        buf, offs, n = msgs.buf, msgs.offs, len(msgs)
    else:
        buf, offs = self._pack_messages(msgs)
        n = len(msgs)
    out = torch.empty((n, 32), dtype=torch.uint8, device=self.device)
    self._chk(self.lib.fz_hash_msg(self._h, _ptr(buf), _ptr(offs), C.c_size_t(n), _ptr(out), self._stream()))
    return out


def format_text(self, mode: int, vk_l, vk_r, digests, c_hat=None):
    vk_l, vk_r = self._dev(vk_l), self._dev(vk_r)
    n = vk_l.numel() // self.d
    digests = self._u8(digests)
    c_hat = None if c_hat is None else self._dev(c_hat)
    self.lib.fz_text_stride.restype = C.c_size_t
    stride = int(self.lib.fz_text_stride(self._h, C.c_int(mode)))
    out = torch.zeros((n, stride), dtype=torch.uint8, device=self.device)
    lens = torch.empty(n, dtype=torch.int32, device=self.device)
    self._chk(self.lib.fz_format_text(self._h, C.c_int(mode), _ptr(vk_l), _ptr(vk_r), _ptr(digests), _ptr(c_hat), C.c_size_t(n), _ptr(out), _ptr(lens),
                                      self._stream()))
    return out, lens


def decode_poly(self, data: torch.Tensor, blen: int, log2_bias: int, norm_bound: int, weight_bound: int) -> torch.Tensor:
    """data: [n, stride] uint8 device tensor; decodes the first blen bytes of each row (fusion.py:422-481)."""
    data = self._u8(data)
    n, stride = data.shape
    out = torch.empty((n, self.d), dtype=torch.int32, device=self.device)
    # clamp on the host what the reference clamps (fusion.py:432: bound = max(1, min(modulus // 2, norm_bound))): ctypes truncates silently
    norm_bound = max(1, min(int(norm_bound), Q // 2))
    if not -(1 << 20) < int(log2_bias) < (1 << 20):
        raise ValueError("log2_bias out of range")
    rc = self.lib.fz_decode_poly(self._h, _ptr(data), C.c_size_t(stride), C.c_size_t(blen), C.c_size_t(n), C.c_longlong(log2_bias), C.c_longlong(norm_bound),
                                 C.c_int(weight_bound), _ptr(out), self._stream())
    if rc == -3:
        raise ValueError(self.lib.fz_last_error().decode())
    self._chk(rc)
    return out


def hash_ch(self, vk_l, vk_r, digests) -> torch.Tensor:
    vk_l, vk_r, digests = self._dev(vk_l), self._dev(vk_r), self._u8(digests)
    n = vk_l.numel() // self.d
    self.lib.fz_hash_ch_workspace.restype = C.c_size_t
    ws = self._ws(self.lib.fz_hash_ch_workspace(self._h, C.c_size_t(n)))
    out = torch.empty((n, self.d), dtype=torch.int32, device=self.device)
    self._chk(self.lib.fz_hash_ch(self._h, _ptr(vk_l), _ptr(vk_r), _ptr(digests), C.c_size_t(n), _ptr(out), _ptr(ws), C.c_size_t(ws.numel()), self._stream()))
    return out


def _agg_off(self, counts):
    off = [0]
    for c in counts:
        off.append(off[-1] + int(c))
    return torch.tensor(off, dtype=torch.int64, device=self.device), off[-1], (max(counts) if len(counts) else 0)


def hash_ag(self, vk_l_s, vk_r_s, dg_s, c_hat_s, counts) -> torch.Tensor:
    """counts[j] = number of (already sorted) signers of aggregate j."""
    vk_l_s, vk_r_s, dg_s, c_hat_s = self._dev(vk_l_s), self._dev(vk_r_s), self._u8(dg_s), self._dev(c_hat_s)
    agg_off, n_total, max_keys = self._agg_off(counts)
    self.lib.fz_hash_ag_workspace.restype = C.c_size_t
    ws = self._ws(self.lib.fz_hash_ag_workspace(self._h, C.c_size_t(n_total), C.c_size_t(len(counts)), C.c_size_t(max_keys)))
    out = torch.empty((n_total, self.d), dtype=torch.int32, device=self.device)
    self._chk(self.lib.fz_hash_ag(self._h, _ptr(vk_l_s), _ptr(vk_r_s), _ptr(dg_s), _ptr(c_hat_s), _ptr(agg_off), C.c_size_t(len(counts)), C.c_size_t(n_total),
                                  C.c_size_t(max_keys), _ptr(out), _ptr(ws), C.c_size_t(ws.numel()), self._stream()))
    return out


def sort_keys(self, vk_l, vk_r, counts, sel=None) -> torch.Tensor:
    vk_l, vk_r = self._dev(vk_l), self._dev(vk_r)
    agg_off, n_total, max_keys = self._agg_off(counts)
    if sel is not None:
        sel = sel.to(self.device, dtype=torch.int32).contiguous()
    perm = torch.empty(n_total, dtype=torch.int32, device=self.device)
    self._chk(self.lib.fz_sort_keys(self._h, _ptr(vk_l), _ptr(vk_r), _ptr(sel), _ptr(agg_off), C.c_size_t(len(counts)), C.c_size_t(max_keys), _ptr(perm),
                                    self._stream()))
    return perm


def gather_sorted(self, vk_l, vk_r, digests, perm, sel=None):
    vk_l, vk_r, digests = self._dev(vk_l), self._dev(vk_r), self._u8(digests)
    n_total = perm.numel()
    if sel is not None:
        sel = sel.to(self.device, dtype=torch.int32).contiguous()
    vl = torch.empty((n_total, self.d), dtype=torch.int32, device=self.device)
    vr = torch.empty((n_total, self.d), dtype=torch.int32, device=self.device)
    dg = torch.empty((n_total, 32), dtype=torch.uint8, device=self.device)
    self._chk(self.lib.fz_gather_sorted(self._h, _ptr(vk_l), _ptr(vk_r), _ptr(digests), _ptr(sel), _ptr(perm), C.c_size_t(n_total), _ptr(vl), _ptr(vr), _ptr(dg),
                                        self._stream()))
    return vl, vr, dg


def aggregate_sum(self, sigs, perm, alpha, counts) -> torch.Tensor:
    sigs, alpha = self._dev(sigs), self._dev(alpha)
    agg_off, n_total, max_keys = self._agg_off(counts)
    self.lib.fz_aggregate_sum_workspace.restype = C.c_size_t
    ws = self._ws(self.lib.fz_aggregate_sum_workspace(self._h, C.c_size_t(len(counts)), C.c_size_t(max_keys)))
    out = torch.empty((len(counts), self.ell, self.d), dtype=torch.int32, device=self.device)
    self._chk(self.lib.fz_aggregate_sum(self._h, _ptr(sigs), _ptr(perm), _ptr(alpha), _ptr(agg_off), C.c_size_t(len(counts)), C.c_size_t(max_keys), _ptr(out),
                                        _ptr(ws), C.c_size_t(ws.numel()), self._stream()))
    return out


def fold_partials(self, parts: torch.Tensor, n_parts: int, n_aggs: int) -> torch.Tensor:
    parts = self._dev(parts)
    out = torch.empty((n_aggs, self.ell, self.d), dtype=torch.int32, device=self.device)
    self._chk(self.lib.fz_fold_partials(self._h, _ptr(parts), C.c_int(n_parts), C.c_size_t(n_aggs), _ptr(out), self._stream()))
    return out


def verify_check(self, A, vk_l_s, vk_r_s, c_hat_s, alpha_s, agg, counts) -> torch.Tensor:
    A, vk_l_s, vk_r_s, c_hat_s, alpha_s, agg = (self._dev(t) for t in (A, vk_l_s, vk_r_s, c_hat_s, alpha_s, agg))
    agg_off, n_total, max_keys = self._agg_off(counts)
    self.lib.fz_verify_check_workspace.restype = C.c_size_t
    ws = self._ws(self.lib.fz_verify_check_workspace(self._h, C.c_size_t(len(counts))))
    codes = torch.empty(len(counts), dtype=torch.uint8, device=self.device)
    self._chk(self.lib.fz_verify_check(self._h, _ptr(A), _ptr(vk_l_s), _ptr(vk_r_s), _ptr(c_hat_s), _ptr(alpha_s), _ptr(agg), _ptr(agg_off),
                                       C.c_size_t(len(counts)), _ptr(codes), _ptr(ws), C.c_size_t(ws.numel()), self._stream()))
    return codes


# ---- windowed aggregation hash + split verify (device offsets; no host synchronisation) ----------------------
def hash_ag_begin(self, agg_off: torch.Tensor, n_aggs: int) -> torch.Tensor:
    """One resumable sponge record per aggregate (fusion.py:591 is one serial SHAKE256 per aggregate)."""
    self.lib.fz_hash_ag_state_bytes.restype = C.c_size_t
    state = torch.empty(max(n_aggs, 1) * int(self.lib.fz_hash_ag_state_bytes(self._h)), dtype=torch.uint8, device=self.device)
    self._chk(self.lib.fz_hash_ag_begin(self._h, _ptr(state), _ptr(agg_off), C.c_size_t(n_aggs), self._stream()))
    return state


def hash_ag_absorb(self, state, vk_l_w, vk_r_w, dg_w, c_hat_w, win_off: torch.Tensor, n_aggs: int) -> None:
    self._chk(self.lib.fz_hash_ag_absorb(self._h, _ptr(state), _ptr(vk_l_w), _ptr(vk_r_w), _ptr(dg_w), _ptr(c_hat_w), _ptr(win_off), C.c_size_t(n_aggs),
                                         self._stream()))


def hash_ag_squeeze(self, state, win_off: torch.Tensor, n_aggs: int, n_window: int) -> torch.Tensor:
    out = torch.empty((n_window, self.d), dtype=torch.int32, device=self.device)
    self._chk(self.lib.fz_hash_ag_squeeze(self._h, _ptr(state), _ptr(win_off), C.c_size_t(n_aggs), C.c_size_t(n_window), _ptr(out), self._stream()))
    return out


def aggregate_sum_off(self, sigs, rows, alpha, agg_off: torch.Tensor, n_aggs: int, max_keys: int) -> torch.Tensor:
    """fz_aggregate_sum with device offsets: out[j] = sum over p in [agg_off[j], agg_off[j+1]) of sigs[rows[p]] * alpha[p]."""
    self.lib.fz_aggregate_sum_workspace.restype = C.c_size_t
    ws = self._ws(self.lib.fz_aggregate_sum_workspace(self._h, C.c_size_t(n_aggs), C.c_size_t(max_keys)))
    out = torch.empty((n_aggs, self.ell, self.d), dtype=torch.int32, device=self.device)
    self._chk(self.lib.fz_aggregate_sum(self._h, _ptr(sigs), _ptr(rows), _ptr(alpha), _ptr(agg_off), C.c_size_t(n_aggs), C.c_size_t(max_keys), _ptr(out),
                                        _ptr(ws), C.c_size_t(ws.numel()), self._stream()))
    return out


def verify_target(self, vk_l_w, vk_r_w, c_hat_w, alpha_w, win_off: torch.Tensor, n_aggs: int) -> torch.Tensor:
    out = torch.empty((n_aggs, self.d), dtype=torch.int32, device=self.device)
    self._chk(self.lib.fz_verify_target(self._h, _ptr(vk_l_w), _ptr(vk_r_w), _ptr(c_hat_w), _ptr(alpha_w), _ptr(win_off), C.c_size_t(n_aggs), _ptr(out),
                                        self._stream()))
    return out


def verify_finish(self, A, target, agg) -> torch.Tensor:
    A, target, agg = self._dev(A), self._dev(target), self._dev(agg)
    n_aggs = target.numel() // self.d
    self.lib.fz_verify_finish_workspace.restype = C.c_size_t
    ws = self._ws(self.lib.fz_verify_finish_workspace(self._h, C.c_size_t(n_aggs)))
    codes = torch.empty(n_aggs, dtype=torch.uint8, device=self.device)
    self._chk(self.lib.fz_verify_finish(self._h, _ptr(A), _ptr(target), _ptr(agg), C.c_size_t(n_aggs), _ptr(codes), _ptr(ws), C.c_size_t(ws.numel()),
                                        self._stream()))
    return codes


# ---- ring-fed windowed aggregation hash: text generation and decoding off the serial chain --------------------
def hash_ag_ring(self, n_aggs: int, max_window: int):
    """(ring, ring_cap, text_pos): per-aggregate byte rings for windows of <= max_window signers per aggregate."""
    self.lib.fz_hash_ag_ring_bytes.restype = C.c_size_t
    cap = int(self.lib.fz_hash_ag_ring_bytes(self._h, C.c_size_t(max_window)))
    ring = torch.empty(max(n_aggs, 1) * cap, dtype=torch.uint8, device=self.device)
    return ring, cap, torch.zeros(max(n_aggs, 1), dtype=torch.int64, device=self.device)


def hash_ag_xof_buffer(self, n_aggs: int, max_window: int):
    self.lib.fz_hash_ag_xof_bytes.restype = C.c_size_t
    cap = int(self.lib.fz_hash_ag_xof_bytes(self._h, C.c_size_t(max_window)))
    return torch.empty(max(n_aggs, 1) * cap, dtype=torch.uint8, device=self.device), cap


def hash_ag_text(self, vk_l_w, vk_r_w, dg_w, c_hat_w, win_off, agg_of, agg_off, n_aggs: int, n_window: int, w0: int, text_pos, ring, ring_cap: int, state=None):
    """Writes the text window w0 adds to every aggregate's message into the rings; returns the message lengths after it (bit 63: text complete;
    bit 62, only with `state`: the text overran bytes the aggregate's sponge had not absorbed -- the absorb calls then poison that sponge)."""
    self.lib.fz_hash_ag_text_workspace.restype = C.c_size_t
    ws = self._ws(self.lib.fz_hash_ag_text_workspace(self._h, C.c_size_t(n_window)))
    wend = torch.empty(max(n_aggs, 1), dtype=torch.int64, device=self.device)
    self._chk(self.lib.fz_hash_ag_text(self._h, _ptr(vk_l_w), _ptr(vk_r_w), _ptr(dg_w), _ptr(c_hat_w), _ptr(win_off), _ptr(agg_of), _ptr(agg_off),
                                       C.c_size_t(n_aggs), C.c_size_t(n_window), C.c_uint32(w0), _ptr(text_pos), _ptr(ring), C.c_size_t(ring_cap), _ptr(state), _ptr(wend),
                                       _ptr(ws), C.c_size_t(ws.numel()), self._stream()))
    return wend


def hash_ag_absorb_ring(self, state, ring, ring_cap: int, wend, win_off, n_aggs: int) -> None:
    self._chk(self.lib.fz_hash_ag_absorb_ring(self._h, _ptr(state), _ptr(ring), C.c_size_t(ring_cap), _ptr(wend), _ptr(win_off), C.c_size_t(n_aggs),
                                              self._stream()))


def hash_ag_squeeze_xof(self, state, win_off, n_aggs: int, xof, xof_cap: int) -> torch.Tensor:
    bpos = torch.empty(max(n_aggs, 1), dtype=torch.int32, device=self.device)
    self._chk(self.lib.fz_hash_ag_squeeze_xof(self._h, _ptr(state), _ptr(win_off), C.c_size_t(n_aggs), _ptr(xof), C.c_size_t(xof_cap), _ptr(bpos),
                                              self._stream()))
    return bpos


def hash_ag_decode(self, xof, xof_cap: int, bpos, win_off, agg_of, n_window: int) -> torch.Tensor:
    out = torch.empty((n_window, self.d), dtype=torch.int32, device=self.device)
    self._chk(self.lib.fz_hash_ag_decode(self._h, _ptr(xof), C.c_size_t(xof_cap), _ptr(bpos), _ptr(win_off), _ptr(agg_of), C.c_size_t(n_window), _ptr(out),
                                         self._stream()))
    return out


def hash_ag_absorb_sliced(self, state, ring, ring_cap: int, wend, n_aggs: int, first_active: int, n_active: int, budget_blocks: int) -> None:
    self._chk(self.lib.fz_hash_ag_absorb_sliced(self._h, _ptr(state), _ptr(ring), C.c_size_t(ring_cap), _ptr(wend), C.c_size_t(n_aggs),
                                                C.c_size_t(first_active), C.c_size_t(n_active), C.c_uint64(budget_blocks), self._stream()))


def hash_ag_open_sponges(self, state, n_aggs: int) -> int:
    """Number of aggregates whose message has not been closed yet (phase word of the state records); synchronises."""
    return int((state.view(torch.int32).view(-1, 128)[:n_aggs, 107] != 1).sum().item())


def hash_ag_fast(self, vk_l_s, vk_r_s, dg_s, c_hat_s, counts) -> torch.Tensor:
    """fusion.py:632-652 like hash_ag, through the ring-fed kernels in one window: the message text is written by a warp per signer,
    the sponge chains only permute, the squeezed blocks are decoded in parallel (0.64 s instead of 1.0 s for one capacity aggregate)."""
    vk_l_s, vk_r_s, dg_s, c_hat_s = self._dev(vk_l_s), self._dev(vk_r_s), self._u8(dg_s), self._dev(c_hat_s)
    agg_off, n_total, max_keys = self._agg_off(counts)
    n_aggs = len(counts)
    if n_total == 0:
        return torch.empty((0, self.d), dtype=torch.int32, device=self.device)
    agg_of = torch.repeat_interleave(torch.arange(n_aggs, dtype=torch.int32, device=self.device),
                                     torch.as_tensor(list(counts), dtype=torch.int64, device=self.device), output_size=n_total)
    state = self.hash_ag_begin(agg_off, n_aggs)
    ring, cap, text_pos = self.hash_ag_ring(n_aggs, max_keys)
    wend = self.hash_ag_text(vk_l_s, vk_r_s, dg_s, c_hat_s, agg_off, agg_of, agg_off, n_aggs, n_total, 0, text_pos, ring, cap)
    self.hash_ag_absorb_ring(state, ring, cap, wend, agg_off, n_aggs)
    xof, xcap = self.hash_ag_xof_buffer(n_aggs, max_keys)
    bpos = self.hash_ag_squeeze_xof(state, agg_off, n_aggs, xof, xcap)
    return self.hash_ag_decode(xof, xcap, bpos, agg_off, agg_of, n_total)


def _seed_keys(self, seeds):
    """Python ints (any sign, up to 256 bits) -> ([n, 8] uint32 limbs, [n] int32 limb counts) on the device."""
    import numpy as np

    n = len(seeds)
    keys = np.zeros((n, 8), dtype=np.uint32)
    klen = np.zeros(n, dtype=np.int32)
    for i, s in enumerate(seeds):
        v = abs(int(s))
        if v >> 256:
            raise ValueError("seeds above 256 bits are not supported by the device sampler")
        k = 0
        while True:
            keys[i, k] = v & 0xFFFFFFFF
            v >>= 32
            k += 1
            if not v:
                break
        klen[i] = k
    return torch.from_numpy(keys.view(np.int32)).to(self.device), torch.from_numpy(klen).to(self.device)


def sample_coef(self, seeds, norm_bound: int = 52, weight_bound=None) -> torch.Tensor:
    """algebra/polynomials.py:436-467 for a list of integer seeds -> [n, d]."""
    keys, klen = self._seed_keys(seeds)
    out = torch.empty((len(seeds), self.d), dtype=torch.int32, device=self.device)
    wb = self.d if weight_bound is None else weight_bound
    self._chk(self.lib.fz_sample_coef(self._h, _ptr(keys), C.c_int(8), _ptr(klen), C.c_size_t(len(seeds)), C.c_int(norm_bound), C.c_int(wb), _ptr(out),
                                      self._stream()))
    return out


def sample_ntt(self, seeds) -> torch.Tensor:
    """algebra/polynomials.py:470-488 for a list of integer seeds -> [n, d]."""
    keys, klen = self._seed_keys(seeds)
    out = torch.empty((len(seeds), self.d), dtype=torch.int32, device=self.device)
    self._chk(self.lib.fz_sample_ntt(self._h, _ptr(keys), C.c_int(8), _ptr(klen), C.c_size_t(len(seeds)), _ptr(out), self._stream()))
    return out


def keygen(self, A, seeds, want_sk: bool = True, out=None):
    """fusion.py:338-373 for u64 seeds.  Returns (sk_l, sk_r, sk_hat, vk_l, vk_r); sk_l / sk_r are None unless want_sk.
    `out` may supply (sk_hat [n, 2, d], vk_l [n, d], vk_r [n, d]) buffers to write into."""
    A = self._dev(A)
    if not isinstance(seeds, torch.Tensor):
        import numpy as np
        seeds = torch.from_numpy(np.asarray(seeds, dtype=np.uint64).view(np.int64))
    seeds = seeds.to(self.device).contiguous()
    n = seeds.numel()
    self.lib.fz_keygen_workspace.restype = C.c_size_t
    ws = self._ws(self.lib.fz_keygen_workspace(self._h, C.c_size_t(n)))
    sk_l = torch.empty((n, self.ell, self.d), dtype=torch.int32, device=self.device) if want_sk else None
    sk_r = torch.empty((n, self.ell, self.d), dtype=torch.int32, device=self.device) if want_sk else None
    if out is not None:
        sk_hat, vk_l, vk_r = out
        for t, shape in ((sk_hat, (n, 2, self.d)), (vk_l, (n, self.d)), (vk_r, (n, self.d))):
            if tuple(t.shape) != shape or t.dtype != torch.int32 or not t.is_contiguous() or t.device != self.device:
                raise ValueError("keygen: out buffers must be contiguous int32 device tensors of shapes [n, 2, d], [n, d], [n, d]")
    else:
        sk_hat = torch.empty((n, 2, self.d), dtype=torch.int32, device=self.device)
        vk_l = torch.empty((n, self.d), dtype=torch.int32, device=self.device)
        vk_r = torch.empty((n, self.d), dtype=torch.int32, device=self.device)
    self._chk(self.lib.fz_keygen(self._h, _ptr(A), _ptr(seeds), C.c_size_t(n), _ptr(sk_l), _ptr(sk_r), _ptr(sk_hat), _ptr(vk_l), _ptr(vk_r), _ptr(ws),
                                 C.c_size_t(ws.numel()), self._stream()))
    return sk_l, sk_r, sk_hat, vk_l, vk_r


def keygen_rows(self, A, row_seeds, out=None):
    """fz_keygen_rows: the general ell-row keygen (fusion.py:338-373 without quirk Q1).  row_seeds: [n, 2, ell] u64 (numpy array or an
    int64 device tensor holding the same bits): one sampler seed per row of sk_l (side 0) and sk_r (side 1).
    Returns (sk_l, sk_r, vk_l, vk_r) with sk_*: [n, ell, d]; `out` may supply (sk_l, sk_r) buffers to reuse."""
    A = self._dev(A)
    if not isinstance(row_seeds, torch.Tensor):
        import numpy as np
        row_seeds = torch.from_numpy(np.ascontiguousarray(row_seeds, dtype=np.uint64).view(np.int64))
    row_seeds = row_seeds.to(self.device).contiguous()
    if row_seeds.dim() != 3 or row_seeds.shape[1] != 2 or row_seeds.shape[2] != self.ell:
        raise ValueError(f"keygen_rows: row_seeds must be [n, 2, {self.ell}]")
    n = row_seeds.shape[0]
    self.lib.fz_keygen_rows_workspace.restype = C.c_size_t
    ws = self._ws(self.lib.fz_keygen_rows_workspace(self._h, C.c_size_t(n)))
    if out is not None:
        sk_l, sk_r = out[0][:n], out[1][:n]
    else:
        sk_l = torch.empty((n, self.ell, self.d), dtype=torch.int32, device=self.device)
        sk_r = torch.empty((n, self.ell, self.d), dtype=torch.int32, device=self.device)
    vk_l = torch.empty((n, self.d), dtype=torch.int32, device=self.device)
    vk_r = torch.empty((n, self.d), dtype=torch.int32, device=self.device)
    self._chk(self.lib.fz_keygen_rows(self._h, _ptr(A), _ptr(row_seeds), C.c_size_t(n), _ptr(sk_l), _ptr(sk_r), _ptr(vk_l), _ptr(vk_r), _ptr(ws),
                                      C.c_size_t(ws.numel()), self._stream()))
    return sk_l, sk_r, vk_l, vk_r


def keygen_sign(self, A, seeds: torch.Tensor, msgs: "PackedMessages", vk_l, vk_r, sig) -> None:
    """fz_keygen_sign: seeds (int64 view of u64, device), packed messages on the device; writes vk_l / vk_r [n, d] and sig [n, ell, d]."""
    n = seeds.numel()
    self.lib.fz_keygen_sign_workspace.restype = C.c_size_t
    ws = self._ws(self.lib.fz_keygen_sign_workspace(self._h, C.c_size_t(n)))
    self._chk(self.lib.fz_keygen_sign(self._h, _ptr(A), _ptr(seeds), _ptr(msgs.buf), _ptr(msgs.offs), C.c_size_t(n), _ptr(vk_l), _ptr(vk_r), _ptr(sig),
                                      _ptr(ws), C.c_size_t(ws.numel()), self._stream()))


def sign_mac_rows(self, sk_hat, c_hat, out=None) -> torch.Tensor:
    sk_hat, c_hat = self._dev(sk_hat), self._dev(c_hat)
    n = c_hat.numel() // self.d
    out = torch.empty((n, self.ell, self.d), dtype=torch.int32, device=self.device) if out is None else out
    self._chk(self.lib.fz_sign_mac_rows(self._h, _ptr(sk_hat), _ptr(c_hat), _ptr(out), C.c_size_t(n), self._stream()))
    return out


for _f in (_u8, _pack_messages, pack_messages, _ws, sha3_256, shake256, hash_msg, format_text, decode_poly, hash_ch, _agg_off, hash_ag, sort_keys, gather_sorted,
           aggregate_sum, fold_partials, verify_check, hash_ag_begin, hash_ag_absorb, hash_ag_squeeze, aggregate_sum_off,
           verify_target, verify_finish, hash_ag_ring, hash_ag_xof_buffer, hash_ag_text, hash_ag_absorb_ring, hash_ag_squeeze_xof,
           hash_ag_decode, hash_ag_absorb_sliced, hash_ag_open_sponges, hash_ag_fast, _seed_keys, sample_coef, sample_ntt, keygen, keygen_rows, keygen_sign, sign_mac_rows):
    setattr(Context, _f.__name__, _f)
