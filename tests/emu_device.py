"""A test-side stand-in for the GPU: the library compiled for the CUDA-on-CPU emulator (tests/hostlib.load_emulated_library) behind the product's own
Python layers -- native.Context, batch.Engine, the drop-in fusion.fusion / algebra.* modules -- with "device" tensors living in host memory.  Streams
and events are made trivial (work runs where it is enqueued, which is one of the orders the real streams allow).  Everything is done by patching
torch.cuda and native's library handle inside a pytest fixture; the package itself has no such mode and is not changed."""
import contextlib
import ctypes as C

import torch

import hostlib


class _Stream:
    cuda_stream = 0

    def __init__(self, *a, **k):
        pass

    def wait_stream(self, *a):
        pass

    def wait_event(self, *a):
        pass

    def synchronize(self):
        pass

    def record_event(self, *a):
        return _Event()


class _Event:
    def __init__(self, *a, **k):
        pass

    def record(self, *a):
        pass

    def synchronize(self):
        pass

    def wait(self, *a):
        pass

    def query(self):
        return True

    def elapsed_time(self, other):
        return 0.0


@contextlib.contextmanager
def _null(*a, **k):
    yield


def make_context(native, secpar):
    """native.Context without its CUDA checks: the emulated library's fz_init, tensors on the host."""
    ctx = object.__new__(native.Context)
    ctx.lib = native.load_library()
    ctx.secpar, ctx.device_index, ctx.device = secpar, 0, torch.device("cpu")
    h = C.c_void_p()
    rc = ctx.lib.fz_init(C.c_int(0), C.c_int(secpar), C.byref(h))
    assert rc == 0, ctx.lib.fz_last_error()
    ctx._h = h
    ctx.d, ctx.ell, ctx.capacity = ctx.param("d"), ctx.param("ell"), ctx.param("capacity")
    return ctx


def install(monkeypatch):
    """Patch torch.cuda and fusion_cryptography_b200.native for the running test; returns the native module."""
    from fusion_cryptography_b200 import native

    emu = hostlib.load_emulated_library()
    monkeypatch.setattr(native, "LIB_PATH", emu._name)
    monkeypatch.setattr(native, "_lib", None)
    monkeypatch.setattr(native, "_contexts", {})
    for name, value in (("is_available", lambda: True), ("init", lambda: None), ("current_device", lambda: 0), ("synchronize", lambda *a: None),
                        ("empty_cache", lambda: None), ("current_stream", lambda *a: _Stream()), ("Stream", _Stream), ("Event", _Event),
                        ("stream", _null), ("device", _null)):
        monkeypatch.setattr(torch.cuda, name, value)
    monkeypatch.setattr(torch.Tensor, "record_stream", lambda self, s: None, raising=False)
    for secpar in (128, 256):
        native._contexts[(secpar, 0)] = make_context(native, secpar)
    return native
