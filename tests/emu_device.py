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


def _on_host(factory):
    def make(*args, **kwargs):
        dev = kwargs.get("device")
        if dev is not None and str(dev).startswith("cuda"):
            kwargs["device"] = "cpu"
        return factory(*args, **kwargs)

    return make


class _TorchProxy:
    """`torch` as native.py sees it: torch.device("cuda", i) is the host."""

    def __getattr__(self, name):
        return getattr(torch, name)

    @staticmethod
    def device(*a, **k):
        return torch.device("cpu") if a and a[0] == "cuda" else torch.device(*a, **k)


class Permanent:
    """monkeypatch.setattr for a whole process (the subprocess runs of the reference's own test files)."""

    @staticmethod
    def setattr(obj, name, value, raising=True):
        setattr(obj, name, value)


def install(monkeypatch, native=None):
    """Patch torch.cuda and the native module (fusion_cryptography_b200.native unless given) for the running test; returns the native module."""
    if native is None:
        from fusion_cryptography_b200 import native

    emu = hostlib.load_emulated_library()
    monkeypatch.setattr(native, "LIB_PATH", emu._name)
    monkeypatch.setattr(native, "_lib", None)
    monkeypatch.setattr(native, "_contexts", {})
    if hasattr(native, "_rings"):
        monkeypatch.setattr(native, "_rings", {})
    monkeypatch.setattr(native, "torch", _TorchProxy())
    for name, value in (("is_available", lambda: True), ("init", lambda: None), ("current_device", lambda: 0), ("synchronize", lambda *a: None),
                        ("empty_cache", lambda: None), ("current_stream", lambda *a: _Stream()), ("Stream", _Stream), ("Event", _Event),
                        ("stream", _null), ("device", _null)):
        monkeypatch.setattr(torch.cuda, name, value)
    monkeypatch.setattr(torch.Tensor, "record_stream", lambda self, s: None, raising=False)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self.clone(), raising=False)   # moving inputs "to the device" copies them, as a real transfer does
    for fname in ("empty", "zeros", "ones", "full", "arange", "tensor", "randint"):           # ... or create them there
        monkeypatch.setattr(torch, fname, _on_host(getattr(torch, fname)))
    return native
