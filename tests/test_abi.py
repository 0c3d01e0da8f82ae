"""CPU-side checks of the drop-in boundary: the C-ABI library builds/loads and exports every
symbol include/fusion_b200.h declares (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fusion_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fz_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    for must in ("fz_init", "fz_destroy", "fz_last_error", "fz_ntt_fwd", "fz_ntt_inv", "fz_pointwise_mul", "fz_matvec"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from fusion_cryptography_b200 import build

    lib_path = build.build_library()
    lib = ctypes.CDLL(lib_path)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/fusion_b200.h but not exported"


def test_init_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fusion_cryptography_b200 import native

    with pytest.raises(native.FusionB200Error):
        native.Context(256)
    lib = native.load_library()
    h = ctypes.c_void_p()
    rc = lib.fz_init(0, 256, ctypes.byref(h))
    assert rc != 0 and lib.fz_last_error()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "fusion_cryptography_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} mentions the oracle"
