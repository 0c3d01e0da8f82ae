"""Builds and loads the host harnesses of tests/host/*.cpp -- the product's device headers (csrc/*.cuh) compiled with g++ for the CPU suite -- and the
small call helpers the host tests share.  Test infrastructure only."""
import ctypes as C
import os
import shutil
import subprocess
import tempfile
from math import ceil, log2

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INC = "/usr/local/cuda/include"
Q = 2147465729
_cache = {}
_dir = None


def load(name: str, cuda_headers: bool = False, pthread: bool = False, sources=None, std: str = "c++17", opt: str = "-O2") -> C.CDLL:
    """tests/host/<name>.cpp (or the listed sources under tests/host/) -> shared library, built once per process; skips the calling test when the
    toolchain is missing."""
    global _dir
    if name in _cache:
        return _cache[name]
    if shutil.which("g++") is None:
        pytest.skip("needs g++")
    if cuda_headers and not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("needs the CUDA headers")
    if _dir is None:
        _dir = tempfile.mkdtemp(prefix="fz_host_")
    out = os.path.join(_dir, name + ".so")
    srcs = [os.path.join(ROOT, "tests", "host", f) for f in (sources or [name + ".cpp"])]
    flags = [opt, "-std=" + std, "-fPIC", "-w"] + (["-pthread"] if pthread else []) + (["-I" + CUDA_INC] if cuda_headers else [])
    if len(srcs) == 1:
        subprocess.run(["g++"] + flags + ["-shared", "-o", out] + srcs, check=True)
    else:       # one object per source, compiled side by side, then linked
        objs = [os.path.join(_dir, name + "_" + os.path.basename(f)[:-4] + ".o") for f in srcs]
        procs = [subprocess.Popen(["g++"] + flags + ["-c", f, "-o", o]) for f, o in zip(srcs, objs)]
        if any(p.wait() for p in procs):
            raise RuntimeError("g++ failed building " + name)
        subprocess.run(["g++", "-shared"] + (["-pthread"] if pthread else []) + ["-o", out] + objs, check=True)
    _cache[name] = C.CDLL(out)
    return _cache[name]


def i32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.int64).astype(np.int32))


def ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


# ---- fz_format.cuh on the emulated warp
def fmt_vk(lib, secpar, vl, vr) -> str:
    buf = C.create_string_buffer(11264)
    a, b = i32(vl), i32(vr)
    n = lib.h_format_vk(C.c_int(1 if secpar == 256 else 0), C.c_int(len(a)), ptr(a), ptr(b), buf)
    return buf.raw[:n].decode()


def fmt_ch(lib, secpar, c) -> str:
    buf = C.create_string_buffer(11264)
    a = i32(c)
    n = lib.h_format_ch(C.c_int(1 if secpar == 256 else 0), C.c_int(len(a)), ptr(a), buf)
    return buf.raw[:n].decode()


# ---- fz_decoder.cuh
def byte_budget(log2_bias, d, norm_bound):
    bound = max(1, min(Q // 2, norm_bound))                               # fusion.py:432
    return bound, ceil((log2(bound) + 1 + log2_bias) / 8), ceil((log2(d) + log2_bias) / 8)     # fusion.py:433-434


def host_decode(lib, b: bytes, log2_bias, d, norm_bound, weight):
    bound, bpc, bpi = byte_budget(log2_bias, d, norm_bound)
    out = np.full(d, 99, dtype=np.int32)
    lib.h_decode(b, C.c_uint32(len(b)), C.c_int(d), C.c_int(weight), C.c_uint32(bound), C.c_int(bpc), C.c_int(bpi), ptr(out))
    return out.tolist()


# ---- the whole library on the CUDA-on-CPU emulator (tests/host/cuda_emu.h)
EMU_SOURCES = ["emu/api_emu.cpp", "emu/ntt_emu.cpp", "emu/ring_emu.cpp", "emu/sample_emu.cpp", "emu/scheme_emu.cpp", "emu/stream_emu.cpp", "emu/hash_emu.cpp",
               "emu/hashstream_emu.cpp", "emu/smem_emu.cpp", "emu/emu_runtime.cpp", "emu/pipeline_driver.cpp"]


def load_emulated_library() -> C.CDLL:
    """libfusion_b200 -- every csrc/*.cu, kernels and host code, plus the C ABI of include/fusion_b200.h -- compiled by g++ for the emulator; "device"
    pointers are host pointers."""
    pre = os.environ.get("FZ_EMU_LIBRARY")          # a subprocess of a test that has already built it
    if pre and os.path.exists(pre) and "libfusion_b200_emu" not in _cache:
        _cache["libfusion_b200_emu"] = C.CDLL(pre)
    lib = load("libfusion_b200_emu", cuda_headers=True, pthread=True, std="c++20", opt="-O1", sources=EMU_SOURCES)
    lib.fz_last_error.restype = C.c_char_p
    return lib
