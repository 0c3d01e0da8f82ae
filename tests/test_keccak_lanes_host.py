"""csrc/fz_keccak_lanes.cuh -- Keccak-f[1600] spread over a lane pair (pairag, 16 sponges per warp) and over a whole warp (warpag, lane L holds Keccak
lane L), the two cooperative layouts of hash_ag's serial SHAKE256 chains -- run UNMODIFIED on an emulated warp (32 host threads; a shuffle is
publish / barrier / read / barrier, tests/host/keccak_lanes_host.cpp) and compared with the one-thread permutation of fz_keccak.cuh, which
tests/test_modarith_host.py pins against hashlib.  No GPU needed."""
import ctypes as C
import os

import numpy as np
import pytest

import hostlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    return hostlib.load("keccak_lanes_host", pthread=True)


def reference(lib, states, times):
    r = np.ascontiguousarray(states.copy())
    for _ in range(times):
        lib.h_keccak_f(r.ctypes.data_as(C.c_void_p), C.c_size_t(r.shape[0]))
    return r


def some_states(rng, n):
    s = rng.integers(0, 1 << 64, size=(n, 25), dtype=np.uint64)
    s[0] = 0                                    # the empty sponge
    if n > 1:
        s[1] = 0xFFFFFFFFFFFFFFFF
    if n > 2:
        s[2] = 0; s[2, 0] = 0x1F; s[2, 16] = 0x8000000000000000      # an empty SHAKE256 message after padding
    return s


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_warp_layout_equals_the_one_thread_permutation(lib, variant):
    rng = np.random.default_rng(31 + variant)
    for st in some_states(rng, 5):
        for times in (1, 3):
            got = np.ascontiguousarray(st.copy())
            lib.h_keccak_f_warp(got.ctypes.data_as(C.c_void_p), C.c_int(variant), C.c_int(times))
            assert np.array_equal(got, reference(lib, st[None, :], times)[0]), (variant, times)


def test_lane_pair_layout_equals_the_one_thread_permutation(lib):
    rng = np.random.default_rng(41)
    for times in (1, 2):
        st = some_states(rng, 16)               # 16 sponges on one warp, all different
        got = np.ascontiguousarray(st.copy())
        lib.h_keccak_f_pair(got.ctypes.data_as(C.c_void_p), C.c_int(times))
        assert np.array_equal(got, reference(lib, st, times)), times
