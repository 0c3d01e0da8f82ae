"""The ring-fed aggregation-hash pipeline of the library -- its KERNELS (agw_int_text / agw_len / agw_layout / agw_format, the resumable SHAKE256 chains
in all three layouts: a thread, a warp, a lane pair per aggregate, agw_squeeze_*, agw_decode) and its host LAUNCH code (launch_agw_text,
launch_agw_absorb, launch_agw_squeeze, launch_agw_decode), i.e. csrc/fz_hash.cu and csrc/fz_hashstream.cu as they are -- compiled by g++ on a small
CUDA-on-CPU emulator (tests/host/cuda_emu.h: the threads of a CTA are fibers of one host thread, barriers and shuffles are cooperative) and driven window by
window as Engine._hash_windows / _squeeze_windows drive it on the GPU (tests/host/emu/pipeline_driver.cpp).  The aggregation coefficients must equal
the ones the unmodified reference decoded (fusion.py:573-629) for its golden aggregate at both levels -- windows shorter than the aggregate exercise
the resumable sponge and the byte rings -- and the oracle's on a ragged batch of aggregates.  No GPU needed."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import hostlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LAYOUTS = {"thread": 2, "warp": 3, "pair": 4}


@pytest.fixture(scope="module")
def lib():
    lib = hostlib.load_emulated_library()
    lib.h_last_error.restype = C.c_char_p
    return lib


def run(lib, secpar, layout, counts, vk_l_s, vk_r_s, digests_s, c_hat_s, P):
    d = 64 if secpar == 128 else 256
    agg_off = np.zeros(len(counts) + 1, dtype=np.uint64)
    agg_off[1:] = np.cumsum(counts)
    vl, vr, ch = hostlib.i32(vk_l_s), hostlib.i32(vk_r_s), hostlib.i32(c_hat_s)
    dg = np.ascontiguousarray(np.frombuffer(b"".join(digests_s), dtype=np.uint8))
    out = np.full((int(agg_off[-1]), d), 777, dtype=np.int32)
    rc = lib.h_agw_pipeline(C.c_int(secpar), C.c_int(LAYOUTS[layout]), C.c_size_t(len(counts)), hostlib.ptr(agg_off), hostlib.ptr(vl), hostlib.ptr(vr),
                            hostlib.ptr(dg), hostlib.ptr(ch), C.c_int(P), hostlib.ptr(out))
    assert rc == 0, (rc, lib.h_last_error())
    return out.astype(np.int64)


@pytest.mark.parametrize("secpar,layout,P", [(128, "thread", 1), (128, "thread", 2), (128, "thread", 5), (256, "thread", 1), (256, "thread", 3),
                                               (128, "warp", 2), (128, "pair", 2), (256, "warp", 3), (256, "pair", 1)])
def test_golden_aggregate_through_the_kernels(lib, secpar, layout, P):
    J = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))
    Z = np.load(os.path.join(ROOT, "tests", "golden", "scheme_golden.npz"))
    order = Z[f"order_{secpar}"]
    digests = [int(J[f"prehash_{secpar}"][i]).to_bytes(32, "little") for i in order]
    got = run(lib, secpar, layout, [len(order)], Z[f"vk_l_{secpar}"][order], Z[f"vk_r_{secpar}"][order], digests, Z[f"c_hat_{secpar}"][order], P)
    assert np.array_equal(got, Z[f"alpha_coef_sorted_{secpar}"])


@pytest.mark.parametrize("layout,counts,P", [("thread", [1, 4, 2, 3], 2), ("thread", [3] * 33, 3), ("pair", [1, 2] + [1] * 15, 1)])
def test_ragged_batch_against_the_oracle(lib, layout, counts, P):
    """Several aggregates of different sizes in one batch (the pair case: 17 aggregates = two warps of the pair kernel, the second almost empty;
    the thread case with 33 aggregates: two CTAs); keys from the oracle's keygen, sorted by the oracle."""
    from oracle import oracle as orc

    secpar = 128
    A = orc.setup_A(secpar, 42)
    n = int(sum(counts))
    seeds = (5 * 10**6 + 2 * np.arange(n)).astype(np.uint64)
    _, _, vk_l, vk_r = orc.keygen_batch(secpar, A, seeds)
    msgs = [f"emu-{i}".encode() for i in range(n)]
    vl_s, vr_s, dg_s, ch_s, want = [], [], [], [], []
    lo = 0
    for c in counts:
        sl = slice(lo, lo + c)
        order = orc.sort_by_repr(secpar, vk_l[sl], vk_r[sl])
        kl, kr, ms = vk_l[sl][order], vk_r[sl][order], [msgs[lo + int(i)] for i in order]
        alpha_hat = orc.hash_ag(secpar, kl, kr, ms)
        want.append(orc.ntt_inv(secpar, alpha_hat))
        vl_s.append(kl); vr_s.append(kr)
        dg_s += [orc.hash_msg(secpar, m) for m in ms]
        ch_s.append(np.stack([orc.hash_ch(secpar, kl[i], kr[i], ms[i]) for i in range(c)]))
        lo += c
    got = run(lib, secpar, layout, counts, np.concatenate(vl_s), np.concatenate(vr_s), dg_s, np.concatenate(ch_s), P)
    assert np.array_equal(got, np.concatenate(want).astype(np.int64))
