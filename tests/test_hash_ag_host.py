"""The aggregation hash of the reference's golden aggregate recomposed on the CPU from the PRODUCT's device code compiled for the host: the message
pieces "(" str(vk) ", " str(prehash) ", " str(challenge) ")" come from csrc/fz_format.cuh on the emulated warp, the sorted order from
csrc/fz_repr_order.cuh, SHAKE256 from hashlib (the permutations themselves are pinned by tests/test_keccak_lanes_host.py), the coefficients from
the byte-driven decoder of csrc/fz_decoder.cuh -- against what the unmodified reference computed (fusion.py:573-629: the XOF digest and every
aggregation coefficient, with the decoder running out of bytes at secpar 256).  Only the list punctuation is stated here."""
import ctypes as C
import hashlib
import json
import os
from math import ceil, log2

import numpy as np
import pytest

import hostlib
from hostlib import fmt_ch, fmt_vk, host_decode

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LEVEL = {128: (64, 35), 256: (256, 60)}      # secpar -> (degree, omega_ag)   (fusion.py:28-29)


@pytest.fixture(scope="module")
def fmt():
    return hostlib.load("format_host", pthread=True)


@pytest.fixture(scope="module")
def dec():
    return hostlib.load("decoder_host")


@pytest.fixture(scope="module")
def order():
    return hostlib.load("repr_order_host")


@pytest.mark.parametrize("secpar", [128, 256])
def test_golden_aggregate_hash(fmt, dec, order, secpar):
    J = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))
    Z = np.load(os.path.join(ROOT, "tests", "golden", "scheme_golden.npz"))
    d, omega = LEVEL[secpar]
    vk_l, vk_r, c_hat, pre = Z[f"vk_l_{secpar}"], Z[f"vk_r_{secpar}"], Z[f"c_hat_{secpar}"], J[f"prehash_{secpar}"]
    n = vk_l.shape[0]
    # the sort step (fusion.py:661-663) with the product's comparison
    a = np.ascontiguousarray(vk_l.astype(np.int32)); b = np.ascontiguousarray(vk_r.astype(np.int32))
    perm = np.empty(n, dtype=np.uint32)
    order.h_rank_sort(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), C.c_int(n), C.c_int(d), perm.ctypes.data_as(C.c_void_p))
    assert perm.tolist() == Z[f"order_{secpar}"].tolist()
    # the message (fusion.py:586-590): dst "," str([(vk, int, challenge), ...])
    pieces = ["(" + fmt_vk(fmt, secpar, vk_l[i], vk_r[i]) + ", " + pre[i] + ", " + fmt_ch(fmt, secpar, c_hat[i]) + ")" for i in perm.tolist()]
    msg = bytes([3 if secpar == 256 else 1, 2]) + b"," + ("[" + ", ".join(pieces) + "]").encode()
    bpc, bpi = ceil((log2(1) + 1 + secpar) / 8), ceil((log2(d) + secpar) / 8)
    n_ag = ceil(omega / 8) + (bpc + bpi) * omega                             # fusion.py:583-585
    assert n * n_ag == J[f"ag_xof_len_{secpar}"]
    xof = hashlib.shake_256(msg).digest(n * n_ag)
    assert hashlib.sha256(xof).hexdigest() == J[f"ag_xof_sha_{secpar}"]
    # the coefficients (fusion.py:594-629); at secpar 256 the shuffle needs more bytes than n_ag holds
    got = [host_decode(dec, xof[k * n_ag:(k + 1) * n_ag], secpar, d, 1, omega) for k in range(n)]
    assert np.array_equal(np.array(got, dtype=np.int64), Z[f"alpha_coef_sorted_{secpar}"])
