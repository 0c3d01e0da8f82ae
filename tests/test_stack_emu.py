"""The product's whole stack on the emulated device (tests/emu_device.py): the drop-in `fusion.fusion` API and the batched Engine pipelines -- Python
mirror, ctypes binding, C ABI, launch code, kernels -- with the library compiled for the CUDA-on-CPU emulator.  The reference's README demo must give
the reference's digests (str(agg_sig) -> 130985dc...1d735105, the messages drawn from the global generator keygen re-seeded, quirk Q3),
and Engine's windowed aggregate / verify must equal the oracle.  No GPU needed; the package is not modified, torch.cuda is patched by the fixture."""
import hashlib
import json
import os
import random
import string

import numpy as np
import pytest
import torch

import emu_device

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def native(monkeypatch):
    return emu_device.install(monkeypatch)


def sha(x) -> str:
    return hashlib.sha256(str(x).encode()).hexdigest()


def test_readme_demo_through_the_drop_in_api(native):
    from fusion_cryptography_b200.fusion.fusion import aggregate, fusion_setup, keygen, sign, verify

    J = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))["demo"]
    a = fusion_setup(256, 42)                                   # README.md:27-37 / misc/demo.py of the reference
    keys = [keygen(a, 42) for _ in range(2)]
    messages = ["".join(random.choices(string.ascii_letters + string.digits, k=20)) for _ in range(2)]
    assert messages == J["messages"]                            # keygen leaves the global generator as the reference does (quirk Q3)
    sigs = [sign(a, k, m) for k, m in zip(keys, messages)]
    vks = [k[1] for k in keys]
    agg = aggregate(a, vks, messages, sigs)
    assert sha(a) == J["params_sha"] and sha(keys[0]) == J["key0_sha"] and sha(sigs) == J["sigs_sha"]
    assert sha(agg) == J["agg_sha"] == "130985dc44f2699f1fcffb8fc95a2fbe61682ea97b51ba1d03c2c83c1d735105"
    assert list(verify(a, vks, messages, agg)) == J["verify"] == [True, ""]


def test_engine_pipelines_against_the_oracle(native):
    from fusion_cryptography_b200.batch import Engine, Tuning
    from oracle import oracle as orc

    secpar = 128
    eng = Engine(secpar, seed=42, device=0)
    A = eng.A.numpy()
    assert np.array_equal(A, orc.setup_A(secpar, 42))
    n = 7
    seeds = (8 * 10**6 + 2 * np.arange(n)).astype(np.uint64)
    msgs = [f"stack-{i}".encode() for i in range(n)]
    vk_l, vk_r, sigs = eng.keygen_sign(seeds, msgs, chunk=3)     # three chunks: the two-stream pipeline with alternating buffers
    sk_l, sk_r, ovl, ovr = orc.keygen_batch(secpar, A, seeds)
    want_sigs = orc.sign_batch(secpar, sk_l, sk_r, ovl, ovr, msgs)
    assert np.array_equal(vk_l.numpy(), ovl) and np.array_equal(vk_r.numpy(), ovr) and np.array_equal(sigs.numpy(), want_sigs)
    counts = [4, 1, 2]
    want = np.stack([orc.aggregate(secpar, ovl[lo:hi], ovr[lo:hi], msgs[lo:hi], want_sigs[lo:hi]) for lo, hi in ((0, 4), (4, 5), (5, 7))])
    # the default chain layout (a warp per aggregate) once; the other window shapes with a thread per aggregate, which the emulator runs much faster
    for tuning, max_positions, layout in ((Tuning(), 3, 0), (Tuning(), None, 2), (Tuning(overlap=False), 6, 2), (Tuning(text_stream=True), 2, 2)):
        eng.tuning = tuning
        eng.ctx.set_option("hash_ag_impl", layout)
        aggs = eng.aggregate(vk_l, vk_r, msgs, sigs, counts, max_positions=max_positions)
        assert np.array_equal(aggs.numpy(), want), (tuning, max_positions)
        assert eng.verify(vk_l, vk_r, msgs, aggs, counts, max_positions=max_positions).tolist() == [0, 0, 0]
    bad = aggs.clone()
    bad[1, 3, 5] += 1
    assert eng.verify(vk_l, vk_r, msgs, bad, counts, max_positions=3).tolist() == [0, 3, 0]
    eng.ctx.set_option("hash_ag_impl", 0)


def test_graft_entry_smoke_on_the_emulated_device(native):
    """__graft_entry__.smoke() -- what the driver runs on cuda:0 before the bench -- with the emulated device under it: its own checks against the
    oracle (transforms, fused pass, matvec, keygen + sign, windowed aggregate / verify at both levels, a general ring) must hold here too."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("graft_entry_under_test", os.path.join(ROOT, "__graft_entry__.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.smoke()
