"""The drop-in ``fusion.fusion`` API on the GPU vs the unmodified reference: byte-identical ``str()`` of
every returned object (digests recorded by tests/golden/make_golden.py), the README demo
(BASELINE config[0]), the reference's own relation tests and its error behaviour."""
import hashlib
import os
import random
import string
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sha(s):
    return hashlib.sha256(s.encode() if isinstance(s, str) else s).hexdigest()


@pytest.fixture(scope="module")
def ff():
    from fusion_cryptography_b200.fusion import fusion

    return fusion


@pytest.mark.parametrize("secpar", [128, 256])
def test_lifecycle_matches_reference_strings(ff, golden_json, scheme_golden, secpar):
    J, t = golden_json, str(secpar)
    params = ff.fusion_setup(secpar, 42)
    assert sha(str(params)) == J[f"params_str_sha_{t}"]
    seeds = [int(s) for s in scheme_golden[f"seeds_{t}"]]
    msgs = J[f"msgs_{t}"]
    keys = [ff.keygen(params, s) for s in seeds]
    assert [sha(str(k)) for k in keys] == J[f"key_str_sha_{t}"]
    assert str(keys[0][1]) == J[f"vk_str_{t}_0"]
    assert [str(ff.hash_message_to_int(params, m)) for m in msgs] == J[f"prehash_{t}"]
    n_ch = J[f"n_ch_{t}"]
    pre0 = ff.hash_message_to_int(params, msgs[0])
    assert ff.hash_vk_and_int_to_bytes(params, keys[0][1], pre0, n_ch).hex() == J[f"ch_xof_hex_{t}_0"]
    ch0 = ff.hash_ch(params, keys[0][1], msgs[0])
    assert str(ch0) == J[f"ch_str_{t}_0"]
    sigs = [ff.sign(params, k, m) for k, m in zip(keys, msgs)]
    assert [sha(str(s)) for s in sigs] == J[f"sig_str_sha_{t}"]
    vks = [k[1] for k in keys]
    agg = ff.aggregate(params, vks, msgs, sigs)
    assert sha(str(agg)) == J[f"agg_str_sha_{t}"]
    assert list(ff.verify(params, vks, msgs, agg)) == J[f"verify_ok_{t}"]
    # tamper one value (reference tests/test_fusion.py:861-873)
    agg.signature_hat.matrix[0][0].values[0] += 1
    assert list(ff.verify(params, vks, msgs, agg)) == J[f"verify_tamper_{t}"]
    agg.signature_hat.matrix[0][0].values[0] -= 1
    assert list(ff.verify(params, vks, msgs, agg)) == J[f"verify_ok_{t}"]
    wm = list(msgs)
    wm[0] = "other"
    assert list(ff.verify(params, vks, wm, agg)) == J[f"verify_wrongmsg_{t}"]
    assert list(ff.verify(params, vks, msgs[:-1], agg)) == J[f"verify_count_{t}"]
    assert ff.verify(params, vks * (params.capacity // len(vks) + 1), msgs, agg) == (False, "Too many keys.")
    agg1 = ff.aggregate(params, vks[:1], msgs[:1], sigs[:1])
    assert list(ff.verify(params, vks[:1], msgs[:1], agg1)) == J[f"verify1_{t}"]
    # hash_ag through the public helper (keys in the order given, as the reference)
    order = scheme_golden[f"order_{t}"].tolist()
    alphas = ff.hash_ag(params, [vks[i] for i in order], [msgs[i] for i in order])
    assert [a.alpha_hat.values for a in alphas] == scheme_golden[f"alpha_hat_sorted_{t}"].tolist()


def test_readme_demo_and_global_rng_side_effect(ff, golden_json):
    """README.md:20-71 / misc/demo.py: the demo's messages come from the global generator in the state
    keygen leaves behind."""
    J = golden_json["demo"]
    a = ff.fusion_setup(256, 42)
    keys = [ff.keygen(a, 42) for _ in range(2)]
    messages = ["".join(random.choices(string.ascii_letters + string.digits, k=20)) for _ in range(2)]
    assert messages == J["messages"]
    sigs = [ff.sign(a, k, m) for k, m in zip(keys, messages)]
    vks = [k[1] for k in keys]
    agg = ff.aggregate(a, vks, messages, sigs)
    assert sha(str(a)) == J["params_sha"] and sha(str(keys[0])) == J["key0_sha"]
    assert sha(str(sigs)) == J["sigs_sha"] and sha(str(agg)) == J["agg_sha"]
    assert list(ff.verify(a, vks, messages, agg)) == J["verify"]


def test_reference_relations(ff):
    """reference tests/test_fusion.py:308-349 (keygen), :694-731 (sign): A.sk == vk, A.z == vk_l*c + vk_r, norm / weight bounds."""
    from fusion_cryptography_b200.algebra.polynomials import transform

    for secpar in (128, 256):
        params = ff.fusion_setup(secpar, 8675309)
        sk, vk = ff.keygen(params, 8675309)
        assert params.public_challenge * sk.left_sk_hat == vk.left_vk_hat
        assert params.public_challenge * sk.right_sk_hat == vk.right_vk_hat
        for row in sk.left_sk_hat.matrix:
            c = transform(row[0])
            assert c.norm(p="infty") <= params.beta_sk and c.weight() <= params.omega_sk
        msg = "Hello, world!"
        sig = ff.sign(params, (sk, vk), msg)
        ch = ff.hash_ch(params, vk, msg)
        assert params.public_challenge * sig.signature_hat == vk.left_vk_hat * ch.c_hat + vk.right_vk_hat
        c = transform(ch.c_hat)
        assert c.norm(p="infty") <= params.beta_ch and c.weight() <= params.omega_ch
        assert transform(transform(c)) == c


def test_mocked_hash_helpers_route_through_module_attributes(ff, monkeypatch):
    """reference tests/test_fusion.py:368-435: the helpers call fusion.fusion.sha3_256 / shake_256 by name."""
    params = ff.fusion_setup(128, 1)

    class H:
        def __init__(self, data):
            self.data = data

        def digest(self, n=None):
            return (1234567890).to_bytes(32, "little") if n is None else b"expected_shake_256_result"

    calls = []
    monkeypatch.setattr(ff, "sha3_256", lambda data: calls.append(data) or H(data))
    assert ff.hash_message_to_int(params, "my_message") == 1234567890
    assert calls == [(params.sign_pre_hash_dst.decode("utf-8") + ",my_message").encode()]
    monkeypatch.undo()
    _, vk = ff.keygen(params, 2)
    calls = []
    monkeypatch.setattr(ff, "shake_256", lambda data: calls.append(data) or H(data))
    assert ff.hash_vk_and_int_to_bytes(params, vk, 1234567890, 1) == b"expected_shake_256_result"
    assert calls == [(params.sign_hash_dst.decode("utf-8") + "," + str(vk) + "," + "1234567890").encode()]
    monkeypatch.undo()
    one = [1] + [0] * (params.degree - 1)
    monkeypatch.setattr(ff, "decode_bytes_to_polynomial_coefficients", lambda **kw: one)
    ch = ff.hash_ch(params, vk, "m")
    from fusion_cryptography_b200.algebra.polynomials import PolynomialCoefficientRepresentation as Poly, transform
    assert ch.c_hat == transform(Poly(modulus=params.modulus, degree=params.degree, root=params.root, inv_root=params.inv_root,
                                      root_order=params.root_order, coefficients=one))


def test_error_behaviour(ff):
    from fusion_cryptography_b200.algebra.matrices import GeneralMatrix
    from fusion_cryptography_b200.algebra.polynomials import PolynomialNTTRepresentation as PolyNTT

    params = ff.fusion_setup(128, 3)
    with pytest.raises(ValueError, match="Too few bytes"):
        ff.decode_bytes_to_polynomial_coefficients(b=bytes(10), log2_bias=128, modulus=params.modulus, degree=64, norm_bound=1, weight_bound=27)
    with pytest.raises(ValueError, match="too short"):
        ff.parse_challenge(params, bytes(100))
    with pytest.raises(TypeError):
        PolyNTT(modulus=params.modulus, degree=64, root=params.root, inv_root=params.inv_root, root_order=128, values="x")
    with pytest.raises(ValueError):
        PolyNTT(modulus=params.modulus, degree=64, root=params.root, inv_root=params.inv_root, root_order=128, values=[0] * 63)
    with pytest.raises(ValueError):
        PolyNTT(modulus=params.modulus, degree=64, root=params.root + 1, inv_root=params.inv_root, root_order=128, values=[0] * 64)
    with pytest.raises(ValueError):
        GeneralMatrix(matrix=[])
    with pytest.raises(TypeError):
        ff.keygen(params, None)  # `seed + 1` on None, as in the reference (:352)
    p2 = ff.fusion_setup(256, 3)
    a = params.public_challenge.matrix[0][0]
    b = p2.public_challenge.matrix[0][0]
    with pytest.raises(NotImplementedError):
        a + b
    assert repr(PolyNTT) == "<class 'algebra.polynomials.PolynomialNTTRepresentation'>"  # hashed text, quirk Q9


def test_top_level_dropin_layout():
    """A user of the reference switches by putting the package directory on sys.path: `from fusion.fusion import ...`."""
    code = (
        "import sys, random, string, hashlib\n"
        f"sys.path.insert(0, {os.path.join(ROOT, 'fusion_cryptography_b200')!r})\n"
        "from fusion.fusion import fusion_setup, keygen, sign, aggregate, verify\n"
        "a = fusion_setup(256, 42)\n"
        "keys = [keygen(a, 42) for _ in range(2)]\n"
        "msgs = [''.join(random.choices(string.ascii_letters + string.digits, k=20)) for _ in range(2)]\n"
        "sigs = [sign(a, k, m) for k, m in zip(keys, msgs)]\n"
        "vks = [k[1] for k in keys]\n"
        "agg = aggregate(a, vks, msgs, sigs)\n"
        "print(msgs, hashlib.sha256(str(agg).encode()).hexdigest(), verify(a, vks, msgs, agg))\n"
    )
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "130985dc44f2699f1fcffb8fc95a2fbe61682ea97b51ba1d03c2c83c1d735105" in out.stdout  # SURVEY 8c demo digest
    assert "['3h7ujoQfqJ915fsSLGNU', 'eMPYPDeghzvdj9KQ949J']" in out.stdout and "(True, '')" in out.stdout


def test_wire_format_round_trip(ff):
    """keys / signatures / aggregates as int32 blobs; the decoded objects print and verify exactly like the originals"""
    from fusion_cryptography_b200 import wire

    for secpar in (128, 256):
        params = ff.fusion_setup(secpar, 7)
        keys = [ff.keygen(params, 1000 + 2 * i) for i in range(3)]
        msgs = [f"wire-{i}" for i in range(3)]
        sigs = [ff.sign(params, k, m) for k, m in zip(keys, msgs)]
        vks = [k[1] for k in keys]
        agg = ff.aggregate(params, vks, msgs, sigs)
        vks2 = [wire.vk_from_bytes(params, wire.vk_to_bytes(params, v)) for v in vks]
        sigs2 = [wire.signature_from_bytes(params, wire.signature_to_bytes(params, s)) for s in sigs]
        agg2 = wire.signature_from_bytes(params, wire.signature_to_bytes(params, agg))
        assert [str(v) for v in vks2] == [str(v) for v in vks] and [str(s) for s in sigs2] == [str(s) for s in sigs] and str(agg2) == str(agg)
        assert len(wire.vk_to_bytes(params, vks[0])) == 12 + 2 * params.degree * 4
        assert ff.verify(params, vks2, msgs, agg2) == (True, "")
        assert str(ff.aggregate(params, vks2, msgs, sigs2)) == str(agg)
        blob = bytearray(wire.signature_to_bytes(params, agg))
        with pytest.raises(ValueError):
            wire.signature_from_bytes(params, bytes(blob[:-4]))
        with pytest.raises(ValueError):
            wire.vk_from_bytes(params, bytes(blob))
        blob[12:16] = (2**31 - 1).to_bytes(4, "little")
        with pytest.raises(ValueError):
            wire.signature_from_bytes(params, bytes(blob))


@pytest.mark.parametrize("secpar", [128, 256])
def test_every_kat_kind_on_the_device(ff, secpar):
    """Every kind of row the reference's KAT generator writes (KATs/generate_KAT_values.py), on the CUDA path against digests recorded from the
    unmodified reference (tests/golden/make_kat_golden.py) -- including the kinds whose CSVs the reference's repository is missing, and the
    generator's (sk, vk) tuples as keys of hash_ag / hash_vks_and_ints_and_challs_to_bytes (only str(key) is hashed there)."""
    import json

    def sha(x):
        return hashlib.sha256(str(x).encode()).hexdigest()

    G = json.load(open(os.path.join(ROOT, "tests", "golden", "kat_kinds_golden.json")))[str(secpar)]
    p = ff.fusion_setup(secpar, G["seed_a"])
    assert sha(p) == G["setup"]
    msgs = [str(i) for i in range(10)]
    otks = [ff.keygen(p, s) for s in G["key_seeds"]]
    assert [sha(k) for k in otks] == G["keygen"]
    vks = [k[1] for k in otks]
    pre = [ff.hash_message_to_int(p, m) for m in msgs]
    assert [str(i) for i in pre] == G["prehash"]
    assert [sha(ff.hash_vk_and_int_to_bytes(p, v, i, G["n_ch"])) for v, i in zip(vks, pre)] == G["xof"]
    challs = [ff.hash_ch(p, v, m) for v, m in zip(vks, msgs)]
    assert [sha(c) for c in challs] == G["hash_ch"]
    sigs = [ff.sign(p, k, m) for k, m in zip(otks, msgs)]
    assert [sha(s) for s in sigs] == G["sign"]
    assert sha(ff.hash_vks_and_ints_and_challs_to_bytes(p, vks, pre, challs)) == G["ag_bytes_vks"]
    assert sha(ff.hash_vks_and_ints_and_challs_to_bytes(p, otks, pre, challs)) == G["ag_bytes_tuples"]
    assert sha(ff.hash_ag(p, vks, msgs)) == G["hash_ag_vks"]
    assert sha(ff.hash_ag(p, otks, msgs)) == G["hash_ag_tuples"]
    agg = ff.aggregate(p, vks, msgs, sigs)
    assert sha(agg) == G["aggregate"]
    assert list(ff.verify(p, vks, msgs, agg)) == G["verify"]


def test_key_objects_other_than_verification_keys(ff):
    """aggregate() only sorts by and hashes str(key) in the reference, so it accepts e.g. the (sk, vk) tuples of keygen; verify() reads the key's
    attributes and fails with the reference's AttributeError.  Digest recorded from the reference (fusion_setup(128, 5), keygen seeds 10, 12, 14)."""
    p = ff.fusion_setup(128, 5)
    otks = [ff.keygen(p, 10 + 2 * i) for i in range(3)]
    msgs = ["a", "b", "c"]
    sigs = [ff.sign(p, k, m) for k, m in zip(otks, msgs)]
    assert sha(str(ff.aggregate(p, otks, msgs, sigs))) == "f8a7b2a357b56258e5af70178c024968070a4a55bab59eb06009289f6033244d"
    with pytest.raises(AttributeError, match="'tuple' object has no attribute 'left_vk_hat'"):
        ff.verify(p, otks, msgs, sigs[0])


def test_differential_battery_on_the_device():
    """tests/refsuite/differential_cases.py (about 150 ordinary, edge and erroneous calls over the reference's three modules) in the drop-in layout on
    the CUDA path, against the outcomes recorded from the unmodified reference (tests/golden/make_differential_golden.py)."""
    import json

    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "fusion_cryptography_b200"))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "refsuite", "differential_cases.py")], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    got = json.loads(r.stdout.strip().splitlines()[-1])
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "differential_golden.json")))
    diff = {k: (v, got.get(k)) for k, v in want.items() if got.get(k) != v}
    assert not diff, diff
