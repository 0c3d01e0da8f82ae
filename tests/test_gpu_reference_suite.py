"""SURVEY 8(f) item 2, on the CUDA path: the reference's OWN unit tests (all four files, 53 tests) and its KAT generator, UNMODIFIED, against the
drop-in -- `algebra` and `fusion` resolve to fusion_cryptography_b200's packages, every arithmetic / hashing / sampling call lands in
libfusion_b200.so on the GPU.  The reference's files come from baseline/_ref/_checkout (tools/install_reference.py copies them there from the
reference checkout; the directory is git-ignored and travels to the GPU box with the snapshot) -- nothing of the reference is committed here.
The KAT generator is run twice on the same box, once against the unmodified reference (baseline/_ref, CPython on the host cores) and once
against the drop-in: all 18 CSV files must be byte-identical."""
import hashlib
import os
import shutil
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
CHECKOUT = os.path.join(REF, "_checkout")
DROPIN = os.path.join(ROOT, "fusion_cryptography_b200")
FILES = ["test_ntt.py", "test_polynomials.py", "test_matrices.py", "test_fusion.py"]
need_ref = pytest.mark.skipif(not all(os.path.exists(os.path.join(CHECKOUT, "tests", f)) for f in FILES),
                              reason="baseline/_ref/_checkout is absent (run tools/install_reference.py where the reference checkout exists)")


@need_ref
def test_reference_unit_tests_pass_on_the_cuda_path(tmp_path):
    tdir = tmp_path / "tests"
    tdir.mkdir()
    (tdir / "__init__.py").write_text("")   # test_matrices.py imports `tests.test_ntt`
    shutil.copy(os.path.join(ROOT, "tests", "refsuite", "mocker_shim.py"), tdir / "conftest.py")
    for f in FILES:
        shutil.copy(os.path.join(CHECKOUT, "tests", f), tdir / f)
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([DROPIN, str(tmp_path), str(tdir)]), PYTHONDONTWRITEBYTECODE="1", FUSION_B200_TRACE_LOADS="1")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider"] + [str(tdir / f) for f in FILES], cwd=str(tmp_path), env=env,
                       capture_output=True, text=True, timeout=1500)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert "53 passed" in r.stdout and "failed" not in r.stdout, tail
    # the modules under test were the drop-in's, and the CUDA library was loaded by that process
    probe = subprocess.run([sys.executable, "-c", "import algebra.ntt, fusion.fusion, sys; print(algebra.ntt.__file__); print(fusion.fusion.__file__); "
                            "print(fusion.fusion._native.load_library()._name)"], env=env, capture_output=True, text=True, timeout=300)
    assert probe.returncode == 0, probe.stderr[-2000:]
    lines = probe.stdout.strip().splitlines()
    assert lines[0].startswith(DROPIN) and lines[1].startswith(DROPIN) and lines[2].endswith("libfusion_b200.so"), lines


@need_ref
def test_reference_kat_generator_writes_identical_files_on_the_cuda_path(tmp_path):
    gen = os.path.join(CHECKOUT, "KATs", "generate_KAT_values.py")
    runner = ("import random; random.seed(20231018)\n"
              f"exec(compile(open({gen!r}).read(), 'generate_KAT_values', 'exec'))\n")
    digests = {}
    for name, path in (("reference", REF), ("dropin", DROPIN)):
        wd = tmp_path / ("run_" + name)
        wd.mkdir()
        env = dict(os.environ, PYTHONPATH=path, PYTHONDONTWRITEBYTECODE="1")
        r = subprocess.run([sys.executable, "-c", runner], cwd=str(wd), env=env, capture_output=True, text=True, timeout=1500)
        assert r.returncode == 0, (name, (r.stdout + r.stderr)[-3000:])
        out = wd / "KAT_values"
        digests[name] = {f: hashlib.sha256((out / f).read_bytes()).hexdigest() for f in sorted(os.listdir(out))}
    assert len(digests["reference"]) == 18
    assert digests["dropin"] == digests["reference"]


@need_ref
def test_differential_battery_live_against_the_reference(tmp_path):
    """tests/refsuite/differential_cases.py with the names resolving to baseline/_ref (CPython) and to the drop-in (CUDA), on the same box."""
    import json

    script = os.path.join(ROOT, "tests", "refsuite", "differential_cases.py")
    out = {}
    for name, path in (("reference", REF), ("dropin", DROPIN)):
        env = dict(os.environ, PYTHONPATH=path, PYTHONDONTWRITEBYTECODE="1")
        r = subprocess.run([sys.executable, script], cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=1500)
        assert r.returncode == 0, (name, (r.stdout + r.stderr)[-3000:])
        out[name] = json.loads(r.stdout.strip().splitlines()[-1])
    assert len(out["reference"]) >= 180
    diff = {k: (v, out["dropin"].get(k)) for k, v in out["reference"].items() if out["dropin"].get(k) != v}
    assert not diff, diff
