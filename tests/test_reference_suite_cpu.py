"""The reference's OWN unit tests -- all four files, 53 tests: tests/test_ntt.py, test_polynomials.py, test_matrices.py (moduli below 2^17, degrees
4..64 and one 1024) and test_fusion.py (setup, keygen, the hash helpers incl. the pytest-mock based ones, decoder, sign, aggregate, verify, tampering)
-- run UNMODIFIED against the host mirror: classes, operators, validation, exceptions, ``str()`` formats, sorting, the reason strings and the
module attributes the mocks patch are ours; the calls the mirror makes to the device are served by the oracle stand-in of tests/refsuite, whose
functions are the ones the CUDA kernels are compared with bit for bit (tests/test_gpu_ring.py, test_gpu_scheme.py, test_gpu_dropin.py).

Every test runs with two backends behind the mirror: "standin" -- the oracle stand-in described above, which checks the host mirror alone -- and
"emulated" -- the drop-in's REAL native module and the library compiled for the CUDA-on-CPU emulator (tests/host/cuda_emu.h, tests/emu_device.py),
i.e. mirror, ctypes binding, C ABI, launch code and kernels on the CPU.

Runs only where the reference checkout exists (the build container); the test files are read from there into a pytest scratch
directory -- they are never copied into this repository."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("FUSION_REFERENCE", "/root/reference")
FILES = ["test_ntt.py", "test_polynomials.py", "test_matrices.py", "test_fusion.py"]
BACKENDS = ["standin", "emulated"]


def make_mirror(tmp_path, backend):
    """A directory to put on PYTHONPATH: `algebra` and `fusion` are the drop-in's packages, `native` is the chosen backend."""
    mirror = tmp_path / ("mirror_" + backend)
    mirror.mkdir()
    os.symlink(os.path.join(ROOT, "fusion_cryptography_b200", "algebra"), mirror / "algebra")
    os.symlink(os.path.join(ROOT, "fusion_cryptography_b200", "fusion"), mirror / "fusion")
    src = "device_standin.py" if backend == "standin" else "emu_native_shim.py"
    shutil.copy(os.path.join(ROOT, "tests", "refsuite", src), mirror / "native.py")
    if backend == "emulated":
        import hostlib

        hostlib.load_emulated_library()      # built here (or the test is skipped here) rather than inside the subprocess
        os.environ["FZ_EMU_LIBRARY"] = hostlib.load_emulated_library()._name
    return mirror


def run_all(jobs, timeout=1500):
    """jobs: name -> (argv, cwd, env); the subprocesses run side by side.  Returns name -> CompletedProcess-like (returncode, stdout, stderr)."""
    procs = {name: subprocess.Popen(argv, cwd=cwd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for name, (argv, cwd, env) in jobs.items()}
    out = {}
    for name, p in procs.items():
        try:
            so, se = p.communicate(timeout=timeout)
        except subprocess.TimeoutExpired:
            p.kill()
            so, se = p.communicate()
            se += "\n[timed out]"
        out[name] = (p.returncode, so, se)
    return out


@pytest.mark.skipif(not all(os.path.exists(os.path.join(REF, "tests", f)) for f in FILES), reason="no reference checkout here")
def test_reference_unit_tests_pass_on_host_mirror(tmp_path):
    jobs = {}
    for backend in BACKENDS:
        mirror = make_mirror(tmp_path, backend)
        top = tmp_path / ("suite_" + backend)
        tdir = top / "tests"
        tdir.mkdir(parents=True)
        (tdir / "__init__.py").write_text("")   # test_matrices.py imports `tests.test_ntt`; a regular package beats any other `tests` on the path
        shutil.copy(os.path.join(ROOT, "tests", "refsuite", "mocker_shim.py"), tdir / "conftest.py")
        for f in FILES:
            shutil.copy(os.path.join(REF, "tests", f), tdir / f)
        env = dict(os.environ, PYTHONPATH=os.pathsep.join([str(mirror), str(top), str(tdir)]), FUSION_B200_REPO=ROOT, PYTHONDONTWRITEBYTECODE="1")
        jobs[backend] = ([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider"] + [str(tdir / f) for f in FILES], str(top), env)
    for backend, (rc, so, se) in run_all(jobs).items():
        tail = (so + se)[-3000:]
        assert rc == 0, (backend, tail)
        assert " passed" in so and "failed" not in so and "error" not in so.lower(), (backend, tail)
        assert "53 passed" in so, (backend, tail)


GEN = os.path.join(REF, "KATs", "generate_KAT_values.py")


@pytest.mark.skipif(not os.path.exists(GEN), reason="no reference checkout here")
def test_reference_kat_generator_writes_identical_files(tmp_path):
    """The reference's KAT generator (KATs/generate_KAT_values.py, executed unmodified from the reference checkout with the global generator seeded)
    run once against the reference and once against the host mirror with each backend (the oracle stand-in; the real stack on the emulated device):
    all 18 CSV files -- 64 MB of text, among them the 12 kinds the reference's repository is missing -- must be byte-identical.  This also pins the state every call leaves
    CPython's global generator in (the generator draws its seeds from it between calls)."""
    import hashlib

    runner = ("import random; random.seed(20231018)\n"
              f"exec(compile(open({GEN!r}).read(), 'generate_KAT_values', 'exec'))\n")
    jobs = {}
    for name, path in [("reference", REF)] + [(b, str(make_mirror(tmp_path, b))) for b in BACKENDS]:
        wd = tmp_path / ("run_" + name)
        wd.mkdir()
        jobs[name] = ([sys.executable, "-c", runner], str(wd), dict(os.environ, PYTHONPATH=path, FUSION_B200_REPO=ROOT, PYTHONDONTWRITEBYTECODE="1"))
    digests = {}
    for name, (rc, so, se) in run_all(jobs).items():
        assert rc == 0, (name, (so + se)[-3000:])
        out = tmp_path / ("run_" + name) / "KAT_values"
        digests[name] = {f: hashlib.sha256((out / f).read_bytes()).hexdigest() for f in sorted(os.listdir(out))}
    assert len(digests["reference"]) == 18
    for b in BACKENDS:
        assert digests[b] == digests["reference"], b


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "fusion", "fusion.py")), reason="no reference checkout here")
def test_differential_battery_against_the_reference(tmp_path):
    """tests/refsuite/differential_cases.py -- about 150 ordinary, edge and erroneous calls over algebra.ntt, algebra.polynomials, algebra.matrices
    and fusion.fusion (digests of str(result), or the exception's type and text) -- run with the names resolving to the unmodified reference and to
    the host mirror with each backend: every outcome must be identical.  It found (and now pins) the reference's exact NotImplementedError texts for mixed-ring
    operands, hash_ag's behaviour for unequal list lengths and for arbitrary key objects, and its handling of non-int seeds and non-str messages."""
    import json

    script = os.path.join(ROOT, "tests", "refsuite", "differential_cases.py")
    jobs = {name: ([sys.executable, script], str(tmp_path), dict(os.environ, PYTHONPATH=path, FUSION_B200_REPO=ROOT, PYTHONDONTWRITEBYTECODE="1"))
            for name, path in [("reference", REF)] + [(b, str(make_mirror(tmp_path, b))) for b in BACKENDS]}
    out = {}
    for name, (rc, so, se) in run_all(jobs).items():
        assert rc == 0, (name, (so + se)[-3000:])
        out[name] = json.loads(so.strip().splitlines()[-1])
    assert len(out["reference"]) >= 173
    for b in BACKENDS:
        diff = {k: (v, out[b].get(k)) for k, v in out["reference"].items() if out[b].get(k) != v}
        assert not diff, (b, diff)
