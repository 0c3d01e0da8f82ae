"""The `-m gpu` test files themselves, run in the CPU suite with the emulated device under them (FZ_EMULATED_DEVICE=1 makes tests/conftest.py install
tests/emu_device.py: the library compiled for the CUDA-on-CPU emulator behind the product's Python layers).  The tests sized for a GPU (2^20
polynomials, capacity aggregates), the two that start a fresh interpreter (which, rightly, refuses to run without a CUDA device) and the few
that take tens of seconds on the emulator are left to the GPU box; the rest -- the drop-in API, the transforms and the fused pass, the scheme's stages against
goldens / KATs / the oracle, batched keygen + sign, the general rings, guard-word and overrun checks -- must pass here as they do there."""
import os
import subprocess
import sys

import hostlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = ["tests/test_gpu_dropin.py", "tests/test_gpu_scheme.py", "tests/test_gpu_ring.py", "tests/test_gpu_ntt.py", "tests/test_gpu_batch.py"]
LEFT_TO_THE_GPU = [
    "tests/test_gpu_dropin.py::test_top_level_dropin_layout",                 # fresh interpreter: no emulation there, the product demands a GPU
    "tests/test_gpu_dropin.py::test_differential_battery_on_the_device",     # the same (covered by tests/test_reference_suite_cpu.py, backend "emulated")
    "tests/test_gpu_scheme.py::test_buffered_and_streaming_hash_agree",      # 18 s + 60 s on the emulator
    "tests/test_gpu_scheme.py::test_aggregate_many_vs_oracle[256]",          # 26 s
    "tests/test_gpu_scheme.py::test_hash_pipeline_stays_inside_its_byte_buffers[3-256]",   # 10 s each; the secpar-128 cases of the same test run here
    "tests/test_gpu_scheme.py::test_hash_pipeline_stays_inside_its_byte_buffers[4-256]",
    "tests/test_gpu_ntt.py::test_ntt_full_size_properties",                  # 2^20 polynomials
    "tests/test_gpu_ntt.py::test_fused_pass_full_size_vs_oracle",            # 2^20 polynomials
    "tests/test_gpu_ntt.py::test_ntt_host_path",                             # pinned host memory needs the real driver
    "tests/test_gpu_batch.py::test_aggregate_verify_batch",                  # hundreds of signers per aggregate and capacity-sized batches: minutes on the emulator
    "tests/test_gpu_batch.py::test_pair_layout_many_ragged_aggregates",
    "tests/test_gpu_batch.py::test_chain_layout_is_chosen_by_the_number_of_aggregates",
    "tests/test_gpu_batch.py::test_single_aggregate_sharded_world1",
    "tests/test_gpu_batch.py::test_capacity_aggregate_properties",
    "tests/test_gpu_batch.py::test_norm_bound_rejection_in_a_batch[256]",    # 9 s; the secpar-128 case runs here
]


def test_gpu_test_files_pass_on_the_emulated_device():
    env = dict(os.environ, FZ_EMULATED_DEVICE="1", FZ_EMU_LIBRARY=hostlib.load_emulated_library()._name, PYTHONDONTWRITEBYTECODE="1")
    cmd = [sys.executable, "-m", "pytest", "-m", "gpu", "-q", "-p", "no:cacheprovider"] + FILES + ["--deselect=" + t for t in LEFT_TO_THE_GPU]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    tail = (r.stdout + r.stderr)[-4000:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "failed" not in r.stdout, tail
    n_passed = int(r.stdout.strip().splitlines()[-1].split(" passed")[0].split()[-1])
    assert n_passed >= 95, tail
