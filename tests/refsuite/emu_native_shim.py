"""`native` for subprocess runs of the reference's own test files / KAT generator / differential battery against the drop-in ON THE EMULATED DEVICE:
copied as native.py to the front of PYTHONPATH, it loads the drop-in's real native module under that name and points it at the library compiled
for the CUDA-on-CPU emulator (tests/emu_device.py).  Test infrastructure only."""
import importlib.util
import os
import sys

_repo = os.environ["FUSION_B200_REPO"]
sys.path.insert(0, os.path.join(_repo, "tests"))
sys.path.insert(0, _repo)
_spec = importlib.util.spec_from_file_location("native", os.path.join(_repo, "fusion_cryptography_b200", "native.py"))
_mod = importlib.util.module_from_spec(_spec)
sys.modules["native"] = _mod
_spec.loader.exec_module(_mod)
import emu_device  # noqa: E402

emu_device.install(emu_device.Permanent, native=_mod)
