"""Generates tests/golden/matrices_host_golden.json by running the UNMODIFIED reference (algebra/matrices.py) on plain-int
matrices in the build container: PYTHONPATH is /root/reference.  The cases are listed in tests/test_matrices_host.py."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, "/root/reference")
from algebra.matrices import GeneralMatrix  # noqa: E402  (the reference)

from test_matrices_host import run_cases  # noqa: E402

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "matrices_host_golden.json")
json.dump(run_cases(GeneralMatrix), open(out, "w"), indent=0)
print("wrote", out)
