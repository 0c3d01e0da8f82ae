#!/usr/bin/env python3
"""Outcomes of tests/refsuite/differential_cases.py under the UNMODIFIED reference -> tests/golden/differential_golden.json
(the -m gpu tests run the same battery on the CUDA path and compare).  Build container only (needs /root/reference)."""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("FUSION_REFERENCE", "/root/reference")
script = os.path.join(os.path.dirname(HERE), "refsuite", "differential_cases.py")
env = dict(os.environ, PYTHONPATH=REF, PYTHONDONTWRITEBYTECODE="1")
out = subprocess.run([sys.executable, script], env=env, capture_output=True, text=True, check=True).stdout.strip().splitlines()[-1]
json.dump(json.loads(out), open(os.path.join(HERE, "differential_golden.json"), "w"), indent=0, sort_keys=True)
print(len(json.loads(out)), "cases")
