#!/usr/bin/env python3
"""Installs the UNMODIFIED reference into baseline/_ref (git-ignored; it travels to the GPU box with the gpurun snapshot).

    python tools/install_reference.py [/root/reference]

1. `pip install --no-index --no-build-isolation --target baseline/_ref <copy of the checkout>`: the reference's two
   packages, `algebra` and `fusion`, exactly as its own setup.py ships them (the checkout is read-only, so pip builds from a
   scratch copy under /tmp).
2. The parts of the checkout setup.py leaves out -- its test files, KAT generator + committed KAT CSVs, benchmark harness and
   demo -- are copied verbatim to baseline/_ref/_checkout/, because the `-m gpu` tests run the reference's own tests and KAT
   generator against the CUDA drop-in, and bench.py / tools/api_benchmarks.py time the reference on the GPU box's host cores.

Nothing here is product code and nothing under baseline/ is committed.
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")


def install(src: str = "/root/reference", force: bool = False) -> str:
    if not os.path.isdir(src):
        raise SystemExit(f"reference checkout {src} not found")
    if os.path.isdir(os.path.join(DEST, "fusion")) and os.path.isdir(os.path.join(DEST, "_checkout", "tests")) and not force:
        return DEST
    shutil.rmtree(DEST, ignore_errors=True)
    os.makedirs(DEST, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        work = os.path.join(tmp, "reference")
        shutil.copytree(src, work)
        for d, _, files in os.walk(work):
            os.chmod(d, 0o755)
            for f in files:
                os.chmod(os.path.join(d, f), 0o644)
        cmd = [sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation", "--no-deps", "--find-links", "/opt/wheelhouse",
               "--target", DEST, work]
        subprocess.run(cmd, check=True)
        co = os.path.join(DEST, "_checkout")
        os.makedirs(co, exist_ok=True)
        for part in ("tests", "KATs", "benchmarks", "misc", "README.md", "LICENSE"):
            p = os.path.join(work, part)
            if os.path.isdir(p):
                shutil.copytree(p, os.path.join(co, part))
            elif os.path.exists(p):
                shutil.copy(p, os.path.join(co, part))
    return DEST


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    print(install(args[0] if args else "/root/reference", force="--force" in sys.argv))
