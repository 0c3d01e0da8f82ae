#!/bin/bash
# One gpurun call: GPU tests, the bench line, the ncu launch list of the short bench and a --set full capture of the fused NTT pass kernel.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r01f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r01f_pytest.log
python bench.py > gpurun_out/r01f_bench.json 2> gpurun_out/r01f_bench.err; echo "bench rc=$?"
SHORT="python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline"
$SHORT > gpurun_out/r01f_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01f_launches.csv $SHORT > gpurun_out/r01f_ncu1.log 2>&1
$SHORT > gpurun_out/r01f_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ntt_pass_kernel -s 3 -c 2 -f -o gpurun_out/prof_ntt_pass $SHORT > gpurun_out/r01f_ncu2.log 2>&1
tail -3 gpurun_out/r01f_pytest.log
