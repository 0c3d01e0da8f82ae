#!/bin/bash
# One gpurun call: GPU tests, smoke, the bench line, the ncu launch list of the short bench and a --set full capture of the fused NTT pass kernel.
TAG=${1:-r01g}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${TAG}_smoke.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python tools/ring_times.py > gpurun_out/${TAG}_ring_times.log 2>&1
SHORT="python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline"
$SHORT > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $SHORT > gpurun_out/${TAG}_ncu1.log 2>&1
$SHORT > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ntt_pass_kernel -s 3 -c 2 -f -o gpurun_out/prof_ntt_pass_${TAG} $SHORT > gpurun_out/${TAG}_ncu2.log 2>&1
for f in pytest smoke ring_times; do tail -n 3 gpurun_out/${TAG}_$f.log; done
