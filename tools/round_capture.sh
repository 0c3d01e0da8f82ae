#!/bin/bash
# One gpurun call: the ncu evidence of a round.  Launch lists (--metrics gpu__time_duration.sum) of the short bench and of the scheme probe, and
# --set full captures of the dominant kernel of every leg.  Each ncu command runs only after the same command exited 0 without ncu.
TAG=${1:-r02}
mkdir -p gpurun_out
NCU="ncu --clock-control none"
SHORT="python bench.py --steps 3 --warmup 3 --no-legs --no-cpu-baseline"
SCHEME="python tools/profile_scheme.py"
$SHORT > gpurun_out/${TAG}_plain_bench.log 2>&1 &&
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench.csv $SHORT > gpurun_out/${TAG}_ncu_l1.log 2>&1
$SHORT > /dev/null 2>&1 &&
$NCU --set full --import-source on -k regex:ntt_pass_kernel -s 3 -c 2 -f -o gpurun_out/prof_${TAG}_ntt_pass $SHORT > gpurun_out/${TAG}_ncu_f1.log 2>&1
$SCHEME > gpurun_out/${TAG}_plain_scheme.log 2>&1 &&
$NCU --metrics gpu__time_duration.sum -c 4000 --csv --log-file gpurun_out/${TAG}_launches_scheme.csv $SCHEME > gpurun_out/${TAG}_ncu_l2.log 2>&1
$SCHEME > /dev/null 2>&1 &&
$NCU --set full --import-source on -k regex:'sample_dense_kernel|hash_ch_stream_kernel|sign_rows_kernel|aggregate_kernel|agw_absorb|agw_squeeze|agw_format|agw_decode|verify_' -c 48 -f -o gpurun_out/prof_${TAG}_scheme $SCHEME > gpurun_out/${TAG}_ncu_f2.log 2>&1
for f in gpurun_out/prof_${TAG}_*.ncu-rep; do
  ncu -i $f --page raw --csv > ${f%.ncu-rep}_raw.csv 2> /dev/null
  # gpurun merges at most 64 MiB back: keep the raw page, drop reports above 8 MiB
  [ $(stat -c %s $f) -gt 8388608 ] && rm -f $f
done
ls -la gpurun_out | tail -n 20
