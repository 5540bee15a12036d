#!/bin/bash
# ncu --set full of the statistics / predict pass of config 2 (summarised on the box: the report is 30 MB)
set -x
mkdir -p gpurun_out
P2="python tools/passbench.py --case c2_logistic --rows 8000000 --profile --reps 2"
$P2 > gpurun_out/plain_p2.log 2>&1 && ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:stats_pass -f -o gpurun_out/r2_ncu_stats_c2 $P2 > gpurun_out/ncu_p2.log 2>&1
tail -2 gpurun_out/plain_p2.log
python tools/ncu_summary.py gpurun_out/r2_ncu_stats_c2.ncu-rep > gpurun_out/r2_ncu_stats_pass_c2_full.txt 2>&1
rm -f gpurun_out/r2_ncu_stats_c2.ncu-rep
head -40 gpurun_out/r2_ncu_stats_pass_c2_full.txt
