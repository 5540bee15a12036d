set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest44.log 2>&1; echo "pytest exit $?"
tail -n 15 gpurun_out/pytest44.log
for c in c4_linear:32000000 c2_logistic:16000000 c5_expectile:32000000 c3_poisson:4000000; do
timeout 300 python tools/passbench.py --case ${c%%:*} --rows ${c##*:} >> gpurun_out/pb5.log 2>&1
done
cat gpurun_out/pb5.log
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:stats_pass -f -o gpurun_out/prof_stats_c2_v4 python tools/passbench.py --case c2_logistic --rows 8000000 --reps 1 --profile > gpurun_out/ncu26.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:stats_pass -f -o gpurun_out/prof_stats_c4_v4 python tools/passbench.py --case c4_linear --rows 16000000 --reps 1 --profile > gpurun_out/ncu27.log 2>&1
tail -n 2 gpurun_out/ncu26.log gpurun_out/ncu27.log
