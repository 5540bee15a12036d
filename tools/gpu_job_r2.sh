#!/bin/bash
# Round-2 measurement job (one gpurun call on a 1-GPU box):  gpurun --timeout 1500 -- 'bash tools/gpu_job_r2.sh'
# GPU parity suite, the bench line of both arms, kernel timings, the ncu launch list of the bench command and the
# ncu --set full captures of the Gram kernels; everything goes to gpurun_out/ (summaries are copied to profiles/).
# gpurun copies at most 64 MiB back: the capture of the statistics pass is its own job (tools/gpu_job_r2_stats.sh).
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -n 3 gpurun_out/r2_pytest_gpu.log
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err; echo "ref exit $?"
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench exit $?"
for c in c2_logistic:8000000 c3_poisson:4000000 c4_linear:8000000 c5_expectile:8000000; do
  timeout 300 python tools/kbench.py --case ${c%%:*} --rows ${c##*:}
done > gpurun_out/r2_kbench.log 2>&1
for c in c4_linear:32000000 c2_logistic:16000000 c5_expectile:32000000 c3_poisson:4000000; do
  timeout 300 python tools/passbench.py --case ${c%%:*} --rows ${c##*:}
done > gpurun_out/r2_passbench.log 2>&1
# launch list of the bench command (cold-cache, serialised: shares, not absolutes)
B="python bench.py --steps 2 --warmup 3 --configs none --no-cpu-baseline --no-e2e"
$B > gpurun_out/plain_b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file gpurun_out/r2_launches_bench.csv $B > gpurun_out/ncu_b.log 2>&1
# the Gram kernels of config 2 (prepass + ring) and of config 3 (sorted + generic), full sets with source
K2="python tools/kbench.py --case c2_logistic --rows 8000000 --only-gram --reps 1 --profile"
$K2 > gpurun_out/plain_k2.log 2>&1 && ncu --set full --clock-control none \
  --profile-from-start off -k regex:"gram_cells_ring|stats_pass_kernel" -f -o gpurun_out/r2_ncu_c2_gram $K2 > gpurun_out/ncu_k2.log 2>&1
K3="python tools/kbench.py --case c3_poisson --rows 2000000 --only-gram --reps 1 --profile"
$K3 > gpurun_out/plain_k3.log 2>&1 && ncu --set full --clock-control none \
  --profile-from-start off -k regex:"gram_cells_sorted|gram_cells_generic|stats_pass_kernel" -f -o gpurun_out/r2_ncu_c3_gram $K3 > gpurun_out/ncu_k3.log 2>&1
# the reports embed the whole module (tens of MB each) and gpurun copies 64 MiB back: summarise here, keep the text
python tools/ncu_summary.py gpurun_out/r2_ncu_c2_gram.ncu-rep > gpurun_out/r2_ncu_c2_gram_full.txt 2>&1
python tools/traffic_json.py gpurun_out/r2_ncu_c2_gram.ncu-rep 8000000 > gpurun_out/r2_ncu_gram_cells_traffic.json
python tools/ncu_summary.py gpurun_out/r2_ncu_c3_gram.ncu-rep > gpurun_out/r2_ncu_c3_gram_full.txt 2>&1
rm -f gpurun_out/r2_ncu_c2_gram.ncu-rep gpurun_out/r2_ncu_c3_gram.ncu-rep
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
