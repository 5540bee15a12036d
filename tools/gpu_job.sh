#!/bin/bash
# What one gpurun call of the development loop runs on the B200 box (development tool):
#   gpurun --timeout 1200 -- 'bash tools/gpu_job.sh'
# GPU parity suite, the bench line with its CPU baseline and end-to-end leg, the O(k) passes against the HBM peak,
# smoke(); everything it writes goes to gpurun_out/ (merged back by gpurun).
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_last.log 2>&1; echo "pytest exit $?"
tail -n 6 gpurun_out/pytest_last.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_last_n1.json 2> gpurun_out/bench_last_n1.err; echo "bench exit $?"
python tools/show_bench.py gpurun_out/bench_last_n1.json
for c in c4_linear:32000000 c2_logistic:16000000 c5_expectile:32000000 c3_poisson:4000000; do
  timeout 300 python tools/passbench.py --case ${c%%:*} --rows ${c##*:}
done 2>&1 | tee gpurun_out/passbench_last.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
