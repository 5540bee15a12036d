#!/bin/bash
# Final round-2 measurement job (one gpurun call, 1 GPU):  gpurun --timeout 1500 -- 'bash tools/gpu_job_r2_final.sh'
# GPU parity suite, both bench arms, host-upload timings, the ncu launch list of the bench command.  The ncu --set full
# captures of the Gram / statistics kernels are in tools/gpu_job_r2.sh and tools/gpu_job_r2_stats.sh (kernels unchanged since).
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -n 3 gpurun_out/r2_pytest_gpu.log
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err; echo "ref exit $?"
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench exit $?"
timeout 200 python tools/uploadbench.py > gpurun_out/r2_uploadbench.log 2>&1
B="python bench.py --steps 2 --warmup 3 --configs none --no-cpu-baseline --no-e2e"
$B > gpurun_out/plain_b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file gpurun_out/r2_launches_bench.csv $B > gpurun_out/ncu_b.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python tools/show_bench.py gpurun_out/r2_bench_n1.json | head -40
