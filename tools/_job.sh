set -x
export PATH=/usr/local/cuda/bin:$PATH
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "stats_pass or gram_pass_matches_oracle or host_buffer" --timeout 800 > gpurun_out/memcheck.log 2>&1; echo "memcheck exit $?"
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/memcheck.log; tail -12 gpurun_out/memcheck.log
