set -x
mkdir -p gpurun_out
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r1_head.json 2> gpurun_out/bench_r1_head.err; echo "bench exit $?"
python tools/show_bench.py gpurun_out/bench_r1_head.json
