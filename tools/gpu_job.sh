set -x
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1_final_n1.json 2> gpurun_out/bench_r1_final_n1.err; echo "bench exit $?"
python tools/show_bench.py gpurun_out/bench_r1_final_n1.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
