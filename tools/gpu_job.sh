set -x
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1_final_n1.json 2> gpurun_out/bench_r1_final_n1.err; echo "bench exit $?"
cat gpurun_out/bench_r1_final_n1.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_final_ref.json 2> gpurun_out/bench_r1_final_ref.err; echo "ref exit $?"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu28.log 2>&1
tail -n 3 gpurun_out/ncu28.log
