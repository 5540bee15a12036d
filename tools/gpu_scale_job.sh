#!/bin/bash
# bench.py on N GPUs of one box (development tool): gpurun --gpus N -- 'PGB_N=N bash tools/gpu_scale_job.sh'
set -x
mkdir -p gpurun_out
N=${PGB_N:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench exit $?"
python tools/show_bench.py gpurun_out/r2_bench_n$N.json
