set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest47.log 2>&1; echo "pytest exit $?"
tail -n 8 gpurun_out/pytest47.log
timeout 300 python tools/passbench.py --case c3_poisson --rows 4000000 > gpurun_out/pb7.log 2>&1
timeout 300 python tools/passbench.py --case c2_logistic --rows 16000000 >> gpurun_out/pb7.log 2>&1
cat gpurun_out/pb7.log
