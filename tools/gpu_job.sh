set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest55.log 2>&1; echo "full pytest exit $?"
tail -n 6 gpurun_out/pytest55.log
timeout 300 python tools/passbench.py --case c2_logistic --rows 16000001 > gpurun_out/pb9.log 2>&1
timeout 300 python tools/passbench.py --case c4_linear --rows 32000001 >> gpurun_out/pb9.log 2>&1
timeout 300 python tools/passbench.py --case c2_logistic --rows 16000000 >> gpurun_out/pb9.log 2>&1
cat gpurun_out/pb9.log
