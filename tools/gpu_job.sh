set -x
mkdir -p gpurun_out
timeout 300 python tools/passbench.py --case c2_logistic --rows 16000001 > gpurun_out/pb8.log 2>&1
timeout 300 python tools/passbench.py --case c4_linear --rows 32000001 >> gpurun_out/pb8.log 2>&1
cat gpurun_out/pb8.log
