set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 300 -k "more_terms" > gpurun_out/pytest53.log 2>&1; rc=$?; echo "pytest exit $rc"
tail -n 25 gpurun_out/pytest53.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest54.log 2>&1; echo "full pytest exit $?"
tail -n 6 gpurun_out/pytest54.log
