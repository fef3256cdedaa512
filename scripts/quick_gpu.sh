#!/bin/bash
# quick GPU check of a kernel change: parity subset + phase profile of the 1024-cell C2 launch
#   gpurun -- 'bash scripts/quick_gpu.sh <tag>'
TAG=${1:-x}
mkdir -p gpurun_out/r2
python -m pytest tests/test_gpu_parity.py tests/test_fluxeq.py -m gpu -x -q > gpurun_out/r2/pytest_$TAG.log 2>&1
tail -4 gpurun_out/r2/pytest_$TAG.log
CATINT_PHASES=1 python scripts/profile_case.py 1000000 1024 2>&1 | grep -v transport.info > gpurun_out/r2/phases_$TAG.txt
cat gpurun_out/r2/phases_$TAG.txt
