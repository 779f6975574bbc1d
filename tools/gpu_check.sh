#!/bin/bash
# Run every GPU test file in its own process (a kernel trap poisons the CUDA context of its process only), with timeouts.
# Usage (on the GPU box): bash tools/gpu_check.sh [file ...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu_info.txt 2>&1
files="$@"
[ -z "$files" ] && files=$(ls tests/test_gpu_*.py tests/test_boundary_logger.py)
rc_all=0
for f in $files; do
  name=$(basename $f .py)
  echo "=== $f" | tee -a gpurun_out/gpu_check.log
  timeout 600 python -m pytest $f -m gpu -q -x --timeout 300 -p no:cacheprovider > gpurun_out/$name.log 2>&1
  rc=$?
  tail -n 25 gpurun_out/$name.log | tee -a gpurun_out/gpu_check.log
  echo "rc=$rc" | tee -a gpurun_out/gpu_check.log
  [ $rc -ne 0 ] && rc_all=1
done
exit $rc_all
