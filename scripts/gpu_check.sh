#!/bin/bash
# Runs the GPU parity suites one pytest process per file (a CUDA fault in one kernel must not take
# the other suites down with it) and leaves the logs in gpurun_out/.
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for f in ${@:-tests/test_gpu_kernels.py tests/test_gpu_model.py}; do
  name=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu ${PYTEST_X:-} --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/$name.log 2>&1
  echo "$name exit $?" | tee -a gpurun_out/summary.txt
  tail -5 gpurun_out/$name.log
done
