#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.txt gpurun_out/parity_fullwidth.txt gpurun_out/summary.txt
bash scripts/gpu_check.sh tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_fullwidth.py
SRNN_GEMM_STORE64=1 timeout 600 python scripts/profile_step.py > gpurun_out/step_breakdown_store64.txt 2>&1; echo "breakdown64 exit $?"
timeout 600 python scripts/profile_step.py > gpurun_out/step_breakdown_store128.txt 2>&1; echo "breakdown128 exit $?"
for f in store64 store128; do echo $f; grep -E "^step|NT m=1024000x1 n=1024 k=1024|NT m=1024000x1 n=1024 k=256|NT m=16000x64|NT m=256000x1 n=4096|NT m=256000x1 n=1024 k=56|NT m=256000x1 n=3072" gpurun_out/step_breakdown_$f.txt; done
SRNN_GEMM_STORE64=1 timeout 600 python bench.py --steps 8 --warmup 3 --no-eager --no-cpu > gpurun_out/bench_store64.json 2>/dev/null; cut -c1-200 gpurun_out/bench_store64.json
timeout 600 python bench.py --steps 8 --warmup 3 --no-eager --no-cpu > gpurun_out/bench_store128.json 2>/dev/null; cut -c1-200 gpurun_out/bench_store128.json
