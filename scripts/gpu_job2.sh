#!/bin/bash
# round-2 GPU call 2: fence-free exchange - kernel tests, stress, microbenchmark, model parity, bench
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.txt gpurun_out/parity_fullwidth.txt gpurun_out/summary.txt
timeout 600 python scripts/gru_microbench.py --flags 0,16,32 --ts-flags 0,16 > gpurun_out/gru_mb_r2b.txt 2>&1; echo "gru_mb exit $?" | tee -a gpurun_out/summary.txt
head -8 gpurun_out/gru_mb_r2b.txt
bash scripts/gpu_check.sh tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_fullwidth.py
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; echo "bench exit $?" | tee -a gpurun_out/summary.txt
cut -c1-300 gpurun_out/bench_r2b.json
