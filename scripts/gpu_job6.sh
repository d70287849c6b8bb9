#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/gru_microbench.py --reps 4 --flags 0,16777216,16,4194304,20971520 --ts-flags 0 > gpurun_out/gru_mb_r2e.txt 2>&1; echo "gru_mb exit $?"
grep "flags=" gpurun_out/gru_mb_r2e.txt | head -12
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "gru or lstm" -p no:cacheprovider > gpurun_out/test_gru_r2e.log 2>&1; echo "gru tests exit $?"; tail -2 gpurun_out/test_gru_r2e.log
bash scripts/gpu_job_multi.sh 1
timeout 600 python bench.py --steps 8 --warmup 3 --no-eager --no-cpu > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; echo "bench exit $?"; cut -c1-200 gpurun_out/bench_r2d.json
