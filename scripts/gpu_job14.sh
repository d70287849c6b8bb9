#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "gru or lstm or recurrence" -p no:cacheprovider > gpurun_out/test_gru_multi.log 2>&1; echo "gru tests exit $?"; tail -6 gpurun_out/test_gru_multi.log
timeout 300 python scripts/gru_microbench.py --reps 3 --flags 0 --ts-flags '' --batch 128 > gpurun_out/gru_mb_b128.txt 2>&1; grep "flags=" gpurun_out/gru_mb_b128.txt
timeout 300 python scripts/gru_microbench.py --reps 3 --flags 0 --ts-flags '' --batch 64 > gpurun_out/gru_mb_b64.txt 2>&1; grep "flags=" gpurun_out/gru_mb_b64.txt
timeout 300 python scripts/gru_microbench.py --reps 3 --flags 0 --ts-flags '' --batch 256 > gpurun_out/gru_mb_b256.txt 2>&1; grep "flags=" gpurun_out/gru_mb_b256.txt
