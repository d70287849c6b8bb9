#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_fullwidth.py -q -m gpu -x -k "generation" -p no:cacheprovider > gpurun_out/test_gen_r2.log 2>&1; echo "gen tests exit $?"; tail -5 gpurun_out/test_gen_r2.log
grep generation gpurun_out/parity_report.txt gpurun_out/parity_fullwidth.txt | tail -8
timeout 600 python scripts/gen_bench.py --frames 1000 > gpurun_out/gen_bench_r2b.log 2>&1; echo "gen_bench exit $?"; tail -2 gpurun_out/gen_bench_r2b.log
timeout 600 python scripts/gen_bench.py --frames 1000 --batch 64 >> gpurun_out/gen_bench_r2b.log 2>&1; tail -1 gpurun_out/gen_bench_r2b.log
