#!/bin/bash
# round-2 GPU call: generation changes (multi-group recurrent step, device Philox) + full suites + config5 line
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.txt gpurun_out/parity_fullwidth.txt gpurun_out/summary.txt
bash scripts/gpu_check.sh tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_fullwidth.py
timeout 600 python scripts/gen_bench.py --frames 1000 > gpurun_out/gen_bench_r2.log 2>&1; echo "gen_bench exit $?"; tail -2 gpurun_out/gen_bench_r2.log
timeout 600 python scripts/gen_bench.py --frames 1000 --batch 64 >> gpurun_out/gen_bench_r2.log 2>&1
timeout 900 python bench.py --workload config5 --steps 2 --warmup 3 > gpurun_out/bench_config5.json 2> gpurun_out/bench_config5.err; echo "config5 exit $?"; cut -c1-400 gpurun_out/bench_config5.json
