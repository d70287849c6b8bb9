#!/bin/bash
# round-2 GPU call 1: parity suites (incl. full width), recurrence microbenchmark, bench line, smoke
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.txt gpurun_out/parity_fullwidth.txt gpurun_out/summary.txt
python scripts/nan_probe.py > gpurun_out/nan_probe.log 2>&1; echo "nan_probe exit $?" | tee -a gpurun_out/summary.txt
bash scripts/gpu_check.sh tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_fullwidth.py
timeout 600 python scripts/gru_microbench.py --flags 0,32,16 --ts-flags 0,32 > gpurun_out/gru_mb_r2.txt 2>&1; echo "gru_mb exit $?" | tee -a gpurun_out/summary.txt
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "bench exit $?" | tee -a gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/gru_mb_r2.txt; cut -c1-600 gpurun_out/bench_r2a.json
