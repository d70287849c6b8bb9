#!/bin/bash
# round-2 GPU call: final-build verification: suites, step breakdown, bench (with eager + cpu legs), ncu launch list (recurrent kernels excluded)
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.txt gpurun_out/parity_fullwidth.txt gpurun_out/summary.txt
bash scripts/gpu_check.sh tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_fullwidth.py
timeout 600 python scripts/profile_step.py > gpurun_out/step_breakdown_r2.txt 2>&1; echo "breakdown exit $?"; head -12 gpurun_out/step_breakdown_r2.txt
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err; echo "bench exit $?"; cut -c1-200 gpurun_out/bench_r2e.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r2.json 2> gpurun_out/bench_ref_r2.err; echo "ref exit $?"; cut -c1-200 gpurun_out/bench_ref_r2.json
python bench.py --steps 2 --warmup 3 --no-eager --no-cpu > gpurun_out/plain_for_ncu.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(?!.*gru_kernel)' -c 2000 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 2 --warmup 3 --no-eager --no-cpu > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
