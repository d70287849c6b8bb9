#!/bin/bash
# What the driver runs at round end (whole GPU suite in one process, smoke, both bench arms) plus the secondary bench lines
# and microbenchmarks the profiles quote:  gpurun -- bash scripts/gpu_verify.sh
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.txt gpurun_out/parity_fullwidth.txt gpurun_out/parity_fp32_mode.txt
( time timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider ) > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest -m gpu exit $?"; tail -4 gpurun_out/pytest_gpu_all.log
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"; cut -c1-160 gpurun_out/bench_default.json
python bench.py --impl reference > gpurun_out/bench_default_ref.json 2> gpurun_out/bench_default_ref.err; echo "ref exit $?"; cut -c1-160 gpurun_out/bench_default_ref.json
python bench.py --precision fp32 --steps 3 --no-cpu --no-eager > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; echo "fp32 mode exit $?"; cut -c1-200 gpurun_out/bench_fp32.json
python bench.py --workload config5 --steps 3 > gpurun_out/bench_config5.json 2> gpurun_out/bench_config5.err; echo "config5 exit $?"; cut -c1-200 gpurun_out/bench_config5.json
python bench.py --workload config1 --steps 8 --no-eager > gpurun_out/bench_config1.json 2> gpurun_out/bench_config1.err; echo "config1 exit $?"; cut -c1-200 gpurun_out/bench_config1.json
python scripts/gen_bench.py --frames 1000 > gpurun_out/gen_bench_final.log 2>&1; python scripts/gen_bench.py --frames 1000 --batch 64 >> gpurun_out/gen_bench_final.log 2>&1; cat gpurun_out/gen_bench_final.log | cut -c1-220
python scripts/gru_microbench.py --reps 4 --flags 0,16 --ts-flags 0 > gpurun_out/gru_mb_final.txt 2>&1; grep "flags=" gpurun_out/gru_mb_final.txt
