#!/bin/bash
# round-2 GPU call: ncu launch list of the bench command + full captures of the recurrent kernel (benched size) and comb GEMM
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-eager > gpurun_out/plain_for_ncu.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 2 --warmup 3 --no-eager > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
python scripts/gru_microbench.py --steps 4000 --flags 0 --reps 1 --ts-flags '' > gpurun_out/gru_plain4000.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gru_kernel -c 2 -o gpurun_out/gru_r2 \
    python scripts/gru_microbench.py --steps 4000 --flags 0 --reps 1 --ts-flags '' > gpurun_out/ncu_gru.log 2>&1
echo "gru capture exit $?"
ROWS=1024000 python scripts/ncu_targets.py > gpurun_out/targets_plain_r2.txt 2>&1 &&
ROWS=1024000 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -c 2 -o gpurun_out/comb_r2 \
    python scripts/ncu_targets.py > gpurun_out/ncu_comb.log 2>&1
echo "comb capture exit $?"
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/ncu_gru.log
