python scripts/gru_microbench.py --flags 0 --ts-flags "" --steps 4000 > gpurun_out/gru_mb4000.log 2>&1
python scripts/gru_microbench.py --flags 0 --ts-flags "" --steps 1000 --reps 1 >> gpurun_out/gru_mb4000.log 2>&1
