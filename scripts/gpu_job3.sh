#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/gru_microbench.py --reps 4 --flags 0,2097152,4194304,6291456,8388608,10485760,12582912 --ts-flags 6291456 > gpurun_out/gru_mb_r2d.txt 2>&1; echo "gru_mb exit $?"
grep "flags=" gpurun_out/gru_mb_r2d.txt | head -24
timeout 900 python bench.py --steps 8 --warmup 3 --no-eager > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; echo "bench exit $?"
cut -c1-200 gpurun_out/bench_r2c.json
