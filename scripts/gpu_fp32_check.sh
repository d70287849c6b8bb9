#!/bin/bash
# fp32-tolerance mode: parity tests, the worst per-tensor numbers, and the config-2 bench line in that mode
mkdir -p gpurun_out
rm -f gpurun_out/parity_fp32_mode.txt
python -m pytest tests/test_gpu_fp32_mode.py -q -m gpu 2>&1 | tail -3
head -3 gpurun_out/parity_fp32_mode.txt
grep "full-width config2 chunk .: loss" gpurun_out/parity_fp32_mode.txt
echo "worst full-width tensors:"; grep "full-width config2 chunk . grad" gpurun_out/parity_fp32_mode.txt | sort -t' ' -k8 -g | tail -4 | cut -c1-150
echo "worst golden tensors:"; grep "worst gradient" gpurun_out/parity_fp32_mode.txt | sort -t' ' -k9 -g | tail -2
timeout 600 python bench.py --precision fp32 --steps 3 --warmup 3 --no-cpu --no-eager > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/bench_fp32.json'));print(d['ms_per_step'],d['value']);[print(r['kernel'],round(r['ms_per_step'],1),round(r['us_per_timestep'],1)) for r in d['recurrence']]"
