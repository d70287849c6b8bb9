python bench.py --steps 16 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
python bench.py --steps 8 --warmup 3 --workload config3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
python bench.py --steps 16 --warmup 3 --workload config4 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
