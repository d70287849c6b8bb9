PYTEST_X=-x bash scripts/gpu_check.sh
python bench.py --steps 16 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
