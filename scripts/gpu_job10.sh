#!/bin/bash
# exactly what the driver runs at round end: the whole GPU suite in one process, smoke(), the default bench (both arms)
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.txt gpurun_out/parity_fullwidth.txt
( time timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider ) > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest -m gpu exit $?"; tail -4 gpurun_out/pytest_gpu_all.log
( time python -c "import __graft_entry__ as g; g.build(); g.smoke()" ) > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/smoke.log
( time python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"; cut -c1-160 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
( time python bench.py --impl reference ) > gpurun_out/bench_default_ref.json 2> gpurun_out/bench_default_ref.err; echo "ref exit $?"; cut -c1-160 gpurun_out/bench_default_ref.json; tail -3 gpurun_out/bench_default_ref.err
python bench.py --workload config5 --steps 2 > gpurun_out/bench_config5.json 2> gpurun_out/bench_config5.err; echo "config5 exit $?"; cut -c1-200 gpurun_out/bench_config5.json
