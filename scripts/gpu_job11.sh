#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_fullwidth.py -q -m gpu -x -k "generation" -p no:cacheprovider > gpurun_out/test_gen_r2.log 2>&1; echo "gen tests exit $?"; tail -5 gpurun_out/test_gen_r2.log
python scripts/gen_variance.py 2>&1 | tail -12
