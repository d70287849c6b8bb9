#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.txt gpurun_out/parity_fullwidth.txt gpurun_out/summary.txt
bash scripts/gpu_check.sh tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_fullwidth.py
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/dp_equivalence.py > gpurun_out/dp_equivalence_n2.txt 2>&1; echo "dp eq exit $?"; grep -E "chunk|ok|Error|error" gpurun_out/dp_equivalence_n2.txt | tail -6
