#!/bin/bash
# multi-GPU bench lines: weak scaling of config 2 and the fixed-global-batch configs 3 / 4.   usage: gpu_job_multi.sh N
N=$1
mkdir -p gpurun_out
run() {  # name, extra args
  name=$1; shift
  if [ "$N" = "1" ]; then
    python bench.py --gpus 1 --steps 6 --warmup 3 --no-eager --no-cpu "$@" > gpurun_out/r2_${name}_n$N.json 2> gpurun_out/r2_${name}_n$N.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 6 --warmup 3 "$@" > gpurun_out/r2_${name}_n$N.json 2> gpurun_out/r2_${name}_n$N.err
  fi
  echo "$name n=$N exit $?"; cut -c1-180 gpurun_out/r2_${name}_n$N.json
}
[ "$N" != "1" ] && run weak
run config3_g128 --workload config3 --global-batch 128
run config4_g256 --workload config4 --global-batch 256
