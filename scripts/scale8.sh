#!/bin/bash
# one 8-GPU box: C2 with the peer-memory collective (default) and with NCCL, C5 with the default
mkdir -p gpurun_out
run() { # name N extra-env workload
  local name=$1 N=$2 envs=$3 W=$4
  env $envs python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) \
    bench.py --gpus $N --steps 20 --warmup 3 --workload $W --no-cpu-baseline --no-e2e > gpurun_out/$name.log 2>&1
  tail -1 gpurun_out/$name.log > gpurun_out/$name.json
  python -c "
import json; d=json.load(open('gpurun_out/$name.json')); print('$name', round(d['value']), 'samples/s', round(d['ms_per_step'],3), 'ms/step')" || tail -5 gpurun_out/$name.log
}
NG=$(nvidia-smi -L | wc -l)
run scale_C2_n${NG}_peer $NG "MGB_DP_PEER=1" C2
run scale_C2_n${NG}_nccl $NG "MGB_DP_PEER=0" C2
run scale_C5_n${NG}_peer $NG "MGB_DP_PEER=1" C5
