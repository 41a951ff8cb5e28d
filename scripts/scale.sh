#!/bin/bash
# weak-scaling sweep on one node: bench.py at N = 1, 2, 4, 8 (whatever fits the visible GPUs)
# usage: scripts/scale.sh [workload] [steps]
W=${1:-C2}; K=${2:-20}
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
for N in 1 2 4 8; do
  [ "$N" -gt "$NG" ] && break
  if [ "$N" -eq 1 ]; then
    python bench.py --gpus 1 --steps $K --warmup 3 --workload $W --no-cpu-baseline > gpurun_out/scale_${W}_n$N.log 2>&1
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
      bench.py --gpus $N --steps $K --warmup 3 --workload $W --no-cpu-baseline > gpurun_out/scale_${W}_n$N.log 2>&1
  fi
  tail -1 gpurun_out/scale_${W}_n$N.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('N=$N', '$W', round(d['value']), 'samples/s', round(d['ms_per_step'],3), 'ms/step', 'e2e', round(d['e2e']['value']) if d.get('e2e') else None)
except Exception as e: print('N=$N failed', e)
"
done
