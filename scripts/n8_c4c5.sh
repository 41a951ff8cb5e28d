# 8-GPU re-run of the two workloads the folded wgrad changed (run under gpurun --gpus 8): config 4 strong scaling, config 5 weak
N=8
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29623 bench.py --gpus $N --steps 10 --warmup 3 --workload C4 --scaling strong --no-e2e > gpurun_out/scale_C4_strong_n$N.json 2> gpurun_out/scale_C4_strong_n$N.err; echo "C4 strong rc=$?"
run 29624 bench.py --gpus $N --steps 10 --warmup 3 --workload C5 > gpurun_out/scale_C5_weak_n$N.json 2> gpurun_out/scale_C5_weak_n$N.err; echo "C5 weak rc=$?"
for f in C4_strong C5_weak; do python -c "
import json; d=json.loads(open('gpurun_out/scale_${f}_n8.json').read().strip().splitlines()[-1]); print('$f', round(d['value']), round(d['ms_per_step'],3), d.get('parity_ok'))"; done
