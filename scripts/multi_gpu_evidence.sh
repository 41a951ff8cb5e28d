#!/bin/bash
# Multi-GPU evidence on one node (run under `gpurun --gpus N`): data-parallel parity (tests/dp_check.py with
# the peer-memory collective and with NCCL only), then bench.py weak and strong scaling for the workloads
# BASELINE.json names.  Logs go to gpurun_out/ (copy the ones to keep into profiles/).
# usage: scripts/multi_gpu_evidence.sh N [steps]
N=${1:-2}; K=${2:-20}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
MGB_DP_PEER=1 run 29611 tests/dp_check.py > gpurun_out/dp_check_n${N}_peer.log 2>&1; echo "dp_check peer rc=$?"; tail -2 gpurun_out/dp_check_n${N}_peer.log
MGB_DP_PEER=0 run 29612 tests/dp_check.py > gpurun_out/dp_check_n${N}_nccl.log 2>&1; echo "dp_check nccl rc=$?"; tail -1 gpurun_out/dp_check_n${N}_nccl.log
run 29621 bench.py --gpus $N --steps $K --warmup 3 --workload C2 > gpurun_out/scale_C2_weak_n$N.json 2> gpurun_out/scale_C2_weak_n$N.err; echo "C2 weak rc=$?"
run 29622 bench.py --gpus $N --steps $K --warmup 3 --workload C2 --scaling strong --no-e2e > gpurun_out/scale_C2_strong_n$N.json 2> gpurun_out/scale_C2_strong_n$N.err; echo "C2 strong rc=$?"
run 29623 bench.py --gpus $N --steps $K --warmup 3 --workload C4 --scaling strong --no-e2e > gpurun_out/scale_C4_strong_n$N.json 2> gpurun_out/scale_C4_strong_n$N.err; echo "C4 strong rc=$?"
run 29624 bench.py --gpus $N --steps $K --warmup 3 --workload C5 > gpurun_out/scale_C5_weak_n$N.json 2> gpurun_out/scale_C5_weak_n$N.err; echo "C5 weak rc=$?"
run 29625 scripts/allreduce_bench.py > gpurun_out/allreduce_n$N.json 2> gpurun_out/allreduce_n$N.err; echo "allreduce bench rc=$?"; tail -1 gpurun_out/allreduce_n$N.json | cut -c1-600
for f in C2_weak C2_strong C4_strong C5_weak; do
  python - "$f" "$N" <<'PY'
import json, sys
f, n = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open(f"gpurun_out/scale_{f}_n{n}.json").read().strip().splitlines()[-1])
    print(f, "N=" + n, round(d["value"]), d["unit"], round(d["ms_per_step"], 3), "ms/step", "parity_ok", d.get("parity_ok"),
          "fused", d.get("fused", {}).get("ms_per_step"), "e2e", d.get("e2e") and round(d["e2e"]["value"]))
except Exception as e:
    print(f, "N=" + n, "FAILED", e)
    print(open(f"gpurun_out/scale_{f}_n{n}.err").read()[-1500:])
PY
done
