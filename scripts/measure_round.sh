# One-GPU measurement pass behind profiles/ (run under gpurun): bench lines for every BASELINE config, the
# reference arm, ncu launch lists and full captures of the hot kernels.  usage: scripts/measure_round.sh [tag]
TAG=${1:-r02b}
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
python bench.py --steps 50 --warmup 3 > gpurun_out/${TAG}_bench_c2.log 2>&1; tail -1 gpurun_out/${TAG}_bench_c2.log > gpurun_out/${TAG}_bench_c2.json
for W in C4 C3 C1 C5; do
  python bench.py --workload $W --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_$W.log 2>&1; tail -1 gpurun_out/${TAG}_bench_$W.log > gpurun_out/${TAG}_bench_$W.json
done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_c2_ref.log 2>&1; tail -1 gpurun_out/${TAG}_bench_c2_ref.log > gpurun_out/${TAG}_bench_c2_ref.json
# launch lists (cold-cache, serialised: shares matter) and full captures; numbers printed under ncu are never bench values
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_c2_f32.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-f64 > gpurun_out/ncu_launches.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches_c2_f64.csv python bench.py --workload C2f64 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches64.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"conv_halo2_kernel|wgrad_rows_kernel|pool_bwd" --launch-skip 10 -c 6 -f -o gpurun_out/${TAG}_c2_f32_full python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-f64 > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"f64_fwdlike_tma_kernel|f64_wgrad_tma_kernel|pool_max" --launch-skip 6 -c 5 -f -o gpurun_out/${TAG}_c2_f64_full python bench.py --workload C2f64 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full64.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_c5.csv python bench.py --workload C5 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_c5.log 2>&1
ls -la gpurun_out | tail -20
