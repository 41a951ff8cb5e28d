set -x
mkdir -p gpurun_out
python bench.py --steps 50 --warmup 3 > gpurun_out/bench_c2.log 2>&1; tail -1 gpurun_out/bench_c2.log > gpurun_out/bench_c2.json
for W in C2f64 C4 C3 C1 C5; do
  python bench.py --workload $W --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$W.log 2>&1; tail -1 gpurun_out/bench_$W.log > gpurun_out/bench_$W.json
done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_c2_ref.log 2>&1; tail -1 gpurun_out/bench_c2_ref.log > gpurun_out/bench_c2_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2_f32.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"conv_halo_kernel|conv_tma_kernel|wgrad_tma_kernel|nchw_to_nhwc_split|pool_max" --launch-skip 12 -c 8 -o gpurun_out/c2_f32_full python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
