set -x
mkdir -p gpurun_out
QI_BENCH_METHOD=exact QI_BENCH_EXTRAS=0 timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2r_bench_exact_f32.json 2> gpurun_out/r2r_bench_exact_f32.err; echo "bench exact rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2r_launches_cfg3.csv python tools/bench_configs.py cfg3 > gpurun_out/r2r_ncu_cfg3.log 2>&1; echo "ncu rc=$?"
