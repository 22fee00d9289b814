set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r2d_tests.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r2d_bench.err
QI_BENCH_DTYPE=float64 QI_BENCH_EXTRAS=0 timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2d_bench_f64.json 2> gpurun_out/r2d_bench_f64.err; echo "bench64 rc=$?"
tail -3 gpurun_out/r2d_bench_f64.err
QI_BENCH_CHECKS=0 QI_BENCH_EXTRAS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2d_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2d_ncu.log 2>&1; echo "ncu rc=$?"
QI_BENCH_DTYPE=float64 QI_BENCH_CHECKS=0 QI_BENCH_EXTRAS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2d_launches_f64.csv python bench.py --steps 1 --warmup 1 > gpurun_out/r2d_ncu_f64.log 2>&1; echo "ncu64 rc=$?"
