set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r2g_tests.log
QI_BENCH_EXTRAS=0 timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r2g_bench.err
QI_BENCH_DTYPE=float64 QI_BENCH_EXTRAS=0 timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2g_bench_f64.json 2> gpurun_out/r2g_bench_f64.err; echo "bench64 rc=$?"
tail -3 gpurun_out/r2g_bench_f64.err
QI_BENCH_CHECKS=0 QI_BENCH_EXTRAS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2g_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2g_ncu.log 2>&1; echo "ncu rc=$?"
QI_BENCH_DTYPE=float64 QI_BENCH_CHECKS=0 QI_BENCH_EXTRAS=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2g_launches_f64.csv python bench.py --steps 1 --warmup 1 > gpurun_out/r2g_ncu_f64.log 2>&1; echo "ncu64 rc=$?"
python tools/bench_configs.py cfg2 cfg3 > gpurun_out/r2g_configs.jsonl 2> gpurun_out/r2g_configs.err; echo "configs rc=$?"
