set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "multirate or mr or entropy or cwt or shard or config or smoke" > gpurun_out/r2n_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r2n_tests.log
QI_BENCH_EXTRAS=0 timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r2n_bench.err
QI_BENCH_CHECKS=0 QI_BENCH_EXTRAS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2n_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2n_ncu.log 2>&1; echo "ncu rc=$?"
