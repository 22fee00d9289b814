set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r2c_bench.err
QI_BENCH_CHECKS=0 QI_BENCH_EXTRAS=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2c_ncu.log 2>&1; echo "ncu rc=$?"
