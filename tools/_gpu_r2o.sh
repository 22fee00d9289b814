set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2o_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r2o_tests.log
QI_BENCH_DTYPE=float64 QI_BENCH_EXTRAS=0 timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/r2o_bench_f64.json 2> gpurun_out/r2o_bench_f64.err; echo "bench64 rc=$?"
python tools/bench_configs.py cfg3 > gpurun_out/r2o_configs.jsonl 2> gpurun_out/r2o_configs.err; echo "configs rc=$?"
for hc in 2 4 8; do
QI_BENCH_HOST_CHUNKS=$hc QI_BENCH_EXTRAS=0 QI_BENCH_CHECKS=0 timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2o_bench_hc$hc.json 2> gpurun_out/r2o_bench_hc$hc.err; echo "bench hc$hc rc=$?"
done
QI_BENCH_DTYPE=float64 QI_BENCH_CHECKS=0 QI_BENCH_EXTRAS=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2o_launches_f64.csv python bench.py --steps 1 --warmup 1 > gpurun_out/r2o_ncu_f64.log 2>&1; echo "ncu64 rc=$?"
