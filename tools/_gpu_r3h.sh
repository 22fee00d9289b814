set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r3h_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r3h_tests.log
python tools/bench_configs.py cfg1 cfg3 > gpurun_out/r3h_configs.jsonl 2> gpurun_out/r3h_configs.err; echo "configs rc=$?"
QI_BENCH_METHOD=exact QI_BENCH_EXTRAS=0 timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r3h_bench_exact_f32.json 2> gpurun_out/r3h_bench_exact_f32.err; echo "bench exact rc=$?"
python tools/quick_step.py 10 final > gpurun_out/r3h_quick.json 2>&1; cat gpurun_out/r3h_quick.json
