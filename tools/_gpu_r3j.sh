set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r3j_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r3j_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3j_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r3j_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r3j_bench.json 2> gpurun_out/r3j_bench.err; echo "bench rc=$?"
python tools/bench_configs.py cfg1 cfg2 cfg3 cfg4 cfg5 > gpurun_out/r3j_configs.jsonl 2> gpurun_out/r3j_configs.err; echo "configs rc=$?"
QI_BENCH_METHOD=exact QI_BENCH_EXTRAS=0 timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r3j_bench_exact_f32.json 2> gpurun_out/r3j_bench_exact_f32.err; echo "bench exact rc=$?"
