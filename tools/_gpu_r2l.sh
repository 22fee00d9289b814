set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r2l_tests.log
python tools/bench_configs.py cfg3 > gpurun_out/r2l_configs.jsonl 2> gpurun_out/r2l_configs.err; echo "configs rc=$?"
cat gpurun_out/r2l_configs.jsonl
