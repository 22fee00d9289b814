set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r3a_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r3a_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3a_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r3a_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r3a_bench.json 2> gpurun_out/r3a_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r3a_bench.err
