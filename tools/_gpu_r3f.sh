set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r3f_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r3f_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3f_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r3f_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r3f_bench.json 2> gpurun_out/r3f_bench.err; echo "bench rc=$?"
QI_BENCH_DTYPE=float64 QI_BENCH_EXTRAS=0 timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r3f_bench_f64.json 2> gpurun_out/r3f_bench_f64.err; echo "bench64 rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r3f_reference_arm.json 2> gpurun_out/r3f_reference_arm.err; echo "ref rc=$?"
python tools/bench_configs.py cfg1 cfg2 cfg3 cfg4 cfg5 > gpurun_out/r3f_configs.jsonl 2> gpurun_out/r3f_configs.err; echo "configs rc=$?"
QI_BENCH_CHECKS=0 QI_BENCH_EXTRAS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r3f_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r3f_ncu.log 2>&1; echo "ncu rc=$?"
timeout 300 python tools/profile_step.py 1 && \
timeout 900 ncu --set full --clock-control none -k regex:'mr_' --launch-skip 0 --launch-count 40 -f -o /tmp/r3f_multirate_full python tools/profile_step.py 1 > gpurun_out/r3f_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/r3f_multirate_full.ncu-rep --page raw --csv > gpurun_out/r3f_multirate_full_raw.csv 2>/dev/null; ls -la /tmp/r3f_multirate_full.ncu-rep gpurun_out/r3f_multirate_full_raw.csv
./tools/bin/wprobe > gpurun_out/r3f_wprobe.txt 2>&1; ./tools/bin/wprobe2 > gpurun_out/r3f_wprobe2.txt 2>&1
du -sh gpurun_out
