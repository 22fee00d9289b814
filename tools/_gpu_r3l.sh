set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r3l_bench.json 2> gpurun_out/r3l_bench.err; echo "bench rc=$?"
python tools/bench_configs.py cfg1 cfg4 cfg5 > gpurun_out/r3l_configs.jsonl 2> gpurun_out/r3l_configs.err; echo "configs rc=$?"
QI_BENCH_CHECKS=0 QI_BENCH_EXTRAS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r3l_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r3l_ncu.log 2>&1; echo "ncu rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:'mr_' --launch-skip 0 --launch-count 40 -f -o /tmp/r3l_multirate_full python tools/profile_step.py 1 > gpurun_out/r3l_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/r3l_multirate_full.ncu-rep --page raw --csv > gpurun_out/r3l_multirate_full_raw.csv 2>/dev/null; ls -la gpurun_out/r3l_multirate_full_raw.csv
