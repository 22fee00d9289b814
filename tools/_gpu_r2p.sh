set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2p_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r2p_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2p_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2p_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench rc=$?"
QI_BENCH_DTYPE=float64 QI_BENCH_EXTRAS=0 timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2p_bench_f64.json 2> gpurun_out/r2p_bench_f64.err; echo "bench64 rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2p_reference_arm.json 2> gpurun_out/r2p_reference_arm.err; echo "ref rc=$?"
python tools/bench_configs.py cfg1 cfg2 cfg3 cfg4 cfg5 > gpurun_out/r2p_configs.jsonl 2> gpurun_out/r2p_configs.err; echo "configs rc=$?"
QI_BENCH_CHECKS=0 QI_BENCH_EXTRAS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2p_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2p_ncu.log 2>&1; echo "ncu rc=$?"
QI_BENCH_DTYPE=float64 QI_BENCH_CHECKS=0 QI_BENCH_EXTRAS=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2p_launches_f64.csv python bench.py --steps 1 --warmup 1 > gpurun_out/r2p_ncu_f64.log 2>&1; echo "ncu64 rc=$?"
timeout 300 python tools/profile_step.py 1 && \
timeout 900 ncu --set full --clock-control none -k regex:'mr_' --launch-skip 0 --launch-count 30 -f -o /tmp/r2p_multirate_full python tools/profile_step.py 1 > gpurun_out/r2p_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/r2p_multirate_full.ncu-rep --page raw --csv > gpurun_out/r2p_multirate_full_raw.csv 2>/dev/null; ls -la /tmp/r2p_multirate_full.ncu-rep gpurun_out/r2p_multirate_full_raw.csv
QI_TC_BLOCKS=16384 timeout 600 ncu --set full --clock-control none -k regex:'gemm|nvjet|cutlass|sm100|sm90|xmma' --launch-count 6 -f -o /tmp/r2p_tc_gemm python tools/tc_dft_experiment.py > gpurun_out/r2p_ncu_tc.log 2>&1; echo "ncu tc rc=$?"
ncu -i /tmp/r2p_tc_gemm.ncu-rep --page raw --csv > gpurun_out/r2p_tc_gemm_raw.csv 2>/dev/null
du -sh gpurun_out
