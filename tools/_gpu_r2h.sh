set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2h_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r2h_tests.log
QI_BENCH_EXTRAS=0 timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"
timeout 600 python tools/fft_probe.py > gpurun_out/r2h_fft_probe.jsonl 2> gpurun_out/r2h_fft_probe.err; echo "probe rc=$?"
cat gpurun_out/r2h_fft_probe.jsonl
python tools/bench_configs.py cfg2 cfg3 > gpurun_out/r2h_configs.jsonl 2> gpurun_out/r2h_configs.err; echo "configs rc=$?"
QI_BENCH_DTYPE=float64 QI_BENCH_EXTRAS=0 QI_BENCH_CHECKS=0 timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/r2h_bench_f64.json 2> gpurun_out/r2h_bench_f64.err; echo "bench64 rc=$?"
QI_BENCH_CHECKS=0 QI_BENCH_EXTRAS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2h_ncu.log 2>&1; echo "ncu rc=$?"
timeout 300 python tools/profile_exact.py float64 1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'cwtf_interp|cwtf_os|info_plane' --launch-count 6 -f -o gpurun_out/r2h_exact_f64_a python tools/profile_exact.py float64 1 > gpurun_out/r2h_ncu_a.log 2>&1; echo "ncu a rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fft_pass' --launch-count 9 -f -o gpurun_out/r2h_exact_f64_b python tools/profile_exact.py float64 1 > gpurun_out/r2h_ncu_b.log 2>&1; echo "ncu b rc=$?"
ls -la gpurun_out/*.ncu-rep
