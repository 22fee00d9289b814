set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2k_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r2k_tests.log
python tools/bench_configs.py cfg1 cfg2 cfg3 > gpurun_out/r2k_configs.jsonl 2> gpurun_out/r2k_configs.err; echo "configs rc=$?"
QI_BENCH_DTYPE=float64 QI_BENCH_EXTRAS=0 timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/r2k_bench_f64.json 2> gpurun_out/r2k_bench_f64.err; echo "bench64 rc=$?"
timeout 600 python tools/fft_probe.py > gpurun_out/r2k_fft_probe.jsonl 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'stft_kernel' --launch-skip 4 --launch-count 2 -f -o gpurun_out/r2k_stft python tools/bench_configs.py cfg2 > gpurun_out/r2k_ncu_stft.log 2>&1; echo "ncu stft rc=$?"
