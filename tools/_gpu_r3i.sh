set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_kernel --launch-skip 3 --launch-count 1 -f -o gpurun_out/r3i_stft python tools/bench_configs.py cfg2 > gpurun_out/r3i_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r3i_stft.ncu-rep --page source --csv > gpurun_out/r3i_stft_source.csv 2> gpurun_out/r3i_src.err
ncu -i gpurun_out/r3i_stft.ncu-rep --page raw --csv > gpurun_out/r3i_stft_raw.csv 2>> gpurun_out/r3i_src.err
tail -3 gpurun_out/r3i_ncu.log
