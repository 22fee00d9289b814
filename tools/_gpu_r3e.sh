set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cwtf_interp_kernel --launch-skip 0 --launch-count 1 -f -o gpurun_out/r3e_interp python tools/bench_configs.py cfg3 > gpurun_out/r3e_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r3e_interp.ncu-rep --page source --csv > gpurun_out/r3e_interp_source.csv 2> gpurun_out/r3e_src.err
ncu -i gpurun_out/r3e_interp.ncu-rep --page raw --csv > gpurun_out/r3e_interp_raw.csv 2>> gpurun_out/r3e_src.err
tail -3 gpurun_out/r3e_ncu.log
