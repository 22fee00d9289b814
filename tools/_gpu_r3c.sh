set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mr_level2k_kernel --launch-skip 6 --launch-count 1 -f -o gpurun_out/r3c_l0 python tools/profile_step.py 1 > gpurun_out/r3c_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r3c_l0.ncu-rep --page source --csv > gpurun_out/r3c_l0_source.csv 2> gpurun_out/r3c_src.err
ncu -i gpurun_out/r3c_l0.ncu-rep --page raw --csv > gpurun_out/r3c_l0_raw.csv 2>> gpurun_out/r3c_src.err
tail -3 gpurun_out/r3c_ncu.log
