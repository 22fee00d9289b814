set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mr_expand_kernel --launch-skip 3 --launch-count 3 -f -o gpurun_out/r3k_small python tools/profile_step.py 1 > gpurun_out/r3k_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r3k_small.ncu-rep --page raw --csv > gpurun_out/r3k_small_raw.csv 2> gpurun_out/r3k_src.err
ncu -i gpurun_out/r3k_small.ncu-rep --page source --csv > gpurun_out/r3k_small_source.csv 2>> gpurun_out/r3k_src.err
tail -3 gpurun_out/r3k_ncu.log; ls -la gpurun_out | tail -4
