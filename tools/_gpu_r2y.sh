set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mr_expand_kernel --launch-skip 4 --launch-count 1 -f -o gpurun_out/r2z_expand python tools/profile_step.py 1 > gpurun_out/r2z_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r2z_expand.ncu-rep --page source --csv > gpurun_out/r2z_expand_source.csv 2> gpurun_out/r2z_src.err
ncu -i gpurun_out/r2z_expand.ncu-rep --page raw --csv > gpurun_out/r2z_expand_raw.csv 2>> gpurun_out/r2z_src.err
ls -la gpurun_out/ | tail -5
tail -3 gpurun_out/r2z_ncu.log
