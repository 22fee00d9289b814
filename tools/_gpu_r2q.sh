set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k regex:'mr_expand_kernel|mr_info_rows|mr_total' --launch-skip 0 --launch-count 10 -f -o /tmp/r2q_full python tools/profile_step.py 1 > gpurun_out/r2q_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/r2q_full.ncu-rep --page raw --csv > gpurun_out/r2q_expand_full_raw.csv 2>/dev/null; ls -la gpurun_out/r2q_expand_full_raw.csv
