set -x
mkdir -p gpurun_out
timeout 300 python tools/quick_step.py 10 side > gpurun_out/r2t_side.json 2> gpurun_out/r2t_side.err; echo "rc=$?"
QI_MR_NO_SIDE_STREAM=1 QI_QUICK_CHECK=0 timeout 300 python tools/quick_step.py 10 noside > gpurun_out/r2t_noside.json 2> gpurun_out/r2t_noside.err; echo "rc=$?"
cat gpurun_out/r2t_side.json gpurun_out/r2t_noside.json
tail -3 gpurun_out/r2t_side.err
