set -x
mkdir -p gpurun_out
timeout 300 python tools/quick_step.py 10 packed > gpurun_out/r2s_packed.json 2> gpurun_out/r2s_packed.err; echo "rc=$?"
QI_L2K_SCALAR=1 QI_QUICK_CHECK=0 timeout 300 python tools/quick_step.py 10 scalar > gpurun_out/r2s_scalar.json 2> gpurun_out/r2s_scalar.err; echo "rc=$?"
cat gpurun_out/r2s_packed.json gpurun_out/r2s_scalar.json
tail -3 gpurun_out/r2s_packed.err
