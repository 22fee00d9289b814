set -x
mkdir -p gpurun_out
python tools/bench_configs.py cfg2 > gpurun_out/r2j_cfg2.jsonl 2> gpurun_out/r2j_cfg2.err; echo "cfg2 rc=$?"
cat gpurun_out/r2j_cfg2.jsonl
timeout 600 python -m pytest tests -m gpu -q -x -k "stft or welch or tukey" > gpurun_out/r2j_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r2j_tests.log
timeout 600 python tools/tc_dft_experiment.py > gpurun_out/r2j_tc_dft.json 2> gpurun_out/r2j_tc_dft.err; echo "tc rc=$?"
tail -2 gpurun_out/r2j_tc_dft.err
cp quantum_inferno_b200/libqi_b200.so /tmp/libqi_default.so
cd quantum_inferno_b200/csrc
make -j16 > /dev/null 2>&1; echo "make rc=$?"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr"
for v in "1 96" "4 96" "8 96" "16 96" "4 48" "8 48" "16 48" "8 32" "16 32" "16 24"; do
  set -- $v
  nvcc $FLAGS -DQI_FFT_LOADS_IN_FLIGHT=$1 -DQI_FFT_TILE_BUDGET_KB=$2 -c qi_capi.cu -o build/qi_capi.o 2>/dev/null && nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libqi_b200.so build/*.o
  echo "VARIANT LU=$1 BUDGET=$2" >> ../../gpurun_out/r2j_fft_variants.txt
  (cd ../..; for c in "float32 8 25" "float64 8 25" "float32 672 18" "float64 336 18"; do python tools/fft_probe.py $c; done) >> ../../gpurun_out/r2j_fft_variants.txt 2>/dev/null
done
cp /tmp/libqi_default.so ../libqi_b200.so
