set -x
mkdir -p gpurun_out
cd quantum_inferno_b200/csrc
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr"
for v in "1 96" "4 96" "8 96" "2 48" "4 48" "8 48" "4 32" "8 32"; do
  set -- $v
  nvcc $FLAGS -DQI_FFT_LOADS_IN_FLIGHT=$1 -DQI_FFT_TILE_BUDGET_KB=$2 -c qi_capi.cu -o build/qi_capi.o 2>/dev/null && nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libqi_b200.so build/*.o
  echo "VARIANT LU=$1 BUDGET=$2" >> ../../gpurun_out/r2i_fft_variants.txt
  (cd ../..; for c in "float32 8 25" "float64 8 25" "float32 672 18" "float64 336 18" "float32 4096 13"; do python tools/fft_probe.py $c; done) >> ../../gpurun_out/r2i_fft_variants.txt 2>&1
done
nvcc $FLAGS -c qi_capi.cu -o build/qi_capi.o 2>/dev/null && nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libqi_b200.so build/*.o
cd ../..
QI_BENCH_DTYPE=float64 QI_BENCH_EXTRAS=0 QI_BENCH_CHECKS=0 timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/r2i_bench_f64.json 2> gpurun_out/r2i_bench_f64.err; echo "bench64 rc=$?"
QI_BENCH_EXTRAS=0 timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?"
timeout 600 python tools/tc_dft_experiment.py > gpurun_out/r2i_tc_dft.json 2> gpurun_out/r2i_tc_dft.err; echo "tc rc=$?"
tail -2 gpurun_out/r2i_tc_dft.err
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2i_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r2i_tests.log
