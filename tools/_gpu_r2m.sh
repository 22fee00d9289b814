set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2m_topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2m_bench_2gpu.json 2> gpurun_out/r2m_bench_2gpu.err; echo "bench2 rc=$?"
tail -5 gpurun_out/r2m_bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r2m_ref_2gpu.json 2> gpurun_out/r2m_ref_2gpu.err; echo "ref2 rc=$?"
