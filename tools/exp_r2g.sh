#!/bin/bash
# round-2 visit G (2 GPUs): the sharded path on real GPUs — torchrun parity test, configs[3] strong scaling line, e2e per rank
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2g_topo.txt 2>&1
python -m pytest tests/test_multi_gpu.py -m gpu -q > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2g_tests.log
tail -5 gpurun_out/r2g_tests.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 \
  > gpurun_out/r2g_bench_c4s_2gpu.json 2> gpurun_out/r2g_bench_c4s_2gpu.err; echo "bench c4s x2 rc=$?"; tail -3 gpurun_out/r2g_bench_c4s_2gpu.err
cat gpurun_out/r2g_bench_c4s_2gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 3 \
  > gpurun_out/r2g_bench_c4s_2gpu_k20.json 2> gpurun_out/r2g_bench_c4s_2gpu_k20.err; echo "bench c4s x2 k20 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --impl reference --steps 5 --warmup 1 \
  > gpurun_out/r2g_ref_c4s_2gpu.json 2> gpurun_out/r2g_ref_c4s_2gpu.err; echo "ref x2 rc=$?"; cat gpurun_out/r2g_ref_c4s_2gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29536 bench.py --gpus 2 --workload c3 --no-cpu-baseline \
  > gpurun_out/r2g_bench_c3_2gpu.json 2> gpurun_out/r2g_bench_c3_2gpu.err; echo "bench c3 x2 rc=$?"
echo done
