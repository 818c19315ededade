#!/bin/bash
# round-2 visit D: full GPU suite (scores, trajectory, rollout, 1000-step N=32 window), wave shaping A/B, register variants, bench lines
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2d_tests.log
tail -12 gpurun_out/r2d_tests.log
{
for ws in 0 1; do
  export UAVCA_WAVE_SHAPE=$ws
  echo "== UAVCA_WAVE_SHAPE=$ws"
  for nb in "8 65536" "16 32768" "4 131072" "10 16384" "32 16384"; do STREAMS=1 python tools/quick_time.py $nb 2000; STREAMS=2 python tools/quick_time.py $nb 2000; done
  for cfg in "multi 8 65536 32 block" "multi 8 65536 32 philox" "multi 10 16384 32 block" "multi 16 32768 32 block" "single 1 65536 64 block"; do python tools/rollout_time.py $cfg; done
done
unset UAVCA_WAVE_SHAPE
for lib in r48 r56; do
  export UAVCA_LIB=$PWD/build/variants/libuavca_$lib.so
  for nb in "32 131072" "32 1048576" "8 65536"; do STREAMS=1 python tools/quick_time.py $nb 600; done
  python tools/rollout_time.py multi 32 131072 16 block
done
unset UAVCA_LIB
python tools/rollout_time.py multi 32 131072 16 block
python tools/rollout_time.py multi 32 1048576 4 block 5
} 2>&1 | tee gpurun_out/r2d_times.log
python bench.py > gpurun_out/r2d_bench_c3.json 2> gpurun_out/r2d_bench_c3.err; echo "bench c3 rc=$?"; tail -2 gpurun_out/r2d_bench_c3.err
python bench.py --workload c4 --no-cpu-baseline > gpurun_out/r2d_bench_c4.json 2> gpurun_out/r2d_bench_c4.err; echo "bench c4 rc=$?"; tail -2 gpurun_out/r2d_bench_c4.err
python bench.py --workload c2 --no-cpu-baseline > gpurun_out/r2d_bench_c2.json 2> gpurun_out/r2d_bench_c2.err; echo "bench c2 rc=$?"
python bench.py --workload c1 > gpurun_out/r2d_bench_c1.json 2> gpurun_out/r2d_bench_c1.err; echo "bench c1 rc=$?"; cat gpurun_out/r2d_bench_c1.json
python bench.py --workload c5r > gpurun_out/r2d_bench_c5r.json 2> gpurun_out/r2d_bench_c5r.err; echo "bench c5r rc=$?"; cat gpurun_out/r2d_bench_c5r.json
python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/r2d_ref_c3.json 2> gpurun_out/r2d_ref_c3.err; cat gpurun_out/r2d_ref_c3.json
python bench.py --impl reference --workload c1 > gpurun_out/r2d_ref_c1.json 2> gpurun_out/r2d_ref_c1.err; cat gpurun_out/r2d_ref_c1.json
echo done
