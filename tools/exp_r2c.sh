#!/bin/bash
# round-2 visit C: prefetch kernel parity + A/B timing; new bench.py smoke
mkdir -p gpurun_out
python -m pytest tests/test_cuda_parity.py -m gpu -q -k "prefetch or reset_modes" > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2c_tests.log
tail -5 gpurun_out/r2c_tests.log
for path in plain prefetch; do
  for nb in "32 131072" "32 1048576" "8 65536" "8 1048576" "16 131072" "16 1048576" "2 4194304"; do
    UAVCA_STEP_PATH=$path STREAMS=1 python tools/quick_time.py $nb 600
  done
done 2>&1 | tee gpurun_out/r2c_times.log
UAVCA_STEP_PATH=prefetch timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_multi_pf -s 10 -c 1 -f \
  -o gpurun_out/r2c_full_c4_pf python tools/quick_time.py 32 131072 200 > gpurun_out/r2c_ncu_full_c4.log 2>&1
python bench.py --steps 400 --warmup 5 > gpurun_out/r2c_bench_c3.json 2> gpurun_out/r2c_bench_c3.err; echo "bench rc=$?"; tail -3 gpurun_out/r2c_bench_c3.err
cat gpurun_out/r2c_bench_c3.json
python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/r2c_ref_c3.json 2> gpurun_out/r2c_ref_c3.err; cat gpurun_out/r2c_ref_c3.json
echo done
