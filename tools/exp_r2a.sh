#!/bin/bash
# round-2 visit A: parity of the new pair scan + A/B timings of CTA shapes
mkdir -p gpurun_out
python -m pytest tests/test_cuda_parity.py -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2a_tests.log
tail -5 gpurun_out/r2a_tests.log
for lib in r1 "" w2 w1 w8 w4r56; do
  if [ -n "$lib" ]; then export UAVCA_LIB=$PWD/build/variants/libuavca_$lib.so; else unset UAVCA_LIB; fi
  for nb in "32 131072" "32 1048576" "8 65536" "16 131072"; do
    STREAMS=1 python tools/quick_time.py $nb 1000
    STREAMS=2 python tools/quick_time.py $nb 1000
  done
done 2>&1 | tee gpurun_out/r2a_times.log
unset UAVCA_LIB
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_multi -s 10 -c 1 -f \
  -o gpurun_out/r2a_full_c4 python tools/quick_time.py 32 131072 200 > gpurun_out/r2a_ncu_full_c4.log 2>&1
echo done
