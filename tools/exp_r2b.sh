#!/bin/bash
# round-2 visit B: full GPU test suite with the new scan + rollout; rollout timings
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2b_tests.log
tail -15 gpurun_out/r2b_tests.log
for lib in "" ro8 ro5; do
  if [ -n "$lib" ]; then export UAVCA_LIB=$PWD/build/variants/libuavca_$lib.so; else unset UAVCA_LIB; fi
  for cfg in "multi 8 65536 32 philox" "multi 8 65536 32 block" "multi 8 65536 8 block" "multi 32 131072 16 philox" "multi 32 131072 16 block" \
             "multi 32 1048576 4 block 5" "multi 10 16384 32 philox" "multi 16 131072 16 block" "single 1 65536 64 philox" "single 1 65536 64 block"; do
    python tools/rollout_time.py $cfg
  done
done 2>&1 | tee gpurun_out/r2b_rollout.log
unset UAVCA_LIB
echo done
