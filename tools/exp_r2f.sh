#!/bin/bash
# round-2 visit F: small-N mixed sweep + policy kernel v2 (heads on the CUDA cores) + 3-launch acting step: full suite, A/B timings
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2f_tests.log
tail -8 gpurun_out/r2f_tests.log
python tools/policy_bench.py > gpurun_out/r2f_policy_bench.log 2>&1; cat gpurun_out/r2f_policy_bench.log
python tools/rollout_bench.py > gpurun_out/r2f_rollout_c5.jsonl 2> gpurun_out/r2f_rollout_c5.err; cat gpurun_out/r2f_rollout_c5.jsonl
{
for lib in "" tree ro7 ro6; do
  if [ -n "$lib" ]; then export UAVCA_LIB=$PWD/build/variants/libuavca_$lib.so; else unset UAVCA_LIB; fi
  if [ "$lib" == "" ] || [ "$lib" == "tree" ]; then
    for nb in "32 131072" "32 1048576" "16 131072" "8 65536" "10 16384"; do STREAMS=1 python tools/quick_time.py $nb 600; done
    STREAMS=2 python tools/quick_time.py 8 65536 2000
  fi
  python tools/rollout_time.py multi 32 131072 16 block
  python tools/rollout_time.py multi 16 131072 16 block
  python tools/rollout_time.py multi 8 65536 32 block
  python tools/rollout_time.py multi 8 65536 32 philox
  python tools/rollout_time.py multi 10 16384 32 block
  python tools/rollout_time.py multi 4 131072 32 block
done
unset UAVCA_LIB
} 2>&1 | tee gpurun_out/r2f_ab.log
python bench.py > gpurun_out/r2f_bench_c3.json 2> gpurun_out/r2f_bench_c3.err; echo "bench c3 rc=$?"; tail -2 gpurun_out/r2f_bench_c3.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:policy_act -s 3 -c 1 -f \
  -o gpurun_out/r2f_full_policy python tools/policy_bench.py > gpurun_out/r2f_ncu_policy.log 2>&1
python tools/summarize_ncu.py full gpurun_out/r2f_full_policy.ncu-rep gpurun_out/r2f_full_policy.md --title "r2 policy_act_kernel v2" > /dev/null 2>&1
rm -f gpurun_out/r2f_full_policy.ncu-rep
timeout 600 ncu --set full --clock-control none -k regex:step_multi_kernel -s 30 -c 1 -f -o gpurun_out/r2f_full_step_n8 python tools/quick_time.py 8 65536 400 > /dev/null 2>&1
python tools/summarize_ncu.py full gpurun_out/r2f_full_step_n8.ncu-rep gpurun_out/r2f_full_step_n8.md --title "r2 step_multi_kernel<8> (mixed sweep)" > /dev/null 2>&1
rm -f gpurun_out/r2f_full_step_n8.ncu-rep
echo done
