#!/bin/bash
# round-2 visit E: full GPU suite (float64 world, N > 32), the new default bench line, ncu launch list + full captures
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2e_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2e_tests.log
tail -12 gpurun_out/r2e_tests.log
python bench.py > gpurun_out/r2e_bench_c3.json 2> gpurun_out/r2e_bench_c3.err; echo "bench c3 rc=$?"; tail -2 gpurun_out/r2e_bench_c3.err
python bench.py --steps 20 --warmup 3 --no-acting > gpurun_out/r2e_bench_c3_k20.json 2> gpurun_out/r2e_bench_c3_k20.err; echo "bench c3 k20 rc=$?"
python bench.py --workload c4 --no-cpu-baseline > gpurun_out/r2e_bench_c4.json 2> gpurun_out/r2e_bench_c4.err; echo "bench c4 rc=$?"; tail -2 gpurun_out/r2e_bench_c4.err
python bench.py --workload c2 --no-cpu-baseline > gpurun_out/r2e_bench_c2.json 2> gpurun_out/r2e_bench_c2.err; echo "bench c2 rc=$?"
python bench.py --workload c5 --no-cpu-baseline > gpurun_out/r2e_bench_c5.json 2> gpurun_out/r2e_bench_c5.err; echo "bench c5 rc=$?"
python bench.py --workload c1 > gpurun_out/r2e_bench_c1.json 2> gpurun_out/r2e_bench_c1.err; echo "bench c1 rc=$?"
# launch list of the default bench command (per-launch times are cold-cache and serialised)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2e_launches_c3.csv \
  python bench.py --steps 200 --warmup 3 --no-cpu-baseline --no-acting > gpurun_out/r2e_ncu_bench.log 2>&1
# full captures (one launch each), summarised on the box; two reports travel back for the source pages
cap() {  # name kernel-regex skip command...
  name=$1; k=$2; skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/$name "$@" > gpurun_out/$name.log 2>&1
  python tools/summarize_ncu.py full gpurun_out/$name.ncu-rep gpurun_out/$name.md --title "$name" > /dev/null 2>&1
}
cap r2e_full_rollout_n8 rollout_multi 2 python tools/rollout_time.py multi 8 65536 32 block 3
cap r2e_full_rollout_n32 rollout_multi 2 python tools/rollout_time.py multi 32 131072 16 block 3
cap r2e_full_step_n8 step_multi_kernel 30 python tools/quick_time.py 8 65536 400
cap r2e_full_step_n32 step_multi_kernel 10 python tools/quick_time.py 32 131072 200
cap r2e_full_rollout_single rollout_single 2 python tools/rollout_time.py single 1 65536 64 block 3
rm -f gpurun_out/r2e_full_rollout_n8.ncu-rep gpurun_out/r2e_full_step_n8.ncu-rep gpurun_out/r2e_full_rollout_single.ncu-rep
{
for lib in "" tree; do
  if [ -n "$lib" ]; then export UAVCA_LIB=$PWD/build/variants/libuavca_$lib.so; else unset UAVCA_LIB; fi
  for nb in "32 131072" "32 1048576" "16 131072" "8 65536"; do STREAMS=1 python tools/quick_time.py $nb 600; done
  python tools/rollout_time.py multi 32 131072 16 block
  python tools/rollout_time.py multi 8 65536 32 block
done
unset UAVCA_LIB
} 2>&1 | tee gpurun_out/r2e_tree.log
python tools/policy_bench.py > gpurun_out/r2e_policy_bench.log 2>&1; cat gpurun_out/r2e_policy_bench.log
python bench.py --workload c5r > gpurun_out/r2e_bench_c5r.json 2> gpurun_out/r2e_bench_c5r.err; echo "bench c5r rc=$?"
ls -la gpurun_out | head -40
echo done
