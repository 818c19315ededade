#!/bin/bash
# One GPU-box visit: parity tests, bench lines of every workload (+ the CPU arms), ncu launch list of the default bench
# command, full ncu captures of the hot kernels (summarised on the box: only two reports travel back).
# usage: tools/gpu_round.sh TAG     (outputs under gpurun_out/TAG_*)
tag=${1:-r2}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${tag}_tests.log
tail -4 gpurun_out/${tag}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
t0=$(date +%s); python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/${tag}_ref_c3.json 2> gpurun_out/${tag}_ref_c3.err; echo "reference arm wall=$(( $(date +%s) - t0 )) s"
t0=$(date +%s); python bench.py > gpurun_out/${tag}_bench_c3.json 2> gpurun_out/${tag}_bench_c3.err; echo "bench rc=$? wall=$(( $(date +%s) - t0 )) s"; tail -2 gpurun_out/${tag}_bench_c3.err
python bench.py --steps 20 --warmup 3 --no-acting --no-cpu-baseline > gpurun_out/${tag}_bench_c3_k20.json 2>/dev/null
for w in c4 c2 c5; do python bench.py --workload $w --no-cpu-baseline > gpurun_out/${tag}_bench_$w.json 2>/dev/null; done
python bench.py --workload c1 > gpurun_out/${tag}_bench_c1.json 2>/dev/null
python bench.py --workload c5r > gpurun_out/${tag}_bench_c5r.json 2>/dev/null
python bench.py --impl reference --workload c1 > gpurun_out/${tag}_ref_c1.json 2>/dev/null
python tools/rollout_bench.py > gpurun_out/${tag}_rollout_c5.jsonl 2>/dev/null
# the configs[3] shard one of eight GPUs holds (N=32, B=131,072), one launch per step and K steps per launch
{ STREAMS=1 python tools/quick_time.py 32 131072 600; python tools/rollout_time.py multi 32 131072 31 block; python tools/rollout_time.py multi 32 131072 31 philox; } > gpurun_out/${tag}_c4_shard.log 2>&1
# launch list of the default bench command (per-launch times are cold-cache and serialised)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${tag}_launches_c3.csv \
  python bench.py --steps 200 --warmup 3 --no-cpu-baseline --no-acting > gpurun_out/${tag}_ncu_bench.log 2>&1
python tools/summarize_ncu.py list gpurun_out/${tag}_launches_c3.csv gpurun_out/${tag}_launches_c3.md > /dev/null 2>&1
rm -f gpurun_out/${tag}_launches_c3.csv
# full captures (one launch each)
cap() {  # name kernel-regex skip command...
  name=$1; k=$2; skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/$name "$@" > gpurun_out/$name.log 2>&1
  python tools/summarize_ncu.py full gpurun_out/$name.ncu-rep gpurun_out/$name.md --title "$name" > /dev/null 2>&1
}
cap ${tag}_full_rollout_n8 rollout_multi 2 python tools/rollout_time.py multi 8 65536 32 block 3
cap ${tag}_full_rollout_n32 rollout_multi 2 python tools/rollout_time.py multi 32 131072 31 block 3
cap ${tag}_full_step_n8 step_multi_kernel 30 python tools/quick_time.py 8 65536 400
cap ${tag}_full_step_n32 step_multi_kernel 10 python tools/quick_time.py 32 131072 200
cap ${tag}_full_rollout_single rollout_single 2 python tools/rollout_time.py single 1 65536 64 block 3
cap ${tag}_full_policy policy_act 3 python tools/policy_bench.py
cap ${tag}_full_step_ring_n10 step_multi_ring 6 python tools/rollout_bench.py 16384 10 100 fused
rm -f gpurun_out/${tag}_full_step_ring_n10.ncu-rep
rm -f gpurun_out/${tag}_full_rollout_n8.ncu-rep gpurun_out/${tag}_full_step_n8.ncu-rep gpurun_out/${tag}_full_rollout_single.ncu-rep gpurun_out/${tag}_full_policy.ncu-rep
ls gpurun_out | head -60
echo done
