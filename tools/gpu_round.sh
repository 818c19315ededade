#!/bin/bash
# One GPU-box visit: parity tests, default bench (+ reference arm), ncu launch list, full ncu captures.
# usage: tools/gpu_round.sh TAG [quick]     (outputs under gpurun_out/TAG_*)
tag=${1:-r1}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${tag}_tests.log
tail -3 gpurun_out/${tag}_tests.log
python bench.py --impl reference --steps 50 --warmup 3 > gpurun_out/${tag}_ref_c3.json 2> gpurun_out/${tag}_ref_c3.err
python bench.py > gpurun_out/${tag}_bench_c3.json 2> gpurun_out/${tag}_bench_c3.err; echo "bench rc=$?"
cat gpurun_out/${tag}_bench_c3.json
for w in c4 c2 c5; do python bench.py --workload $w --no-cpu-baseline > gpurun_out/${tag}_bench_$w.json 2>/dev/null; done
python tools/rollout_bench.py > gpurun_out/${tag}_rollout_c5.json 2>/dev/null
# launch list of the bench command (per-launch times are cold-cache and serialised)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${tag}_launches_c3.csv \
  python bench.py --steps 400 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
# full captures of the step kernels (one launch each)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_multi -s 30 -c 1 -f \
  -o gpurun_out/${tag}_full_c3 python tools/quick_time.py 8 65536 400 > gpurun_out/${tag}_ncu_full_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_multi -s 10 -c 1 -f \
  -o gpurun_out/${tag}_full_c4 python tools/quick_time.py 32 131072 200 > gpurun_out/${tag}_ncu_full_c4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_single -s 30 -c 1 -f \
  -o gpurun_out/${tag}_full_c2 python bench.py --workload c2 --steps 400 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_full_c2.log 2>&1
for nb in "8 65536" "8 1048576" "32 131072" "32 1048576" "10 16384"; do STREAMS=2 python tools/quick_time.py $nb 1000; STREAMS=1 python tools/quick_time.py $nb 1000; done
