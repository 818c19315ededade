#!/bin/bash
# One GPU-box visit: parity tests, default bench (+ reference arm), ncu launch list, full ncu captures.
# usage: tools/gpu_round.sh TAG     (outputs under gpurun_out/TAG_*)
tag=${1:-r1}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${tag}_tests.log
tail -3 gpurun_out/${tag}_tests.log
python bench.py --impl reference --steps 50 --warmup 3 > gpurun_out/${tag}_ref_c3.json 2> gpurun_out/${tag}_ref_c3.err
python bench.py > gpurun_out/${tag}_bench_c3.json 2> gpurun_out/${tag}_bench_c3.err; echo "bench rc=$?"
cat gpurun_out/${tag}_bench_c3.json
UAVCA_STEP_PATH=lanes python bench.py --no-cpu-baseline > gpurun_out/${tag}_bench_c3_lanes.json 2>/dev/null
python bench.py --workload c4 --no-cpu-baseline > gpurun_out/${tag}_bench_c4.json 2>/dev/null
UAVCA_STEP_PATH=lanes python bench.py --workload c4 --no-cpu-baseline > gpurun_out/${tag}_bench_c4_lanes.json 2>/dev/null
python bench.py --workload c2 --no-cpu-baseline > gpurun_out/${tag}_bench_c2.json 2>/dev/null
# launch list of the bench command (per-launch times are cold-cache and serialised)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_c3.csv \
  python bench.py --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
# full captures of the step kernels
for p in lanes auto; do
  UAVCA_STEP_PATH=$p timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_multi -s 30 -c 2 -f \
    -o gpurun_out/${tag}_full_c3_$p python tools/quick_time.py 8 65536 400 > gpurun_out/${tag}_ncu_full_$p.log 2>&1
done
UAVCA_STEP_PATH=lanes timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_multi -s 10 -c 1 -f \
  -o gpurun_out/${tag}_full_n32_lanes python tools/quick_time.py 32 131072 200 > gpurun_out/${tag}_ncu_full_n32.log 2>&1
for p in lanes auto; do for nb in "8 65536" "8 1048576" "32 131072" "32 1048576" "10 16384"; do UAVCA_STEP_PATH=$p python tools/quick_time.py $nb 1000; done; done
