#!/bin/bash
# Round-2 session 9: new defaults (split epilogue, cluster-of-four weight gradients, schedule in its own block): suite, breakdown,
# bench, ncu launch list of the bench command, ncu --set full of every kernel of the step and of the two DSAC* kernels.
set +e
mkdir -p gpurun_out
S=gpurun_out/s9_summary.txt
: > $S
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" >> $S; }
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/s9_suite.log 2>&1
stamp "full GPU suite rc=$?"; tail -n 4 gpurun_out/s9_suite.log | cut -c1-300 >> $S
timeout 120 python tools/probe_step_breakdown.py > gpurun_out/s9_breakdown.log 2>&1
stamp "breakdown rc=$?"; cat gpurun_out/s9_breakdown.log >> $S
timeout 600 python bench.py --steps 300 --warmup 5 > gpurun_out/s9_bench.json 2> gpurun_out/s9_bench.err
stamp "bench rc=$?"; cut -c1-2500 gpurun_out/s9_bench.json >> $S; tail -n 3 gpurun_out/s9_bench.err | cut -c1-300 >> $S
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s9_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-torch-baseline --no-pipeline > gpurun_out/s9_ncu_list.log 2>&1
stamp "ncu launch list rc=$?"; tail -n 2 gpurun_out/s9_ncu_list.log | cut -c1-300 >> $S
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"head_chain4|gemm2cta|head_tail|adamw|gather_rows" -s 18 -c 6 -o gpurun_out/s9_step -f python tools/probe_step_breakdown.py > gpurun_out/s9_ncu_step.log 2>&1
stamp "ncu step kernels rc=$?"; tail -n 2 gpurun_out/s9_ncu_step.log | cut -c1-300 >> $S
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"dsac_" -c 2 -o gpurun_out/s9_dsac -f python tools/probe_dsac_time.py > gpurun_out/s9_ncu_dsac.log 2>&1
stamp "ncu dsac rc=$?"; tail -n 2 gpurun_out/s9_ncu_dsac.log | cut -c1-300 >> $S
stamp done
cat $S
