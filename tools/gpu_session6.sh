#!/bin/bash
# Round-2 session 6: one-wave tail, tangent-space DSAC* refinement, batched registration loader, gemm2cta cycle counters.
set +e
mkdir -p gpurun_out
S=gpurun_out/s6_summary.txt
: > $S
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" >> $S; }
timeout 500 python -m pytest tests -m gpu -q > gpurun_out/s6_suite.log 2>&1
stamp "full GPU suite rc=$?"; grep -E "^FAILED|^ERROR|passed|failed|Error" gpurun_out/s6_suite.log | cut -c1-220 | head -30 >> $S
timeout 100 python tools/probe_step_breakdown.py > gpurun_out/s6_breakdown.log 2>&1
stamp "breakdown rc=$?"; cat gpurun_out/s6_breakdown.log >> $S
ACEZ_GEMM2_DBG=1 timeout 150 python tools/probe_gemm2cta.py > gpurun_out/s6_gemm2cta.log 2>&1
stamp "gemm2cta probe rc=$?"; tail -n 9 gpurun_out/s6_gemm2cta.log | cut -c1-300 >> $S
timeout 100 python tools/probe_dsac_time.py > gpurun_out/s6_dsac.log 2>&1
stamp "DSAC probe rc=$?"; tail -n 3 gpurun_out/s6_dsac.log >> $S
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"dsac_" -c 2 -o gpurun_out/s6_dsac -f python tools/probe_dsac_time.py > gpurun_out/s6_ncu_dsac.log 2>&1
stamp "ncu dsac rc=$?"
timeout 400 python bench.py --steps 300 --warmup 5 > gpurun_out/s6_bench.json 2> gpurun_out/s6_bench.err
stamp "bench rc=$?"; cat gpurun_out/s6_bench.json >> $S; tail -n 3 gpurun_out/s6_bench.err >> $S
stamp done
cat $S
