#!/bin/bash
# Round-2 session 4: shared-space ld/st in the chain / GEMM epilogues, device schedule, graph-captured refinement, FP32 DSAC scoring.
set +e
mkdir -p gpurun_out
S=gpurun_out/s4_summary.txt
: > $S
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" >> $S; }
export ACEZ_CHAIN_V4=1 ACEZ_CHAIN_EPI_GROUPS=2
for o in arrival own; do
  ACEZ_CHAIN_ORDER=$o timeout 100 python tools/probe_step_breakdown.py > gpurun_out/s4_breakdown_$o.log 2>&1
  stamp "breakdown order=$o rc=$?"; cat gpurun_out/s4_breakdown_$o.log >> $S
done
ACEZ_PROBE_COMBOS="1:0" timeout 100 python tools/probe_chain_time.py > gpurun_out/s4_probe.log 2>&1
stamp "chain probe (arrival) rc=$?"; cat gpurun_out/s4_probe.log >> $S
ACEZ_CHAIN_ORDER=own ACEZ_PROBE_COMBOS="1:0" timeout 100 python tools/probe_chain_time.py > gpurun_out/s4_probe_own.log 2>&1
stamp "chain probe (own) rc=$?"; cat gpurun_out/s4_probe_own.log >> $S
ACEZ_CHAIN_ORDER=own timeout 500 python -m pytest tests -m gpu -q > gpurun_out/s4_suite.log 2>&1
stamp "full GPU suite (V4 g=2 own-first) rc=$?"; grep -E "^FAILED|^ERROR|passed|failed|Error" gpurun_out/s4_suite.log | cut -c1-220 | head -30 >> $S
timeout 100 python tools/probe_dsac_time.py > gpurun_out/s4_dsac.log 2>&1
stamp "DSAC probe rc=$?"; tail -n 3 gpurun_out/s4_dsac.log >> $S
timeout 200 python tools/probe_fill.py > gpurun_out/s4_fill_profile.log 2>&1
stamp "fill profile rc=$?"; head -n 45 gpurun_out/s4_fill_profile.log | cut -c1-160 >> $S
timeout 100 python tools/probe_host_loop.py > gpurun_out/s4_host_loop.log 2>&1
stamp "host loop rc=$?"; tail -n 2 gpurun_out/s4_host_loop.log >> $S
timeout 300 python bench.py --steps 300 --warmup 5 > gpurun_out/s4_bench.json 2> gpurun_out/s4_bench.err
stamp "bench (arrival) rc=$?"; cut -c1-400 gpurun_out/s4_bench.json >> $S; tail -n 3 gpurun_out/s4_bench.err >> $S
stamp done
cat $S
