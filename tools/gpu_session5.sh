#!/bin/bash
# Round-2 session 5: templated epilogue halves + LOP3 mask layout, balanced bias UMMAs in the 2-CTA wgrad, DSAC scoring ILP,
# per-kernel launch list, full bench.
set +e
mkdir -p gpurun_out
S=gpurun_out/s5_summary.txt
: > $S
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" >> $S; }
timeout 150 python tools/probe_gemm2cta.py > gpurun_out/s5_gemm2cta.log 2>&1
stamp "gemm2cta probe rc=$?"; tail -n 8 gpurun_out/s5_gemm2cta.log >> $S
timeout 500 python -m pytest tests -m gpu -q > gpurun_out/s5_suite.log 2>&1
stamp "full GPU suite (defaults: V4 g=2 own-first, wgrad2) rc=$?"; grep -E "^FAILED|^ERROR|passed|failed|Error" gpurun_out/s5_suite.log | cut -c1-220 | head -30 >> $S
for cfg in "2 own" "2 arrival" "4 own"; do
  set -- $cfg
  ACEZ_CHAIN_EPI_GROUPS=$1 ACEZ_CHAIN_ORDER=$2 timeout 100 python tools/probe_step_breakdown.py > gpurun_out/s5_breakdown_g$1_$2.log 2>&1
  stamp "breakdown groups=$1 order=$2 rc=$?"; cat gpurun_out/s5_breakdown_g$1_$2.log >> $S
done
ACEZ_PROBE_COMBOS="1:0" timeout 100 python tools/probe_chain_time.py > gpurun_out/s5_probe.log 2>&1
stamp "chain probe (default) rc=$?"; cat gpurun_out/s5_probe.log >> $S
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s5_launches.csv python tools/probe_step_breakdown.py > gpurun_out/s5_ncu_list.log 2>&1
stamp "ncu launch list rc=$?"
timeout 100 python tools/probe_dsac_time.py > gpurun_out/s5_dsac.log 2>&1
stamp "DSAC probe rc=$?"; tail -n 3 gpurun_out/s5_dsac.log >> $S
timeout 200 python tools/bench_buffer_fill.py 64 4 > gpurun_out/s5_fill.log 2>&1
stamp "buffer fill rc=$?"; tail -n 2 gpurun_out/s5_fill.log >> $S
timeout 400 python bench.py --steps 300 --warmup 5 > gpurun_out/s5_bench.json 2> gpurun_out/s5_bench.err
stamp "bench rc=$?"; cat gpurun_out/s5_bench.json >> $S; tail -n 3 gpurun_out/s5_bench.err >> $S
stamp done
cat $S
