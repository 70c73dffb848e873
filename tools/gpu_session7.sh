#!/bin/bash
# Round-2 session 7: split-K 256x256 2-CTA wgrad, carve-out hints, point-cloud export, pipeline bench.
set +e
mkdir -p gpurun_out
S=gpurun_out/s7_summary.txt
: > $S
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" >> $S; }
ACEZ_GEMM2_SPLITK=2 timeout 150 python tools/probe_gemm2cta.py > gpurun_out/s7_gemm2cta_split.log 2>&1
stamp "gemm2cta probe, split-K 2 rc=$?"; tail -n 10 gpurun_out/s7_gemm2cta_split.log | cut -c1-300 >> $S
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/s7_suite.log 2>&1
stamp "full GPU suite rc=$?"; grep -E "^FAILED|^ERROR|passed|failed|Error" gpurun_out/s7_suite.log | cut -c1-220 | head -30 >> $S
timeout 100 python tools/probe_step_breakdown.py > gpurun_out/s7_breakdown.log 2>&1
stamp "breakdown (defaults) rc=$?"; cat gpurun_out/s7_breakdown.log >> $S
ACEZ_WGRAD_2CTA_BN=128 timeout 100 python tools/probe_step_breakdown.py > gpurun_out/s7_breakdown_bn128.log 2>&1
stamp "breakdown wgrad 256x128 rc=$?"; cat gpurun_out/s7_breakdown_bn128.log >> $S
ACEZ_SMEM_CARVEOUT=0 timeout 100 python tools/probe_step_breakdown.py > gpurun_out/s7_breakdown_nocarve.log 2>&1
stamp "breakdown without carve-out hints rc=$?"; cat gpurun_out/s7_breakdown_nocarve.log >> $S
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/s7_launches.csv python tools/probe_step_breakdown.py > gpurun_out/s7_ncu_list.log 2>&1
stamp "ncu launch list rc=$?"
timeout 500 python bench.py --steps 300 --warmup 5 > gpurun_out/s7_bench.json 2> gpurun_out/s7_bench.err
stamp "bench rc=$?"; cat gpurun_out/s7_bench.json >> $S; tail -n 3 gpurun_out/s7_bench.err >> $S
stamp done
cat $S
