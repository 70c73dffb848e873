#!/bin/bash
# Round-2 session 8: on-chip split-K wgrad (cluster of four), SPLIT chain epilogue, fused gather + schedule.
set +e
mkdir -p gpurun_out
S=gpurun_out/s8_summary.txt
: > $S
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" >> $S; }
ACEZ_GEMM2_SPLITK=2 timeout 150 python tools/probe_gemm2cta.py > gpurun_out/s8_gemm2cta_c4.log 2>&1
stamp "gemm2cta probe, on-chip split-K 2 (256-column tiles) rc=$?"; tail -n 12 gpurun_out/s8_gemm2cta_c4.log | cut -c1-300 >> $S
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/s8_suite.log 2>&1
stamp "full GPU suite (defaults) rc=$?"; grep -E "^FAILED|^ERROR|passed|failed|Error" gpurun_out/s8_suite.log | cut -c1-220 | head -30 >> $S
timeout 100 python tools/probe_step_breakdown.py > gpurun_out/s8_breakdown.log 2>&1
stamp "breakdown (defaults) rc=$?"; cat gpurun_out/s8_breakdown.log >> $S
ACEZ_WGRAD_2CTA_BN=256 timeout 300 python -m pytest tests/test_head_gpu.py tests/test_head_chain_gpu.py -m gpu -q > gpurun_out/s8_c4_tests.log 2>&1
stamp "wgrad cluster-of-four tests rc=$?"; grep -E "^FAILED|^ERROR|passed|failed|Error" gpurun_out/s8_c4_tests.log | cut -c1-220 | head >> $S
ACEZ_WGRAD_2CTA_BN=256 timeout 100 python tools/probe_step_breakdown.py > gpurun_out/s8_breakdown_c4.log 2>&1
stamp "breakdown wgrad cluster of four rc=$?"; cat gpurun_out/s8_breakdown_c4.log >> $S
ACEZ_CHAIN_EPI_SPLIT=1 timeout 300 python -m pytest tests/test_head_gpu.py tests/test_head_chain_gpu.py -m gpu -q > gpurun_out/s8_split_tests.log 2>&1
stamp "SPLIT epilogue tests rc=$?"; grep -E "^FAILED|^ERROR|passed|failed|Error" gpurun_out/s8_split_tests.log | cut -c1-220 | head >> $S
ACEZ_CHAIN_EPI_SPLIT=1 timeout 100 python tools/probe_step_breakdown.py > gpurun_out/s8_breakdown_split.log 2>&1
stamp "breakdown SPLIT epilogue rc=$?"; cat gpurun_out/s8_breakdown_split.log >> $S
ACEZ_CHAIN_EPI_SPLIT=1 ACEZ_PROBE_COMBOS="1:0" timeout 100 python tools/probe_chain_time.py > gpurun_out/s8_probe_split.log 2>&1
stamp "chain probe SPLIT rc=$?"; cat gpurun_out/s8_probe_split.log >> $S
timeout 500 python bench.py --steps 300 --warmup 5 > gpurun_out/s8_bench.json 2> gpurun_out/s8_bench.err
stamp "bench rc=$?"; cat gpurun_out/s8_bench.json >> $S; tail -n 3 gpurun_out/s8_bench.err >> $S
stamp done
cat $S
