#!/bin/bash
# Round-2 session 3: pipelined-TMEM-load epilogue of the cta_group::2 chain (k-block order: arrival vs own-first), three-phase
# tail kernel, grouped buffer fill, torch-on-GPU baseline.
set +e
mkdir -p gpurun_out
S=gpurun_out/s3_summary.txt
: > $S
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" >> $S; }
export ACEZ_CHAIN_V4=1 ACEZ_CHAIN_EPI_GROUPS=2
for o in arrival own; do
  ACEZ_CHAIN_ORDER=$o timeout 240 python -m pytest tests/test_head_chain_gpu.py tests/test_head_gpu.py tests/test_loss_gemm_gpu.py -m gpu -q > gpurun_out/s3_${o}_tests.log 2>&1
  stamp "chain V4 g=2 order=$o: tests rc=$?"; grep -E "^FAILED|passed|failed|rel L2|cos" gpurun_out/s3_${o}_tests.log | cut -c1-200 >> $S
  ACEZ_CHAIN_ORDER=$o timeout 100 python tools/probe_step_breakdown.py > gpurun_out/s3_breakdown_$o.log 2>&1
  stamp "breakdown order=$o rc=$?"; cat gpurun_out/s3_breakdown_$o.log >> $S
done
ACEZ_PROBE_COMBOS="1:0" timeout 100 python tools/probe_chain_time.py > gpurun_out/s3_probe.log 2>&1
stamp "chain probe (arrival) rc=$?"; cat gpurun_out/s3_probe.log >> $S
ACEZ_CHAIN_EPI_GROUPS=4 timeout 100 python tools/probe_step_breakdown.py > gpurun_out/s3_breakdown_g4.log 2>&1
stamp "breakdown g=4 (pipelined loads) rc=$?"; cat gpurun_out/s3_breakdown_g4.log >> $S
timeout 200 python tools/bench_buffer_fill.py 64 4 > gpurun_out/s3_fill.log 2>&1
stamp "buffer fill (groups of 8) rc=$?"; tail -n 2 gpurun_out/s3_fill.log >> $S
ACEZ_FILL_BATCH=1 timeout 200 python tools/bench_buffer_fill.py 64 4 > gpurun_out/s3_fill1.log 2>&1
stamp "buffer fill (per image) rc=$?"; tail -n 2 gpurun_out/s3_fill1.log >> $S
ACEZ_CHAIN_ORDER=own timeout 400 python -m pytest tests -m gpu -q > gpurun_out/s3_suite.log 2>&1
stamp "full GPU suite (V4 g=2 own-first) rc=$?"; grep -E "^FAILED|passed|failed" gpurun_out/s3_suite.log | cut -c1-200 >> $S
timeout 300 python bench.py --steps 300 --warmup 5 > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err
stamp "bench (arrival) rc=$?"; cut -c1-1500 gpurun_out/s3_bench.json >> $S; tail -n 3 gpurun_out/s3_bench.err >> $S
stamp done
cat $S
