#!/bin/bash
# Round-2 session 2: chain on cta_group::2 (cluster of 4) with 2 / 4 epilogue groups, 2-CTA wgrad, PDL; host-loop cost; ncu.
set +e
mkdir -p gpurun_out
S=gpurun_out/s2_summary.txt
: > $S
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" >> $S; }
export ACEZ_WGRAD_2CTA=1
for g in 4 2; do
  ACEZ_CHAIN_V4=1 ACEZ_CHAIN_EPI_GROUPS=$g timeout 240 python -m pytest tests/test_head_chain_gpu.py tests/test_head_gpu.py -m gpu -q > gpurun_out/s2_v4g${g}_tests.log 2>&1
  stamp "chain V4, $g epilogue groups: tests rc=$?"; tail -n 8 gpurun_out/s2_v4g${g}_tests.log | cut -c1-300 >> $S
  ACEZ_CHAIN_V4=1 ACEZ_CHAIN_EPI_GROUPS=$g timeout 100 python tools/probe_step_breakdown.py > gpurun_out/s2_breakdown_v4g$g.log 2>&1
  stamp "breakdown V4 g=$g rc=$?"; cat gpurun_out/s2_breakdown_v4g$g.log >> $S
done
timeout 100 python tools/probe_step_breakdown.py > gpurun_out/s2_breakdown_v2.log 2>&1
stamp "breakdown chain v2 + wgrad2 rc=$?"; cat gpurun_out/s2_breakdown_v2.log >> $S
ACEZ_CHAIN_V4=1 timeout 100 python tools/probe_host_loop.py > gpurun_out/s2_host_loop.log 2>&1
stamp "host loop rc=$?"; tail -n 2 gpurun_out/s2_host_loop.log >> $S
# the rest of the suite on the candidate default (V4 g=4 + wgrad2): DSAC keys, stage tests, stress
ACEZ_CHAIN_V4=1 timeout 400 python -m pytest tests -m gpu -q -x > gpurun_out/s2_suite.log 2>&1
stamp "full GPU suite (V4 g=4, wgrad2) rc=$?"; tail -n 12 gpurun_out/s2_suite.log | cut -c1-300 >> $S
# ncu: one launch each of the chain (fwd, dgrad) and the 2-CTA wgrad
ACEZ_CHAIN_V4=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"head_chain4|gemm2cta" -s 6 -c 3 -o gpurun_out/s2_chain4 -f python tools/probe_step_breakdown.py > gpurun_out/s2_ncu.log 2>&1
stamp "ncu rc=$?"; tail -n 3 gpurun_out/s2_ncu.log >> $S
ACEZ_CHAIN_V4=1 timeout 200 python bench.py --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/s2_bench_v4.json 2> gpurun_out/s2_bench_v4.err
stamp "bench V4 g=4 rc=$?"; cut -c1-330 gpurun_out/s2_bench_v4.json >> $S
stamp done
cat $S
