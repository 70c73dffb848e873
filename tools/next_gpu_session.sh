#!/bin/bash
# First GPU call of the next session: validates every opt-in experiment written without GPU time at the end of round 1.
#   gpurun --timeout 900 -- 'bash tools/next_gpu_session.sh'          (about 12-15 minutes of run time, one GPU)
# Everything lands in gpurun_out/next_*.{log,txt}; the summary is printed at the end. Ordered by value.
set +e
mkdir -p gpurun_out
S=gpurun_out/next_summary.txt
: > $S
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" >> $S; }

# 0. baseline of the box: default path (fused chain v2), tests + step breakdown
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/next_suite_default.log 2>&1
stamp "GPU suite, default path rc=$?"; tail -n 2 gpurun_out/next_suite_default.log >> $S
timeout 100 python tools/probe_step_breakdown.py > gpurun_out/next_breakdown_default.log 2>&1
stamp "breakdown default rc=$?"; cat gpurun_out/next_breakdown_default.log >> $S

# 1. chain V3 (arrival-order k-blocks, half-box publication, epilogue diet): parity, then timing + timeline
ACEZ_CHAIN_V3=1 timeout 200 python -m pytest tests/test_head_chain_gpu.py tests/test_head_gpu.py -m gpu -x -q > gpurun_out/next_v3_tests.log 2>&1
rc_v3=$?
stamp "chain V3 tests rc=$rc_v3"; tail -n 6 gpurun_out/next_v3_tests.log >> $S
if [ $rc_v3 -ne 0 ]; then
  ACEZ_CHAIN_V3=1 timeout 100 python tools/diag_chain.py 384 > gpurun_out/next_v3_diag.log 2>&1
  stamp "chain V3 diag rc=$?"; tail -n 24 gpurun_out/next_v3_diag.log >> $S
fi
ACEZ_CHAIN_V3=1 ACEZ_PROBE_COMBOS="1:0" timeout 100 python tools/probe_chain_time.py > gpurun_out/next_v3_probe.log 2>&1
stamp "chain V3 probe rc=$?"; cat gpurun_out/next_v3_probe.log >> $S
for o in 1 2 3; do   # attribute the V3 gain: 1 = no half-box publication, 2 = own-first k-block order, 3 = both off (diet only)
  ACEZ_CHAIN_V3=1 ACEZ_CHAIN_V3_OPTS=$o ACEZ_PROBE_COMBOS="1:0" timeout 100 python tools/probe_chain_time.py > gpurun_out/next_v3_probe_o$o.log 2>&1
  stamp "chain V3 probe, V3_OPTS=$o rc=$?"; head -n 2 gpurun_out/next_v3_probe_o$o.log >> $S
done
ACEZ_PROBE_COMBOS="1:0" timeout 100 python tools/probe_chain_time.py > gpurun_out/next_v2_probe.log 2>&1
stamp "chain v2 probe (same box) rc=$?"; head -n 3 gpurun_out/next_v2_probe.log >> $S

# 2. cta_group::2 probe GEMM (correctness cases, then timing on the weight-gradient shape)
timeout 120 python tools/probe_gemm2cta.py > gpurun_out/next_gemm2cta.log 2>&1
stamp "gemm2cta probe rc=$?"; cat gpurun_out/next_gemm2cta.log >> $S

# 2b. the batched weight-gradient GEMM on cta_group::2 tiles (only meaningful if the probe passed)
ACEZ_WGRAD_2CTA=1 timeout 200 python -m pytest tests/test_head_gpu.py tests/test_head_chain_gpu.py -m gpu -x -q > gpurun_out/next_wgrad2_tests.log 2>&1
stamp "wgrad 2-CTA tests rc=$?"; tail -n 3 gpurun_out/next_wgrad2_tests.log >> $S
ACEZ_WGRAD_2CTA=1 timeout 100 python tools/probe_step_breakdown.py > gpurun_out/next_breakdown_wgrad2.log 2>&1
stamp "breakdown wgrad 2-CTA (256x128 tiles) rc=$?"; cat gpurun_out/next_breakdown_wgrad2.log >> $S

# 2c. layer chain on cta_group::2 / cluster of 4 (only meaningful if the 2-CTA probe passed)
ACEZ_CHAIN_V4=1 timeout 200 python -m pytest tests/test_head_chain_gpu.py tests/test_head_gpu.py -m gpu -x -q > gpurun_out/next_v4_tests.log 2>&1
rc_v4=$?
stamp "chain V4 tests rc=$rc_v4"; tail -n 6 gpurun_out/next_v4_tests.log >> $S
if [ $rc_v4 -ne 0 ]; then
  ACEZ_CHAIN_V4=1 timeout 100 python tools/diag_chain.py 384 > gpurun_out/next_v4_diag.log 2>&1
  stamp "chain V4 diag rc=$?"; tail -n 24 gpurun_out/next_v4_diag.log >> $S
fi
ACEZ_CHAIN_V4=1 timeout 100 python tools/probe_step_breakdown.py > gpurun_out/next_breakdown_v4.log 2>&1
stamp "breakdown chain V4 rc=$?"; cat gpurun_out/next_breakdown_v4.log >> $S

# 3. tail kernel at 2 CTAs / SM
ACEZ_TAIL_OCC2=1 timeout 150 python -m pytest tests/test_head_gpu.py -m gpu -x -q > gpurun_out/next_occ2_tests.log 2>&1
stamp "tail occ2 tests rc=$?"; tail -n 2 gpurun_out/next_occ2_tests.log >> $S
ACEZ_TAIL_OCC2=1 timeout 100 python tools/probe_step_breakdown.py > gpurun_out/next_breakdown_occ2.log 2>&1
stamp "breakdown tail occ2 rc=$?"; cat gpurun_out/next_breakdown_occ2.log >> $S

# 3b. stress cases of configs[4] (hypothesis sweep to 4096, 4096 images per call, 1 M-row buffer epochs)
ACEZ_TEST_EXTRA=1 timeout 300 python -m pytest tests/test_stress_gpu.py -m gpu -q > gpurun_out/next_stress.log 2>&1
stamp "stress tests rc=$?"; tail -n 4 gpurun_out/next_stress.log >> $S

# 3b2. fc3 gradient kernels on a side stream under the dgrad chain
ACEZ_FC3_OVERLAP=1 timeout 150 python -m pytest tests/test_head_gpu.py tests/test_head_chain_gpu.py -m gpu -x -q > gpurun_out/next_fc3ov_tests.log 2>&1
stamp "fc3 overlap tests rc=$?"; tail -n 2 gpurun_out/next_fc3ov_tests.log >> $S
ACEZ_FC3_OVERLAP=1 timeout 100 python tools/probe_step_breakdown.py > gpurun_out/next_breakdown_fc3ov.log 2>&1
stamp "breakdown fc3 overlap rc=$?"; cat gpurun_out/next_breakdown_fc3ov.log >> $S

# 3b3. DSAC* kernels at higher occupancy (80 / 128 registers)
ACEZ_DSAC_OCC=1 timeout 200 python -m pytest tests/test_dsac_gpu.py -m gpu -x -q > gpurun_out/next_dsac_occ_tests.log 2>&1
stamp "DSAC occupancy variant tests rc=$?"; tail -n 2 gpurun_out/next_dsac_occ_tests.log >> $S
for v in 0 1; do
  ACEZ_DSAC_OCC=$v timeout 100 python tools/probe_dsac_time.py > gpurun_out/next_dsac_occ$v.log 2>&1
  stamp "DSAC probe occ=$v rc=$?"; tail -n 6 gpurun_out/next_dsac_occ$v.log >> $S
done

# 3b4. buffer-fill rate (encoder + sampling + scatter), the number DESIGN.md section 7 lists as missing
timeout 200 python tools/bench_buffer_fill.py 64 4 > gpurun_out/next_buffer_fill.log 2>&1
stamp "buffer fill rc=$?"; tail -n 2 gpurun_out/next_buffer_fill.log >> $S
ACEZ_FILL_BATCH=8 timeout 200 python -m pytest tests/test_stage_gpu.py -m gpu -x -q > gpurun_out/next_fill_batch_tests.log 2>&1
stamp "batched buffer fill: stage tests (bit-exact indices, mapping + registration) rc=$?"; tail -n 2 gpurun_out/next_fill_batch_tests.log >> $S
ACEZ_FILL_BATCH=8 timeout 200 python tools/bench_buffer_fill.py 64 4 > gpurun_out/next_buffer_fill_b8.log 2>&1
stamp "buffer fill, batches of 8 rc=$?"; tail -n 2 gpurun_out/next_buffer_fill_b8.log >> $S

# 3c. optimiser state (34 MB) pinned in L2
ACEZ_L2_PERSIST=1 timeout 150 python -m pytest tests/test_head_gpu.py -m gpu -x -q > gpurun_out/next_l2_tests.log 2>&1
stamp "L2 persistence tests rc=$?"; tail -n 2 gpurun_out/next_l2_tests.log >> $S
ACEZ_L2_PERSIST=1 timeout 150 python bench.py --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/next_bench_l2.json 2> gpurun_out/next_bench_l2.err
stamp "bench L2 persistence rc=$?"; cut -c1-260 gpurun_out/next_bench_l2.json >> $S

# 4. everything that passed, together: bench line
if [ $rc_v3 -eq 0 ]; then
  ACEZ_CHAIN_V3=1 ACEZ_TAIL_OCC2=1 timeout 150 python bench.py --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/next_bench_v3_occ2.json 2> gpurun_out/next_bench_v3_occ2.err
  stamp "bench V3 + occ2 rc=$?"; cut -c1-400 gpurun_out/next_bench_v3_occ2.json >> $S
fi
timeout 150 python bench.py --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/next_bench_default.json 2> gpurun_out/next_bench_default.err
stamp "bench default rc=$?"; cut -c1-400 gpurun_out/next_bench_default.json >> $S
stamp done
cat $S
# Data parallel (costs 2x): gpurun --gpus 2 --timeout 600 -- 'for g in "0 0" "1 0" "1 1"; do set -- $g; ACEZ_DP_ONE_GRAPH=$1 ACEZ_DP_FUSED_FLAG=$2 \
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 300 \
#   --warmup 5 > gpurun_out/next_dp2_graph$1_flag$2.json; done; python tools/check_dp.py'   (tools/check_dp.py: 1-GPU vs 2-GPU trajectories)
