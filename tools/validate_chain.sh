#!/bin/bash
# One-shot GPU validation of the fused layer-chain kernels (head_chain.cu); everything lands in gpurun_out/.
#   gpurun --timeout 660 -- 'bash tools/validate_chain.sh'
# Ordered by value: the call may be cut short by the GPU budget.
set +e
mkdir -p gpurun_out
S=gpurun_out/chain_summary.txt
: > $S
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" >> $S; }

nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader >> $S 2>&1

# 0. timings + in-kernel timeline of the chain kernels (bulk-copy exchange; release/acquire vs relaxed handshake)
timeout 150 python tools/probe_chain_time.py > gpurun_out/chain_probe.log 2>&1
stamp "chain probe rc=$?"
cat gpurun_out/chain_probe.log >> $S

# 1. where does the chain differ from the per-layer path (prints per layer / per box)?
timeout 150 python tools/diag_chain.py 384 > gpurun_out/chain_diag.log 2>&1
stamp "diag (bulk DSMEM copies) rc=$?"
tail -n 4 gpurun_out/chain_diag.log >> $S

# 2. parity tests of the chain (fallback exchange variant if the bulk-copy variant fails)
XCHG=bulk
ACEZ_TEST_CHAIN=1 timeout 300 python -m pytest tests/test_head_chain_gpu.py -m gpu -q -x > gpurun_out/chain_tests.log 2>&1
rc_tests=$?
stamp "chain tests (bulk) rc=$rc_tests"
tail -n 15 gpurun_out/chain_tests.log >> $S
if [ $rc_tests -ne 0 ]; then
  XCHG=st
  ACEZ_CHAIN_XCHG=st timeout 100 python tools/diag_chain.py 384 > gpurun_out/chain_diag_st.log 2>&1
  stamp "diag (st.shared::cluster exchange) rc=$?"
  tail -n 40 gpurun_out/chain_diag_st.log >> $S
  ACEZ_CHAIN_XCHG=st ACEZ_TEST_CHAIN=1 timeout 300 python -m pytest tests/test_head_chain_gpu.py -m gpu -q -x > gpurun_out/chain_tests_st.log 2>&1
  rc_tests=$?
  stamp "chain tests (st) rc=$rc_tests"
  tail -n 15 gpurun_out/chain_tests_st.log >> $S
fi
export ACEZ_CHAIN_XCHG=$XCHG

# 3. segment timings, per-layer path vs chain
ACEZ_HEAD_CHAIN=1 timeout 120 python tools/probe_step_breakdown.py > gpurun_out/breakdown_chain.log 2>&1
stamp "breakdown chain ($XCHG) rc=$?"; cat gpurun_out/breakdown_chain.log >> $S
ACEZ_HEAD_CHAIN=0 timeout 120 python tools/probe_step_breakdown.py > gpurun_out/breakdown_layer.log 2>&1
stamp "breakdown layer rc=$?"; cat gpurun_out/breakdown_layer.log >> $S

if [ $rc_tests -eq 0 ]; then
  # 4. the bench line with the chain on
  ACEZ_HEAD_CHAIN=1 timeout 240 python bench.py --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/bench_chain.json 2> gpurun_out/bench_chain.err
  stamp "bench chain rc=$?"; cat gpurun_out/bench_chain.json >> $S
  # 5. the whole GPU suite with the chain as the plan's path
  ACEZ_HEAD_CHAIN=1 ACEZ_TEST_CHAIN=1 timeout 420 python -m pytest tests -m gpu -q -x > gpurun_out/suite_chain.log 2>&1
  stamp "full suite (chain on) rc=$?"; tail -n 8 gpurun_out/suite_chain.log >> $S
  # 6. launch list of a chain-enabled step
  ACEZ_HEAD_CHAIN=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none \
      -k regex:"gemm_tcgen05|head_chain|head_tail|fc3_|adamw|gather_rows" -s 32 -c 24 --csv \
      --log-file gpurun_out/launches_chain.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  stamp "ncu launch list rc=$?"
fi

# tail kernel grid probe (rows per warp): 4 CTAs per SM instead of 2
ACEZ_TAIL_BLOCKS_PER_SM=4 ACEZ_HEAD_CHAIN=0 timeout 100 python tools/probe_step_breakdown.py > gpurun_out/breakdown_tail4.log 2>&1
stamp "breakdown tail blocks/SM=4 rc=$?"; cat gpurun_out/breakdown_tail4.log >> $S

if [ $rc_tests -eq 0 ]; then
  # 7. full ncu sets of the kernels round 2 works on: both chains, batched wgrad, tail, AdamW (one launch each)
  ACEZ_HEAD_CHAIN=1 timeout 240 ncu --set full --clock-control none --import-source on \
      -k regex:"head_chain|gemm_tcgen05|head_tail|adamw" -s 20 -c 5 -o gpurun_out/chain_full -f \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
  stamp "ncu full rc=$?"
fi
stamp "done"
cat $S
