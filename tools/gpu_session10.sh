#!/bin/bash
# Round-2 session 10: in-kernel timeline of both chain kernels (default own-first order and arrival order), ncu --set full of
# the seven kernels of one training iteration (eager launches).
set +e
mkdir -p gpurun_out
S=gpurun_out/s10_summary.txt
: > $S
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" >> $S; }
ACEZ_PROBE_COMBOS="1:0" timeout 120 python tools/probe_chain_time.py > gpurun_out/s10_probe_own.log 2>&1
stamp "chain probe (own-first) rc=$?"; cat gpurun_out/s10_probe_own.log >> $S
ACEZ_CHAIN_ORDER=arrival ACEZ_PROBE_COMBOS="1:0" timeout 120 python tools/probe_chain_time.py > gpurun_out/s10_probe_arr.log 2>&1
stamp "chain probe (arrival order) rc=$?"; cat gpurun_out/s10_probe_arr.log >> $S
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'head_chain4|gemm2cta|head_tail|adamw_kernel|fc3_reduce|gather_rows' --launch-skip 35 --launch-count 7 -o gpurun_out/s10_step -f python tools/ncu_step.py > gpurun_out/s10_ncu_step.log 2>&1
stamp "ncu step kernels rc=$?"; tail -n 3 gpurun_out/s10_ncu_step.log | cut -c1-300 >> $S
stamp done
cat $S
