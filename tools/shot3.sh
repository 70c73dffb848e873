#!/bin/bash
set +e
mkdir -p gpurun_out
S=gpurun_out/shot3_summary.txt
: > $S
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" >> $S; }
ACEZ_TEST_CHAIN=1 timeout 200 python -m pytest tests/test_head_chain_gpu.py -m gpu -q -x > gpurun_out/chain_tests3.log 2>&1
stamp "chain tests rc=$?"; tail -n 4 gpurun_out/chain_tests3.log >> $S
timeout 200 python tools/probe_chain_time.py > gpurun_out/chain_probe3.log 2>&1
stamp "probe rc=$?"; cat gpurun_out/chain_probe3.log >> $S
stamp done
cat $S
