#!/bin/bash
# Validation of chain v2 (register-resident residual stream, bit masks) + the data needed to decide the default path.
set +e
mkdir -p gpurun_out
S=gpurun_out/shot4_summary.txt
: > $S
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" >> $S; }
ACEZ_TEST_CHAIN=1 timeout 120 python -m pytest tests/test_head_chain_gpu.py -m gpu -q -x > gpurun_out/chain_tests4.log 2>&1
rc=$?
stamp "chain tests rc=$rc"; tail -n 6 gpurun_out/chain_tests4.log >> $S
ACEZ_PROBE_COMBOS="1:0,1:16,1:62" timeout 100 python tools/probe_chain_time.py > gpurun_out/chain_probe4.log 2>&1
stamp "probe rc=$?"; cat gpurun_out/chain_probe4.log >> $S
if [ $rc -eq 0 ]; then
  ACEZ_HEAD_CHAIN=1 ACEZ_TEST_CHAIN=1 timeout 300 python -m pytest tests -m gpu -q -x > gpurun_out/suite_chain4.log 2>&1
  stamp "full suite (chain on) rc=$?"; tail -n 4 gpurun_out/suite_chain4.log >> $S
  ACEZ_HEAD_CHAIN=1 timeout 120 python bench.py --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/bench_chain4.json 2> gpurun_out/bench_chain4.err
  stamp "bench chain rc=$?"; cat gpurun_out/bench_chain4.json >> $S
  ACEZ_HEAD_CHAIN=0 timeout 120 python bench.py --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/bench_layer4.json 2> gpurun_out/bench_layer4.err
  stamp "bench layer rc=$?"; cat gpurun_out/bench_layer4.json >> $S
  ACEZ_HEAD_CHAIN=1 timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none \
      -k regex:"gemm_tcgen05|head_chain|head_tail|fc3_|adamw|gather_rows" -s 32 -c 24 --csv \
      --log-file gpurun_out/launches_chain4.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench4.log 2>&1
  stamp "ncu launch list rc=$?"
fi
stamp done
cat $S
