#!/bin/bash
set +e
mkdir -p gpurun_out
S=gpurun_out/shot5_summary.txt
: > $S
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" >> $S; }
timeout 100 ncu --set full --clock-control none --import-source on -k regex:"head_chain" -s 4 -c 2 -o gpurun_out/chain_v2_full -f \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full5.log 2>&1
stamp "ncu full (chain v2) rc=$?"
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke5.log 2>&1
stamp "smoke rc=$?"; tail -n 2 gpurun_out/smoke5.log >> $S
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/suite5.log 2>&1
stamp "gpu suite (default path) rc=$?"; tail -n 3 gpurun_out/suite5.log >> $S
stamp done
cat $S
