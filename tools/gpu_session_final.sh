#!/bin/bash
# Final validation of the tree: full GPU suite, smoke(), bench line (one GPU).
set +e
mkdir -p gpurun_out
S=gpurun_out/final_summary.txt
: > $S
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" >> $S; }
timeout 400 python -m pytest tests -x -q -m gpu > gpurun_out/final_suite.log 2>&1
stamp "full GPU suite rc=$?"; tail -n 6 gpurun_out/final_suite.log | cut -c1-400 >> $S
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
stamp "smoke rc=$?"; tail -n 2 gpurun_out/final_smoke.log | cut -c1-300 >> $S
timeout 400 python bench.py --steps 300 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
stamp "bench rc=$?"; cut -c1-600 gpurun_out/final_bench.json >> $S; tail -n 3 gpurun_out/final_bench.err | cut -c1-300 >> $S
stamp done
cat $S
