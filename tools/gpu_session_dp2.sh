#!/bin/bash
# 2-GPU session: engine-level and product-level data-parallel checks, bench at N = 2 (two-graph default, one-graph NCCL capture).
set +e
mkdir -p gpurun_out
S=gpurun_out/dp2_summary.txt
: > $S
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" >> $S; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29501 tools/check_dp.py > gpurun_out/dp2_check.log 2>&1
stamp "check_dp (engine level) rc=$?"; grep -E "^\[dp|RESULT|Error" gpurun_out/dp2_check.log | cut -c1-250 >> $S
timeout 900 python tools/check_dp_product.py 2 > gpurun_out/dp2_product.log 2>&1
stamp "check_dp_product rc=$?"; tail -n 12 gpurun_out/dp2_product.log | cut -c1-300 >> $S
timeout 300 $TR --master-port 29502 bench.py --gpus 2 --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/dp2_bench.json 2> gpurun_out/dp2_bench.err
stamp "bench N=2 (peer-memory optimiser, eager between the graph and the next iteration) rc=$?"; cut -c1-700 gpurun_out/dp2_bench.json >> $S; tail -n 3 gpurun_out/dp2_bench.err | cut -c1-300 >> $S
ACEZ_DP_PEERS_GRAPH=1 timeout 300 $TR --master-port 29503 bench.py --gpus 2 --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/dp2_bench_peers_graph.json 2> gpurun_out/dp2_bench_peers_graph.err
stamp "bench N=2 (peer-memory optimiser captured in ONE graph) rc=$?"; cut -c1-700 gpurun_out/dp2_bench_peers_graph.json >> $S; tail -n 3 gpurun_out/dp2_bench_peers_graph.err | cut -c1-300 >> $S
ACEZ_DP_PEERS=0 timeout 300 $TR --master-port 29504 bench.py --gpus 2 --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/dp2_bench_nccl.json 2> gpurun_out/dp2_bench_nccl.err
stamp "bench N=2 (NCCL all-reduce path) rc=$?"; cut -c1-700 gpurun_out/dp2_bench_nccl.json >> $S; tail -n 3 gpurun_out/dp2_bench_nccl.err | cut -c1-300 >> $S
stamp done
cat $S
