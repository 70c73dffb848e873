#!/bin/bash
# Multi-GPU session: bash tools/gpu_session_dp.sh N [full]   (run under gpurun --gpus N)
#   full: also the engine-level / product-level 1-rank vs N-rank trajectory checks and the NCCL-fallback bench
set +e
N=${1:-2}
MODE=${2:-bench}
mkdir -p gpurun_out
S=gpurun_out/dp${N}_summary.txt
: > $S
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" >> $S; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/dp${N}_topo.txt 2>&1
if [ "$MODE" = "full" ] || [ "$MODE" = "check" ]; then
  timeout 300 $TR --master-port 29501 tools/check_dp.py > gpurun_out/dp${N}_check.log 2>&1
  stamp "check_dp (engine level) rc=$?"; grep -E "^\[dp|RESULT|Error" gpurun_out/dp${N}_check.log | cut -c1-250 >> $S
fi
if [ "$MODE" = "full" ]; then
  timeout 600 python tools/check_dp_product.py $N > gpurun_out/dp${N}_product.log 2>&1
  stamp "check_dp_product rc=$?"; tail -n 12 gpurun_out/dp${N}_product.log | cut -c1-300 >> $S
fi
timeout 400 $TR --master-port 29502 bench.py --gpus $N --steps 300 --warmup 5 > gpurun_out/dp${N}_bench.json 2> gpurun_out/dp${N}_bench.err
stamp "bench N=$N (peer-memory optimiser) rc=$?"; cut -c1-900 gpurun_out/dp${N}_bench.json >> $S; tail -n 3 gpurun_out/dp${N}_bench.err | cut -c1-300 >> $S
if [ "$MODE" = "check" ]; then
  ACEZ_DP_MULTICAST=0 timeout 400 $TR --master-port 29503 bench.py --gpus $N --steps 300 --warmup 5 > gpurun_out/dp${N}_bench_p2p.json 2> gpurun_out/dp${N}_bench_p2p.err
  stamp "bench N=$N (ACEZ_DP_MULTICAST=0: peer loads / stores instead of the switch reduction) rc=$?"; cut -c1-420 gpurun_out/dp${N}_bench_p2p.json >> $S
fi
if [ "$MODE" = "full" ]; then
  ACEZ_DP_PEERS=0 timeout 400 $TR --master-port 29504 bench.py --gpus $N --steps 300 --warmup 5 > gpurun_out/dp${N}_bench_nccl.json 2> gpurun_out/dp${N}_bench_nccl.err
  stamp "bench N=$N (NCCL all-reduce path) rc=$?"; cut -c1-900 gpurun_out/dp${N}_bench_nccl.json >> $S; tail -n 3 gpurun_out/dp${N}_bench_nccl.err | cut -c1-300 >> $S
fi
[ "$MODE" = "full" ] && timeout 200 $TR --master-port 29505 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/dp${N}_bench_ref.json 2> gpurun_out/dp${N}_bench_ref.err
stamp "bench --impl reference N=$N rc=$?"; cut -c1-600 gpurun_out/dp${N}_bench_ref.json >> $S
stamp done
cat $S
