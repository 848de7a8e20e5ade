#!/bin/bash
# N-GPU evidence run (N = $1, default 8): bench.py under torchrun at N and the fused gradient-exchange check at N ranks.
N=${1:-8}
R=${ROUND:-r01}
mkdir -p gpurun_out
if [ -z "$SKIP_BENCH" ]; then
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_${N}gpu_$R.json 2> gpurun_out/bench_${N}gpu_$R.err
tail -c 1500 gpurun_out/bench_${N}gpu_$R.json
fi
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 \
  tools/p2p_adamw_check.py > gpurun_out/p2p_adamw_${N}gpu_$R.log 2>&1
grep -v "^$" gpurun_out/p2p_adamw_${N}gpu_$R.log | grep -i "world\|error\|mismatch\|differ\|Traceback\|File \|exit" | head -30 | cut -c1-600
