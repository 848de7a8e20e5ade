#!/bin/bash
# N-GPU scaling evidence of the final build: inference replicas + DDP train step with SyncBatchNorm
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus $N --steps 20 --warmup 4 --no-cpu-baseline > gpurun_out/fin_infer_${N}gpu.json 2> gpurun_out/fin_infer_${N}gpu.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 \
  bench.py --train --gpus $N --steps 10 --warmup 3 > gpurun_out/fin_train_${N}gpu_syncbn.json 2> gpurun_out/fin_train_${N}gpu_syncbn.err
tail -c 300 gpurun_out/fin_infer_${N}gpu.json; tail -c 300 gpurun_out/fin_train_${N}gpu_syncbn.json
