#!/bin/bash
# 8-GPU evidence of the final build: inference replicas and the DDP train step (SyncBatchNorm on, then per-GPU BN)
N=${1:-8}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus $N --steps 20 --warmup 4 --no-cpu-baseline > gpurun_out/fin_infer_${N}gpu.json 2> gpurun_out/fin_infer_${N}gpu.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 \
  bench.py --train --gpus $N --steps 10 --warmup 3 > gpurun_out/fin_train_${N}gpu_syncbn.json 2> gpurun_out/fin_train_${N}gpu_syncbn.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 \
  bench.py --train --sync-bn 0 --gpus $N --steps 10 --warmup 3 > gpurun_out/fin_train_${N}gpu_localbn.json 2> gpurun_out/fin_train_${N}gpu_localbn.err
tail -c 600 gpurun_out/fin_infer_${N}gpu.json; tail -c 400 gpurun_out/fin_train_${N}gpu_syncbn.json
