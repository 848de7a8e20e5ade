#!/bin/bash
# First thing to run on a GPU next round: the pieces written after the round-1 GPU budget ended (end-to-end train step, stem
# weight gradient). Bounded by timeouts so that a hang cannot take the box down.
mkdir -p gpurun_out
MF_RUN_UNVERIFIED=1 timeout 240 python -m pytest tests/test_gpu_train.py -q -m gpu --no-header -s -k "end_to_end_train_steps" 2>&1 \
  | tail -25 | tee gpurun_out/unverified.log
