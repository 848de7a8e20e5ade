#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_train.py -q -m gpu --no-header -x -k "captured_train_step" 2>&1 | tail -40 > gpurun_out/r2ac_cap_$i.log; done
