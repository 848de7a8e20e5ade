#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do timeout 400 python -m pytest tests/test_gpu_train.py -q -m gpu --no-header -x -k "full_backward_tape" 2>&1 | tail -30 > gpurun_out/r2ae_tape_$i.log; done
