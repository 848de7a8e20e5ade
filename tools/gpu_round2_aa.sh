#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -q -m gpu --no-header -x 2>&1 | tail -6 > gpurun_out/r2aa_traintests.log
timeout 400 python bench.py --train --steps 10 --warmup 3 > gpurun_out/r2aa_train.json 2> gpurun_out/r2aa_train.err
