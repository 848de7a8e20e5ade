#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -q -m gpu --no-header 2>&1 | tail -12 > gpurun_out/r2ad_traintests.log
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_train.py -q -m gpu --no-header -x -k "captured_train_step" 2>&1 | tail -8 > gpurun_out/r2ad_cap_$i.log; done
timeout 400 python bench.py --train --steps 10 --warmup 3 > gpurun_out/r2ad_train.json 2> gpurun_out/r2ad_train.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2ad_train_launches.csv python tools/profile_train_step.py > gpurun_out/r2ad_train_prof.log 2>&1
