#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -q -m gpu --no-header -s -x 2>&1 | tail -80 > gpurun_out/r2d_train_tests.log
timeout 300 python -m pytest tests/test_gpu_input.py -q -m gpu --no-header 2>&1 | tail -30 > gpurun_out/r2d_input_tests.log
timeout 400 python bench.py --train --graph 1 --steps 10 --warmup 3 > gpurun_out/r2d_train_graph.json 2> gpurun_out/r2d_train_graph.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2d_train_launches.csv python tools/profile_train_step.py > gpurun_out/r2d_train_prof.log 2>&1
