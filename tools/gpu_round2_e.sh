#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --no-header 2>&1 | tail -60 > gpurun_out/r2e_gputests.log
timeout 400 python bench.py --dump-launches gpurun_out/r2e_launches_strict.json > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
timeout 400 python bench.py --train --steps 10 --warmup 3 > gpurun_out/r2e_train.json 2> gpurun_out/r2e_train.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2e_train_launches.csv python tools/profile_train_step.py > gpurun_out/r2e_train_prof.log 2>&1
timeout 200 python bench.py --impl torch_gpu --steps 5 > gpurun_out/r2e_torch_gpu.json 2> gpurun_out/r2e_torch_gpu.err
