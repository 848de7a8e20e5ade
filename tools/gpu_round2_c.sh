#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu --no-header -x -k "strict" 2>&1 | tail -15 > gpurun_out/r2c_strict_ops.log
timeout 400 python -m pytest tests/test_gpu_model.py -q -m gpu --no-header -s -k "strict" 2>&1 | tail -30 > gpurun_out/r2c_model.log
timeout 600 python -m pytest tests/test_gpu_train.py -q -m gpu --no-header -s -k "reference_training_loop or end_to_end_train" 2>&1 | tail -30 > gpurun_out/r2c_train_tests.log
timeout 300 python bench.py --precision strict --no-cpu-baseline --dump-launches gpurun_out/r2c_launches_strict.json > gpurun_out/r2c_bench_strict.json 2> gpurun_out/r2c_bench_strict.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2c_train_launches.csv python tools/profile_train_step.py > gpurun_out/r2c_train_prof.log 2>&1
ONLY=dcn64,offconv64 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r2c_dcn_fast env MF_PRECISION=fast python tools/profile_kernels.py > gpurun_out/r2c_ncu_dcn.log 2>&1
