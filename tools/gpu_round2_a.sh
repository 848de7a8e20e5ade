#!/bin/bash
# round-2 first GPU pass: strict-precision kernels, the un-run trainer pieces, whole suite, bench in both precisions
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu --no-header -x -k "strict" 2>&1 | tail -30 > gpurun_out/r2_strict_ops.log
timeout 400 python -m pytest tests/test_gpu_model.py -q -m gpu --no-header -s 2>&1 | tail -60 > gpurun_out/r2_model.log
MF_RUN_UNVERIFIED=1 timeout 240 python -m pytest tests/test_gpu_train.py -q -m gpu --no-header -s -k "end_to_end_train_steps" 2>&1 | tail -40 > gpurun_out/r2_unverified.log
timeout 600 python -m pytest tests -q -m gpu --no-header 2>&1 | tail -40 > gpurun_out/r2_gputests.log
timeout 300 python bench.py --precision strict --dump-launches gpurun_out/r2_launches_strict.json > gpurun_out/r2_bench_strict.json 2> gpurun_out/r2_bench_strict.err
timeout 300 python bench.py --precision fast --no-cpu-baseline --dump-launches gpurun_out/r2_launches_fast.json > gpurun_out/r2_bench_fast.json 2> gpurun_out/r2_bench_fast.err
