#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -q -m gpu --no-header -s -k "strict" 2>&1 | tail -12 > gpurun_out/r2n_tests.log
timeout 300 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
