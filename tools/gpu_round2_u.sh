#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu --no-header -x -k "strict" 2>&1 | tail -15 > gpurun_out/r2u_ops.log
timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu --no-header -x 2>&1 | tail -8 > gpurun_out/r2u_model.log
timeout 300 python bench.py --no-cpu-baseline --steps 20 --dump-launches gpurun_out/r2u_launches_strict.json > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err
MF_NO_PATCH=1 timeout 300 python bench.py --no-cpu-baseline --steps 20 --precision strict > gpurun_out/r2u_bench_nopatch.json 2> gpurun_out/r2u_bench_nopatch.err
