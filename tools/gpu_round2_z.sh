#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu --no-header -x -k "patch or strict" 2>&1 | tail -5 > gpurun_out/r2z_ops.log
MF_PATCH=1 timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu --no-header -x 2>&1 | tail -5 > gpurun_out/r2z_model_patch.log
MF_PATCH=1 timeout 300 python bench.py --no-cpu-baseline --steps 20 --precision strict > gpurun_out/r2z_bench_patch.json 2> gpurun_out/r2z_bench_patch.err
timeout 300 python bench.py --no-cpu-baseline --steps 20 --precision strict > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err
