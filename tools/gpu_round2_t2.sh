#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_ops.py -q -m gpu --no-header -x 2>&1 | tail -4 > gpurun_out/r2t2_gputests.log
timeout 300 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/r2t2_bench.json 2> gpurun_out/r2t2_bench.err
timeout 300 python bench.py --no-cpu-baseline --steps 20 --precision fast > gpurun_out/r2t2_bench_fast.json 2> gpurun_out/r2t2_bench_fast.err
