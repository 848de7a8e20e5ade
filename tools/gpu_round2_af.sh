#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -q -m gpu --no-header -x -k "decode or nms or golden or forward_async or full_resolution or topk" 2>&1 | tail -5 > gpurun_out/r2af_tests.log
timeout 300 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/r2af_bench.json 2> gpurun_out/r2af_bench.err
ONLY=decode timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2af_decode_launches.csv python tools/profile_kernels.py > gpurun_out/r2af_ncu_decode.log 2>&1
