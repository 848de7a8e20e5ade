#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -q -m gpu --no-header -k "decode or nms or golden or forward_async or full_resolution" 2>&1 | tail -12 > gpurun_out/r2o_tests.log
timeout 300 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2o_infer_launches_strict.csv python tools/profile_infer_step.py > gpurun_out/r2o_prof_strict.log 2>&1
MF_PRECISION=fast timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2o_infer_launches_fast.csv python tools/profile_infer_step.py > gpurun_out/r2o_prof_fast.log 2>&1
