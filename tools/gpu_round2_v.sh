#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --no-header -x 2>&1 | tail -8 > gpurun_out/r2v_gputests.log
MF_NO_PATCH=1 timeout 300 python bench.py --no-cpu-baseline --steps 20 --dump-launches gpurun_out/r2v_launches_strict.json > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err
timeout 300 python bench.py --no-cpu-baseline --steps 20 --precision strict > gpurun_out/r2v_bench_patch.json 2> gpurun_out/r2v_bench_patch.err
MF_NO_PATCH=1 MF_HEAD2_BN128=1 timeout 300 python bench.py --no-cpu-baseline --steps 20 --precision strict > gpurun_out/r2v_bench_bn128.json 2> gpurun_out/r2v_bench_bn128.err
timeout 300 python bench.py --no-cpu-baseline --steps 20 --precision fast > gpurun_out/r2v_bench_fast.json 2> gpurun_out/r2v_bench_fast.err
timeout 400 python bench.py --train --steps 10 --warmup 3 > gpurun_out/r2v_train.json 2> gpurun_out/r2v_train.err
