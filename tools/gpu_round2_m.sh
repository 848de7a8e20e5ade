#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -q -m gpu --no-header -s -k "strict or decode" 2>&1 | tail -30 > gpurun_out/r2m_tests.log
timeout 300 python bench.py --no-cpu-baseline --steps 20 --dump-launches gpurun_out/r2m_launches_strict.json > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
MF_SPLIT_KCONCAT=1 timeout 300 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/r2m_bench_kconcat.json 2> gpurun_out/r2m_bench_kconcat.err
timeout 300 python bench.py --no-cpu-baseline --steps 10 --batch 32 > gpurun_out/r2m_bench_b32.json 2> gpurun_out/r2m_bench_b32.err
