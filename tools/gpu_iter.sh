#!/bin/bash
# quick iteration: tensor-core op tests + small model test + bench with launch breakdown
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu --no-header -x -k "tensor_core or stem or concat or maxpool" 2>&1 | tail -15
timeout 300 python -m pytest tests/test_gpu_model.py -q -m gpu --no-header -x -k "golden or levels" 2>&1 | tail -8
ROUND=${ROUND:-r01} BENCH_ARGS="--no-cpu-baseline ${BENCH_ARGS}" bash tools/gpu_bench.sh 2>&1 | cut -c1-330
