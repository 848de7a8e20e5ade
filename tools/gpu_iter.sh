#!/bin/bash
# quick iteration: op tests + small model tests + bench with launch breakdown
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu --no-header -x -k "${OPS_K:-tensor_core or stem or concat or maxpool or backward or rows}" 2>&1 | tail -15
timeout 300 python -m pytest tests/test_gpu_model.py -q -m gpu --no-header -x -k "${MODEL_K:-golden or levels}" 2>&1 | tail -8
ROUND=${ROUND:-r01} BENCH_ARGS="--no-cpu-baseline ${BENCH_ARGS}" bash tools/gpu_bench.sh 2>&1 | cut -c1-330
