#!/bin/bash
# Runs the GPU tests in stages (a CUDA fault or hang in one stage must not hide the others); logs in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { # name timeout pytest-args...
  local name=$1 t=$2; shift 2
  timeout $t python -m pytest "$@" -q -m gpu --no-header -rA -s > gpurun_out/$name.log 2>&1
  local rc=$?
  echo "== $name exit $rc"; grep -E "passed|failed|error|Error|PASSED|FAILED" gpurun_out/$name.log | tail -${TAIL:-40}
  return $rc
}
run ops_simt 600 tests/test_gpu_ops.py -k "not tensor_core and not stem and not concat"
run ops_tc 240 tests/test_gpu_ops.py -k "tensor_core or stem or concat"
if [ $? -eq 0 ]; then
  run model 1200 tests/test_gpu_model.py
else
  echo "tensor-core stage failed: running the model tests on the CUDA-core cross-check kernels"
  MF_CONV_IMPL=1 run model_simt 1200 tests/test_gpu_model.py
fi
