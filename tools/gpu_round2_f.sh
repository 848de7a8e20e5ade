#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/grad_fidelity.py 128 4096 65536 > gpurun_out/r2f_grad_fidelity.log 2>&1
timeout 300 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
