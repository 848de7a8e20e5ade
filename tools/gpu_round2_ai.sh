#!/bin/bash
mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/r2ai_bench.json 2> gpurun_out/r2ai_bench.err
timeout 400 python bench.py --train --steps 10 --warmup 3 > gpurun_out/r2ai_train.json 2> gpurun_out/r2ai_train.err
