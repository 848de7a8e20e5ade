#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -q -m gpu --no-header -x -s -k "reference_training_loop or end_to_end_train or fused_adamw" 2>&1 | tail -40 > gpurun_out/r2_train_tests.log
timeout 300 python bench.py --train --graph 0 --steps 5 --warmup 3 > gpurun_out/r2_train_eager.json 2> gpurun_out/r2_train_eager.err
timeout 400 python bench.py --train --graph 1 --steps 10 --warmup 3 > gpurun_out/r2_train_graph.json 2> gpurun_out/r2_train_graph.err
