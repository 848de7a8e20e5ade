#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_ops.py -q -m gpu --no-header -x 2>&1 | tail -12 > gpurun_out/r2t_gputests.log
timeout 300 python bench.py --no-cpu-baseline --steps 20 --dump-launches gpurun_out/r2t_launches_strict.json > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err
MF_HEAD2_BN128=1 timeout 300 python bench.py --no-cpu-baseline --steps 20 --precision strict > gpurun_out/r2t_bench_bn128.json 2> gpurun_out/r2t_bench_bn128.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2t_smoke.log 2>&1
