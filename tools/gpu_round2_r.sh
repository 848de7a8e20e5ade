#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --no-header 2>&1 | tail -12 > gpurun_out/r2r_gputests.log
timeout 400 python bench.py --train --steps 10 --warmup 3 > gpurun_out/r2r_train.json 2> gpurun_out/r2r_train.err
timeout 300 python bench.py --no-cpu-baseline --steps 20 --dump-launches gpurun_out/r2r_launches_strict.json > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2r_smoke.log 2>&1
