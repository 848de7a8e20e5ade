#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --no-header -x 2>&1 | tail -8 > gpurun_out/r2w_gputests.log
timeout 300 python bench.py --no-cpu-baseline --steps 20 --dump-launches gpurun_out/r2w_launches_strict.json > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err
timeout 300 python bench.py --no-cpu-baseline --steps 20 --precision fast --dump-launches gpurun_out/r2w_launches_fast.json > gpurun_out/r2w_bench_fast.json 2> gpurun_out/r2w_bench_fast.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2w_smoke.log 2>&1
