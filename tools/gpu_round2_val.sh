#!/bin/bash
# driver-style validation of the committed build: full GPU tests, smoke, default bench, reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/val_gputests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/val_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/val_bench.json 2> gpurun_out/val_bench.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/val_ref.json 2> gpurun_out/val_ref.err
