#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/val2_gputests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/val2_smoke.log 2>&1
