#!/bin/bash
# Round deliverables: full GPU test suite, smoke(), bench (+reference arm), ncu launch list, ncu full capture of the top kernels.
mkdir -p gpurun_out
R=${ROUND:-r01}
timeout 900 python -m pytest tests/ -q -m gpu --no-header 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_$R.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke_$R.log
NCU_LIST=1 NCU_FULL=1 ONLY=head,dcn64,dcn128,conv128,stem,offconv64,decode ROUND=$R bash tools/gpu_bench.sh 2>&1 | cut -c1-400
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$R.json 2>> gpurun_out/bench_$R.err; cat gpurun_out/bench_ref_$R.json | cut -c1-300
