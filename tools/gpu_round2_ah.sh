#!/bin/bash
mkdir -p gpurun_out
MF_PDL=1 timeout 300 python bench.py --no-cpu-baseline --steps 30 > gpurun_out/r2ah_bench_pdl.json 2> gpurun_out/r2ah_bench_pdl.err
timeout 300 python bench.py --no-cpu-baseline --steps 30 > gpurun_out/r2ah_bench.json 2> gpurun_out/r2ah_bench.err
