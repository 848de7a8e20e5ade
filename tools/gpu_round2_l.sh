#!/bin/bash
mkdir -p gpurun_out
ONLY=head,head_1x1,dcn64,dcn128,offconv64,conv128,conv64,conv512,stem,level0 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r2l_strict env MF_PRECISION=strict python tools/profile_kernels.py > gpurun_out/r2l_ncu_strict.log 2>&1
ONLY=head,dcn64,dcn128,offconv64,conv128,conv64,stem,level0 timeout 900 ncu --set full --clock-control none --profile-from-start off -f -o gpurun_out/r2l_fast env MF_PRECISION=fast python tools/profile_kernels.py > gpurun_out/r2l_ncu_fast.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2l_infer_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2l_ncu_bench.log 2>&1
