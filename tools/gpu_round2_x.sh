#!/bin/bash
mkdir -p gpurun_out
ONLY=head,head2_reduce,dcn64,dcn128,offconv64,conv128,conv64,conv512,stem,level0 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r2x_strict env MF_PRECISION=strict python tools/profile_kernels.py > gpurun_out/r2x_ncu_strict.log 2>&1
ONLY=head,dcn64,dcn128,offconv64,conv128,conv64,stem,level0 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r2x_fast env MF_PRECISION=fast python tools/profile_kernels.py > gpurun_out/r2x_ncu_fast.log 2>&1
