#!/bin/bash
mkdir -p gpurun_out
ONLY=head,head_1x1,dcn64,dcn128,offconv64,conv128,conv64,conv512,stem,level0 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r2s_strict env MF_PRECISION=strict python tools/profile_kernels.py > gpurun_out/r2s_ncu_strict.log 2>&1
ONLY=decode timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2s_decode_launches.csv python tools/profile_kernels.py > gpurun_out/r2s_ncu_decode.log 2>&1
