#!/bin/bash
mkdir -p gpurun_out
ONLY=head,offconv64,conv128,conv64 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r2y_patch env MF_PRECISION=strict MF_PATCH=1 python tools/profile_kernels.py > gpurun_out/r2y_ncu_patch.log 2>&1
