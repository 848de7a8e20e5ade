#!/bin/bash
mkdir -p gpurun_out
ONLY=decode timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r2ag_decode env MF_PRECISION=strict python tools/profile_kernels.py > gpurun_out/r2ag_ncu_decode.log 2>&1
