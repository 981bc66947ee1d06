#!/bin/bash
mkdir -p gpurun_out
python scripts/conv_profile_target.py > gpurun_out/r02d_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_conv8 -c 6 -f -o gpurun_out/r02d_conv python scripts/conv_profile_target.py > gpurun_out/r02d_ncu.log 2>&1
tail -2 gpurun_out/r02d_ncu.log
