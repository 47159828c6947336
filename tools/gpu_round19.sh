#!/bin/bash
mkdir -p gpurun_out
MONOREC_B200_TC_HALO=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_halo -s 2 -c 1 -o gpurun_out/prof_k2_halo python tools/profile_conv.py > gpurun_out/ncu_k2h.log 2>&1
tail -2 gpurun_out/ncu_k2h.log
