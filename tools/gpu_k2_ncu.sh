#!/bin/bash
# ncu captures of the other two conv kernel shapes of the final build: streamed-weight halo (32+64->48 3x3) and the tap-refetch
# kernel running the four sub-pixel phases of Refine 192->48 in one launch
mkdir -p gpurun_out
MONOREC_B200_CONV=f16 timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 6 -c 6 -f -o gpurun_out/prof_k2_more python tools/profile_conv.py > gpurun_out/ncu_k2_more.log 2>&1
ncu -i gpurun_out/prof_k2_more.ncu-rep --page details > gpurun_out/prof_k2_more_details.txt 2>&1
grep -E "conv_tc|Duration|DRAM Throughput|L2 Cache Throughput|highest-utilized|Issue Slots Busy|Registers Per|Dynamic Shared" gpurun_out/prof_k2_more_details.txt | cut -c1-170 | head -60
