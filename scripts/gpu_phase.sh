#!/bin/bash
mkdir -p gpurun_out
L4D_LIB_PATH=$PWD/lidar4d_b200/csrc/liblidar4d_b200_dbg.so timeout 600 python scripts/phase_clocks.py 16 2>&1 | tee gpurun_out/phase_clocks.log | grep "\[clk\]"
