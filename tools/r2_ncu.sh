#!/bin/bash
mkdir -p gpurun_out
# (1) 1920x960 16x: comparable with round 1's capture; (2) the full 8K 16x launch: DRAM traffic of the bench workload
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 2 -c 1 -o gpurun_out/r2_render_1920 -f python tools/ncu_small_frame.py horse_and_mug 16 1920 960 > gpurun_out/r2_ncu_1920.log 2>&1; echo "ncu1 rc=$?"
ncu --set full --clock-control none -k regex:render_kernel -s 1 -c 1 -o gpurun_out/r2_render_8k -f python tools/ncu_small_frame.py horse_and_mug 16 7680 3840 > gpurun_out/r2_ncu_8k.log 2>&1; echo "ncu2 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 3 -c 1 -o gpurun_out/r2_render_config3 -f python tools/ncu_small_frame.py horse_and_mug 1 > gpurun_out/r2_ncu_c3.log 2>&1; echo "ncu3 rc=$?"
ls -la gpurun_out/*.ncu-rep
