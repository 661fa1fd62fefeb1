#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 2 -c 1 -o gpurun_out/r2p_render_1920 -f python tools/ncu_small_frame.py horse_and_mug 16 1920 960 > gpurun_out/r2p_ncu_1920.log 2>&1; echo "ncu1 rc=$?"
ls -la gpurun_out/r2p_*.ncu-rep
