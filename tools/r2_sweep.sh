#!/bin/bash
mkdir -p gpurun_out
V='product,product@{"ploc_leaf_cost":0.7},product@{"ploc_leaf_cost":0.85},product@{"ploc_leaf_cost":1.3},product@{"ploc_radius":32},product@{"ploc_radius":8},product@{"builder":5}'
timeout 900 python tools/kernel_ab.py --variants "$V" --json gpurun_out/r2c_sweep.json horse_and_mug:3840:1920:16 horse_and_mug:1920:960:16 horse_and_mug:1440:720:1 car:2048:1536:8 > gpurun_out/r2c_sweep.log 2>&1; echo "rc=$?"
cat gpurun_out/r2c_sweep.log
