#!/bin/bash
# round-2 starting point: cold start, scene creation, small frames (run on the GPU box via gpurun)
set -x
mkdir -p gpurun_out
nvidia-smi -L; nproc
python tools/scene_create_profile.py > gpurun_out/r2_base_create_lazy.jsonl 2> gpurun_out/r2_base_create_lazy.err
CUDA_MODULE_LOADING=EAGER python tools/scene_create_profile.py horse_and_mug > gpurun_out/r2_base_create_eager.jsonl 2>&1
python tools/cli_wall.py --runs 4 --json gpurun_out/r2_base_cli_wall.json --md gpurun_out/r2_base_cli_wall.md > gpurun_out/r2_base_cli_wall.log 2>&1
python tools/ncu_small_frame.py > gpurun_out/r2_base_small_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 3 -c 1 -o gpurun_out/r2_base_config3 -f python tools/ncu_small_frame.py > gpurun_out/r2_base_ncu_small.log 2>&1
tail -3 gpurun_out/r2_base_cli_wall.log
cat gpurun_out/r2_base_create_lazy.jsonl
