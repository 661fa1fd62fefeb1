#!/bin/bash
mkdir -p gpurun_out
RT_B200_LIB=$PWD/raytracer-ceng477-graphics-hw-1_b200/ab/reload.so timeout 200 python -m pytest tests -m gpu -x -q -p no:cacheprovider --deselect tests/test_gpu_parity.py::test_bounds_checked_build > gpurun_out/r2s_pytest_reload.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2s_pytest_reload.log; tail -3 gpurun_out/r2s_pytest_reload.log
timeout 100 python tools/kernel_ab.py --variants product,reload --json gpurun_out/r2s_kernel_ab.json horse_and_mug:3840:1920:16 marbles:2048:2048:4 mirror_spheres:2048:2048:2 car:2048:1536:8 > gpurun_out/r2s_kernel_ab.txt 2>&1; cat gpurun_out/r2s_kernel_ab.txt
