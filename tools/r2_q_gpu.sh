#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2r_pytest.log; tail -4 gpurun_out/r2r_pytest.log
timeout 600 python tools/kernel_ab.py --variants product,sc0,pe0 --json gpurun_out/r2r_kernel_ab.json > gpurun_out/r2r_kernel_ab.txt 2>&1; cat gpurun_out/r2r_kernel_ab.txt
