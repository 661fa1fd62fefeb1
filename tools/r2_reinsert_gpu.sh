#!/bin/bash
# round 2, insertion-based tree optimisation + zero-specular skip: GPU tests, A/B tables, bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2o_pytest.log; tail -5 gpurun_out/r2o_pytest.log
timeout 600 python tools/kernel_ab.py --variants product,spec_all --json gpurun_out/r2o_kernel_ab.json > gpurun_out/r2o_kernel_ab.txt 2>&1; cat gpurun_out/r2o_kernel_ab.txt
timeout 600 python bench.py > gpurun_out/r2o_bench_n1.json 2> gpurun_out/r2o_bench_n1.err; cut -c1-400 gpurun_out/r2o_bench_n1.json
