#!/bin/bash
# round 2, insertion-based tree optimisation: GPU tests, A/B table, bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2n_pytest.log; tail -5 gpurun_out/r2n_pytest.log
timeout 600 python tools/reinsert_ab.py > gpurun_out/r2n_reinsert_ab.txt 2> gpurun_out/r2n_reinsert_ab.err; tail -60 gpurun_out/r2n_reinsert_ab.txt
timeout 600 python bench.py > gpurun_out/r2n_bench_n1.json 2> gpurun_out/r2n_bench_n1.err; cut -c1-900 gpurun_out/r2n_bench_n1.json
