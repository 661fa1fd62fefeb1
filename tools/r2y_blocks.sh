#!/bin/bash
mkdir -p gpurun_out
python tools/partition_experiment.py 8 7680 3840 16 blocks > gpurun_out/r2y_partition_blocks.txt 2>&1; cat gpurun_out/r2y_partition_blocks.txt
python -m pytest tests -m gpu -x -q -k "bands or narrow or aa_factors or golden or full_size" > gpurun_out/r2y_pytest_blocks.log 2>&1; tail -3 gpurun_out/r2y_pytest_blocks.log
