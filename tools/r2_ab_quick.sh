#!/bin/bash
# usage: tools/r2_ab_quick.sh variants [cases...]
mkdir -p gpurun_out
V=$1; shift
timeout 900 python tools/kernel_ab.py --variants "$V" --json gpurun_out/r2_ab_quick.json "$@" > gpurun_out/r2_ab_quick.log 2>&1; echo "rc=$?"
cat gpurun_out/r2_ab_quick.log
