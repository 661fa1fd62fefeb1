#!/bin/bash
# final single-GPU evidence of the round: tests, bench, reference arm, launch list, CLI walls, scene creation, ncu captures
bash tools/r2_final_n1.sh ${1:-r2z}
TAG=${1:-r2z} bash tools/r2z_ncu.sh
