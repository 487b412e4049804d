#!/bin/bash
# ncu --set full of selected kernels of steady-state steps; raw + source pages come back as CSV (reports stay on the box).
# Usage (GPU box): bash tests/gpu_ncu_kernels.sh <tag> "<kernel regex>" <launches> [steps]
TAG=${1:-r02x}; RE=${2:-k_far_H|k_nonbonded|k_spmv2}; CNT=${3:-6}; STEPS=${4:-2}
OUT=gpurun_out; mkdir -p $OUT
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$RE" -c $CNT -f \
    -o /tmp/ncu_${TAG} python tests/gpu_ncu_target.py 8 $STEPS > $OUT/ncu_${TAG}.log 2>&1
ncu -i /tmp/ncu_${TAG}.ncu-rep --page raw --csv > $OUT/ncu_${TAG}.raw.csv 2>/dev/null
ncu -i /tmp/ncu_${TAG}.ncu-rep --page source --csv > $OUT/ncu_${TAG}.source.csv 2>/dev/null
tail -3 $OUT/ncu_${TAG}.log; ls -la $OUT/ncu_${TAG}.*
