#!/bin/bash
# A/B of the register bounds of the angle / torsion item kernels (RXB_ANG_OCC, RXB_TOR_OCC = resident CTAs per SM the
# allocation is bounded for).  Prints the resident step and the serialised per-phase times.
mkdir -p gpurun_out
out=gpurun_out/item_occ_ab.txt
: > $out
for tor in 4 3 5 6; do
  RXB_TOR_OCC=$tor python tests/gpu_perf_probe.py 8 20 0.625 1 2>&1 | grep probe >> $out
done
for ang in 4 5; do
  RXB_ANG_OCC=$ang python tests/gpu_perf_probe.py 8 20 0.625 1 2>&1 | grep probe >> $out
done
cat $out
