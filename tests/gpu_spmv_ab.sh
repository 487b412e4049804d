#!/bin/bash
# A/B of SpMV forms: RXB_SPMV_DEEP=U (loads in flight per lane), RXB_SPMV_SMEM (dummy shared memory = occupancy cap)
for cfg in "RXB_SPMV_DEEP=0 RXB_SPMV_SMEM=0" "RXB_SPMV_DEEP=8 RXB_SPMV_SMEM=0" "RXB_SPMV_DEEP=16 RXB_SPMV_SMEM=0" \
           "RXB_SPMV_DEEP=8 RXB_SPMV_SMEM=57000" "RXB_SPMV_DEEP=16 RXB_SPMV_SMEM=57000" "RXB_SPMV_DEEP=16 RXB_SPMV_SMEM=75000" \
           "RXB_SPMV_DEEP=16 RXB_SPMV_SMEM=113000"; do
  env $cfg python tests/gpu_perf_probe.py 8 20 0.625 1 2>&1 | tail -1 | sed 's/neigh=.*qeq_cg/qeq_cg/; s/bond_list.*spmv=/spmv=/; s/hbond.*//'
done
