#!/bin/bash
# Can the bonded chain co-run with a SpMV that keeps HBM busy from fewer resident warps?  (RXB_SPMV_SMEM = dummy dynamic
# shared memory per CTA -> caps the resident SpMV CTAs per SM; RXB_SPMV_DEEP = loads in flight per lane; RXB_CHAIN_MODE 0 =
# chain on a high-priority stream, 2 = low priority)
run() { env "$@" python tests/gpu_perf_probe.py 8 20 0.625 0 2>&1 | grep probe; }
run RXB_X=base
run RXB_SPMV_DEEP=16 RXB_SPMV_SMEM=50000 RXB_CHAIN_MODE=2
run RXB_SPMV_DEEP=16 RXB_SPMV_SMEM=50000 RXB_CHAIN_MODE=0
run RXB_SPMV_DEEP=16 RXB_SPMV_SMEM=40000 RXB_CHAIN_MODE=0
run RXB_SPMV_DEEP=8 RXB_SPMV_SMEM=40000 RXB_CHAIN_MODE=0
run RXB_SPMV_DEEP=8 RXB_SPMV_SMEM=40000 RXB_CHAIN_MODE=2
run RXB_SPMV_DEEP=16 RXB_CHAIN_MODE=0
