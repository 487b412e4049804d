#!/bin/bash
# Evidence run for profiles/: bench (both arms), ncu launch list of the bench command, ncu --set full of the top kernels.
# Usage (GPU box): bash tests/gpu_profile_round.sh <tag>      e.g. r01c
TAG=${1:-r01x}
OUT=gpurun_out
mkdir -p $OUT
python bench.py --steps 100 --warmup 10 > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_${TAG}_reference.json 2> $OUT/bench_${TAG}_reference.err
# launch list of the bench command's timed region (bench.py brackets it with cudaProfilerStart/Stop when RXB_NCU_RANGE is set)
RXB_NCU_RANGE=1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/launches_${TAG}.csv python bench.py --steps 10 --warmup 10 --no-cpu-baseline > $OUT/launches_${TAG}.log 2>&1
for K in k_spmv2 k_far_H k_nonbonded k_bond_list k_torsion_items k_cg_sweep; do
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$K -c 2 -f -o $OUT/ncu_${TAG}_$K \
      python tests/gpu_ncu_target.py 8 2 > $OUT/ncu_${TAG}_$K.log 2>&1
done
ls -la $OUT | tail -20
