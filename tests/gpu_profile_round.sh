#!/bin/bash
# Evidence run for profiles/: bench (both arms), ncu launch list of the bench command, ncu --set full of every hot kernel.
# Usage (GPU box): bash tests/gpu_profile_round.sh <tag>      e.g. r02 ; then here: python tests/summarise_profiles.py <tag>
TAG=${1:-r02x}
OUT=gpurun_out
mkdir -p $OUT
python bench.py --steps 100 --warmup 10 > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_${TAG}_reference.json 2> $OUT/bench_${TAG}_reference.err
# launch list of the bench command's timed region (bench.py brackets it with cudaProfilerStart/Stop when RXB_NCU_RANGE is set)
RXB_NCU_RANGE=1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/launches_${TAG}.csv python bench.py --steps 10 --warmup 10 --quick --no-cpu-baseline --no-parity > $OUT/launches_${TAG}.log 2>&1
# ncu --set full: (1) every non-CG kernel of steady-state steps incl. a reneighbouring step, (2) the two CG kernels.
# The reports stay on the box (a full-set report with sources is ~2 MB per launch); their raw pages come back as CSV.
ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"k_far_H|k_bond_list|k_enum|k_nonbonded|k_torsion_items|k_angle_items|k_hbond_items|k_build|k_dbond|k_multi|k_bond_orders|k_shadow" \
    -c 26 -f -o /tmp/ncu_${TAG}_step python tests/gpu_ncu_target.py 8 5 > $OUT/ncu_${TAG}_step.log 2>&1
ncu -i /tmp/ncu_${TAG}_step.ncu-rep --page raw --csv > $OUT/ncu_${TAG}_step.raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_spmv2|k_cg_sweep" \
    -c 6 -f -o /tmp/ncu_${TAG}_cg python tests/gpu_ncu_target.py 8 1 > $OUT/ncu_${TAG}_cg.log 2>&1
ncu -i /tmp/ncu_${TAG}_cg.ncu-rep --page raw --csv > $OUT/ncu_${TAG}_cg.raw.csv 2>/dev/null
ls -la $OUT | tail -12
