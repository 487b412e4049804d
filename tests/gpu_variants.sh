#!/bin/bash
# Development aid: A/B timing of kernel variants on one B200 (per-phase event timers of tests/gpu_perf_probe.py).
OUT=gpurun_out
mkdir -p $OUT
LIB=sw_reaxff_b200/librxb200.so
cp $LIB /tmp/libA.so
probe() { # name env...
  local name=$1; shift
  env "$@" timeout 120 python tests/gpu_perf_probe.py 8 20 > $OUT/probe_$name.txt 2>&1
  echo "== $name $*: $(grep 'device ms/step' $OUT/probe_$name.txt) | $(grep -E '^\s+(nonbonded|qeq_farH|bond_list|bond_orders|bonded|angle_torsion_items|hbond_items|multi_body|dbond|enum)\s' $OUT/probe_$name.txt | awk '{printf "%s=%s ", $1, $2}')"
  grep thermo $OUT/probe_$name.txt
}
for v in 0 1 2 3 4 5 6 7; do probe nb$v RXB_NB_VARIANT=$v; done
probe farh1 RXB_FARH_VARIANT=1
for v in 2 4 6; do RXB_NB_VARIANT=$v timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1; done
# everything fast: library B (RXB_FAST_BONDED) + nonbonded variant 3 + far_H variant 1
cp sw_reaxff_b200/_exp/librxb200_fb.so $LIB
probe fb RXB_NB_VARIANT=0
probe allfast RXB_NB_VARIANT=3 RXB_FARH_VARIANT=1
RXB_NB_VARIANT=3 RXB_FARH_VARIANT=1 timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -15 > $OUT/pytest_allfast.txt
cat $OUT/pytest_allfast.txt
cp /tmp/libA.so $LIB
