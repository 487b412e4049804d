timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 120 python tests/gpu_perf_probe.py 8 20 2>&1 | grep -E "ms/step |^\s+(neigh|bonded|dbond|enum|angle_torsion_items|hbond_items|multi_body)\s" | awk '{printf "%s=%s ", $1, $2} END {print ""}'
