#!/bin/bash
# N=2 A/B on one box: bench --quick --no-parity with the given env settings, prints ms/step and the CG phase
run() {
  env "$@" timeout 250 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29540 + RANDOM % 50)) \
    bench.py --gpus 2 --steps 20 --warmup 5 --quick --no-parity 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); k = d['kernel_ms_per_step']
print('$*', round(d['ms_per_step'], 3), 'cg', k['qeq_cg'], 'spmv', k['spmv'], 'non-spmv', round(k['qeq_cg'] - k['spmv'], 3))"
}
for v in ${AB_VALUES:-1 0 1 0}; do run ${AB_VAR:-RXB_PDL}=$v; done
