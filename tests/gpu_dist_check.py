"""Decomposition-invariance check (run under torchrun on N GPUs of one box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/gpu_dist_check.py [nx ny nz steps T p2p]

The N-rank domain-decomposed run and a single-GPU run of the same global system (rank 0) must agree on positions,
forces, charges and energies by atom tag; the comparison itself is sw_reaxff_b200.dist.parity_check,
the same function bench.py runs before timing anything at N > 1.  Exits non-zero on mismatch.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import helpers as H  # noqa: E402
from sw_reaxff_b200 import dist as D  # noqa: E402


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
    cells = tuple(int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (4, 2, 2)
    steps = int(sys.argv[4]) if len(sys.argv) >= 5 else 12
    T = float(sys.argv[5]) if len(sys.argv) >= 6 else 300.0
    p2p = int(sys.argv[6]) if len(sys.argv) >= 7 else 1
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    out = D.parity_check(H, rank, world, lr, cells=cells, steps=steps, T=T, p2p=p2p)
    if rank == 0:
        print("dist check", json.dumps(out))
        print("DIST CHECK", "PASSED" if out["ok"] else "FAILED")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if out["ok"] else 1)


if __name__ == "__main__":
    main()
