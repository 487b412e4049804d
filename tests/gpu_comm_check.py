"""Multi-rank host-planned halo check (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/gpu_comm_check.py

Every rank drives its own library handle through the plugin calls only (set_atoms / comm_set_ghosts / neigh_build /
set_positions / qeq_pre_force / pair_compute), the way a multi-rank LAMMPS would (tests/lammps_comm.py is the stand-in for
its Comm); the trajectory is compared with the plain single-rank plugin run of the same system, computed by every rank on
its own GPU.  Prints one JSON line on rank 0 and exits non-zero on a mismatch."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import helpers as H  # noqa: E402
import lammps_comm as LC  # noqa: E402
from sw_reaxff_b200 import Rxb  # noqa: E402


def main():
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); dev = int(os.environ.get("LOCAL_RANK", 0))
    grids = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}
    cells = {1: (2, 2, 2), 2: (4, 2, 2), 4: (4, 4, 2), 8: (4, 4, 4)}[world]
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    if world > 1:
        dist.init_process_group("gloo")

    def ag(o):
        if world == 1:
            return [o]
        out = [None] * world
        dist.all_gather_object(out, o)
        return out
    uid = [Rxb.dist_unique_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(uid, src=0)
    box, x, t, tag = H.tatb_cell(*cells)
    v = H.maxwell_velocities(t, 1500.0, 4242)
    comm = LC.HostComm(box, grids[world], 12.5)
    serial = LC.HostComm(box, (1, 1, 1), 12.5)
    out = {"world": world, "cells": cells, "atoms": int(len(x)), "steps": steps}
    ok = True
    for label, kw in (("peer", {}), ("peer_async_qeq", {}), ("nccl", {"RXB_PEER": "0"})):
        for k, val in kw.items():
            os.environ[k] = val
        a = LC.host_md(Rxb, H, comm, rank, dev, box, x, v, t, tag, steps, uid=uid[0], use_comm=True, allgather=ag, tol=1e-10,
                       async_qeq=label.endswith("async_qeq"))
        if label == "peer":
            b = LC.host_md(Rxb, H, serial, 0, dev, box, x, v, t, tag, steps, use_comm=False, tol=1e-10)
        c = LC.compare(a, b)
        c["ghosts_rank0"] = a["nghost"]; c["locals_rank0"] = a["nlocal"]
        c["migrated"] = int((a["owner"] != comm.assign(comm.wrap(np.array(x)))).sum())
        c["ok"] = bool(c["pe_rel"] < 1e-10 and c["f_rel"] < 1e-8 and c["dq"] < 1e-8 and c["dx"] < 1e-9 and c["ghost_q_err"] == 0.0)
        ok = ok and c["ok"]
        out[label] = c
        if world > 1:
            # a new communicator for the second transport: fresh unique id
            uid = [Rxb.dist_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
    out["ok"] = ok
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
