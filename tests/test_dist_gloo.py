"""world_size-2 gloo test (CPU) of the multi-GPU host logic in sw_reaxff_b200/dist.py: processor grids, brick-local
lattice generation (union over ranks == the global lattice, global tags unique), decomposition-independent velocities,
and the max/sum-over-ranks reductions bench.py relies on."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H
from sw_reaxff_b200 import dist as D


def test_processor_grids():
    for w in (1, 2, 3, 4, 6, 8, 12, 16):
        g = D.processor_grid(w)
        assert g[0] * g[1] * g[2] == w
    assert D.processor_grid(8) == (2, 2, 2) and D.processor_grid(2) == (2, 1, 1)
    seen = {D.rank_coords(r, (2, 2, 2)) for r in range(8)}
    assert len(seen) == 8


def test_brick_cells_partition():
    cells, grid = (16, 8, 8), (2, 2, 1)
    cover = np.zeros(cells, dtype=int)
    for r in range(4):
        (a, b), (c, d), (e, f) = D.brick_cells(cells, grid, r)
        cover[a:b, c:d, e:f] += 1
    assert (cover == 1).all()
    # uneven split: remainder goes to the last brick
    (a, b), _, _ = D.brick_cells((5, 1, 1), (2, 1, 1), 1)
    assert (a, b) == (2, 5)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, cells, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    grid = D.processor_grid(world)
    box, x, v, t, tag = D.local_lattice(H, cells, grid, rank)
    dev = torch.device("cpu")
    n_total, = D.sum_over_ranks(dist, [float(len(x))], dev)
    slowest = D.max_over_ranks(dist, 10.0 + rank, dev)
    # gather everything on rank 0 through gloo and compare with the single-process lattice
    n_max = int(D.max_over_ranks(dist, float(len(x)), dev))
    pad = np.zeros((n_max, 8)); pad[:len(x), :3] = x; pad[:len(x), 3:6] = v; pad[:len(x), 6] = tag; pad[:len(x), 7] = t
    buf = [torch.zeros(n_max, 8, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(buf, torch.from_numpy(pad))
    if rank == 0:
        allr = np.concatenate([b.numpy() for b in buf])
        allr = allr[allr[:, 6] > 0]
        q.put((n_total, slowest, allr, box))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_lattice_union_equals_global():
    cells = (2, 1, 1)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, cells, q)) for r in range(2)]
    for p in procs:
        p.start()
    n_total, slowest, allr, box = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    gbox, gx, gt, gtag = H.tatb_cell(*cells)
    assert n_total == len(gx) == len(allr)
    assert slowest == 11.0                                   # max over ranks
    np.testing.assert_allclose(box, gbox)
    order = np.argsort(allr[:, 6])
    allr = allr[order]
    assert np.array_equal(allr[:, 6].astype(np.int64), np.arange(1, len(gx) + 1))    # global tags, each exactly once
    np.testing.assert_allclose(allr[:, :3], gx, atol=1e-12)  # same positions as the single-process lattice, by tag
    assert np.array_equal(allr[:, 7].astype(np.int32), gt)
    # velocities depend only on (seed, tag): the single-process draw agrees with the per-rank draws
    v_ref = D.velocities_by_tag(H, gt, gtag, 300.0, 12345)
    np.testing.assert_allclose(allr[:, 3:6], v_ref, atol=1e-15)
    kB, mvv2e = 0.0019872067, 48.88821291 ** 2
    ke = 0.5 * mvv2e * (H.MASS[gt][:, None] * v_ref ** 2).sum()
    T = 2 * ke / (3 * len(gx) * kB)
    assert 270 < T < 330
