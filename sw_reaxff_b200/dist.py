"""Multi-GPU launcher glue (one process per GPU, torchrun): processor grid, brick-local lattice generation, NCCL id
hand-off to the library, max-over-ranks timing.  The halo exchange / all-reduce themselves run inside librxb200.so
(csrc/rxb_dist.cu); torch.distributed is only the rendezvous + result plumbing.

Spatial decomposition over px*py*pz bricks; the exchange itself lives in csrc/rxb_dist.cu.  `parity_check` compares the
N-rank run with a single-GPU run of the same system (positions, forces, charges, energies by atom tag, incl. migration)
and is executed by bench.py before anything is timed at N > 1 and by tests/gpu_dist_check.py.  Host-side logic in this
file is covered by the world_size-2 gloo test (tests/test_dist_gloo.py).
"""
import os
import time

import numpy as np

GRIDS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def processor_grid(world):
    """px,py,pz with px*py*pz == world, as cubic as possible, x fastest (LAMMPS 'processors * * *' spirit)."""
    if world in GRIDS:
        return GRIDS[world]
    best = None
    for px in range(1, world + 1):
        if world % px:
            continue
        for py in range(1, world // px + 1):
            if (world // px) % py:
                continue
            pz = world // px // py
            score = max(px, py, pz) - min(px, py, pz)
            if best is None or score < best[0]:
                best = (score, (px, py, pz))
    return best[1]


def rank_coords(rank, grid):
    px, py, pz = grid
    return rank % px, (rank // px) % py, rank // (px * py)


def brick_cells(cells, grid, rank):
    """Unit cells [lo,hi) of the replicated lattice owned by `rank` (cells must divide evenly for a balanced start;
    any remainder goes to the last brick in each dimension)."""
    out = []
    for c, p, k in zip(cells, grid, rank_coords(rank, grid)):
        base = c // p
        lo = k * base
        hi = c if k == p - 1 else lo + base
        out.append((lo, hi))
    return out


def local_lattice(H, cells, grid, rank, T=300.0, seed=12345, scale=1.0):
    """Atoms of the global TATB cells[0] x cells[1] x cells[2] lattice that start on `rank`, with GLOBAL tags and the
    same velocities the single-GPU run draws (per-atom RNG keyed by the global tag, so any decomposition agrees)."""
    box6, x0, t0, _ = H.read_data_tatb()
    a = np.array([box6[0], 0, 0]); b = np.array([box6[3], box6[1], 0]); c = np.array([box6[4], box6[5], box6[2]])
    (x0c, x1c), (y0c, y1c), (z0c, z1c) = brick_cells(cells, grid, rank)
    xs, ts, tags = [], [], []
    for iz in range(z0c, z1c):
        for iy in range(y0c, y1c):
            for ix in range(x0c, x1c):
                cell = (iz * cells[1] + iy) * cells[0] + ix        # same cell order as helpers.tatb_cell
                xs.append(x0 + ix * a + iy * b + iz * c)
                ts.append(t0)
                tags.append(cell * 384 + np.arange(1, 385))
    x = np.ascontiguousarray(np.concatenate(xs))
    t = np.concatenate(ts).astype(np.int32)
    tag = np.concatenate(tags).astype(np.int32)
    box = np.array([box6[0] * cells[0], box6[1] * cells[1], box6[2] * cells[2], box6[3] * cells[1], box6[4] * cells[2], box6[5] * cells[2]])
    if scale != 1.0:
        x = x * scale
        box = box * scale
    v = velocities_by_tag(H, t, tag, T, seed)
    return box, x, v, t, tag


def velocities_by_tag(H, types, tags, T, seed):
    """Gaussian velocities that depend only on (seed, tag): identical for every decomposition.  (No global momentum or
    temperature rescale: those need a reduction; the bench only needs a thermal state.)"""
    kB, mvv2e = 0.0019872067, 48.88821291 ** 2
    m = H.MASS[types]
    # counter-based normal draws: hash (seed, tag, component) -> two uniforms -> Box-Muller
    def u(k):
        with np.errstate(over="ignore"):
            z = tags.astype(np.uint64) * np.uint64(6364136223846793005) + np.uint64(seed) * np.uint64(1442695040888963407)
            z = z + np.uint64(k) * np.uint64(0x9E3779B97F4A7C15)
            z ^= z >> np.uint64(33); z = z * np.uint64(0xff51afd7ed558ccd)
            z ^= z >> np.uint64(33); z = z * np.uint64(0xc4ceb9fe1a85ec53)
            z ^= z >> np.uint64(33)
        return ((z >> np.uint64(11)).astype(np.float64) + 0.5) / float(1 << 53)
    g = np.empty((len(tags), 3))
    for k in range(3):
        g[:, k] = np.sqrt(-2.0 * np.log(u(2 * k))) * np.cos(2 * np.pi * u(2 * k + 1))
    return np.ascontiguousarray(g * np.sqrt(kB * T / (m * mvv2e))[:, None])


def broadcast_unique_id(dist, rxb_cls, rank, device):
    import torch
    if rank == 0:
        raw = rxb_cls.dist_unique_id()
        t = torch.tensor(list(raw), dtype=torch.uint8, device=device)
    else:
        t = torch.zeros(128, dtype=torch.uint8, device=device)
    dist.broadcast(t, src=0)
    return bytes(t.cpu().tolist())


def max_over_ranks(dist, value, device):
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(dist, values, device):
    import torch
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().tolist()


def setup_distributed(H, rank, world, local_rank, cells, tol=1e-6, thermo=5, T=300.0, p2p=1, dt=0.0625, scale=1.0, every=5):
    import torch
    import torch.distributed as dist
    from .api import Rxb
    grid = processor_grid(world)
    dev = torch.device("cuda", local_rank)
    box, x, v, t, tag = local_lattice(H, cells, grid, rank, T=T, scale=scale)
    r = Rxb(local_rank)
    r.pair_settings(H.CONTROL)
    r.pair_coeff(H.FFIELD, H.ELEMENTS)
    r.fix_qeq(0.0, 10.0, tol)
    uid = broadcast_unique_id(dist, Rxb, rank, dev)
    r.dist_init(rank, world, uid, grid)
    r.dist_set_p2p(p2p)
    r.md_setup(box, x, v, t, tag, H.MASS, dt=dt, every=every, thermo=thermo)
    return r, grid, len(x)


def gather_by_tag(dist, dev, world, r):
    """(tag, x, f, q) of every atom of the decomposed run, sorted by tag, on every rank."""
    import torch
    out = r.md_get()
    n = int(r.counts()[0])
    nmax = int(max_over_ranks(dist, float(n), dev))
    tags = r.local_tags()
    pad = torch.zeros(nmax, 8, dtype=torch.float64, device=dev)
    blk = np.concatenate([tags[:, None].astype(np.float64), out["x"], out["f"], out["q"][:, None]], axis=1)
    pad[:n] = torch.from_numpy(blk).to(dev)
    bufs = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    allr = torch.cat(bufs).cpu().numpy()
    allr = allr[allr[:, 0] > 0]
    return allr[np.argsort(allr[:, 0])]


def parity_check(H, rank, world, local_rank, cells=(4, 4, 2), steps=12, T=3000.0, tol=1e-10, p2p=1, dt=0.25, species=True):
    """Decomposition invariance: the N-rank run and a single-GPU run (on rank 0's GPU) of the same global system must
    agree by atom tag on positions, forces, charges and energies after `steps` MD steps (reneighbouring every 5, so atoms
    migrate between bricks at 3000 K), and on the fix reax/c/species output.  Returns the deviations on every rank
    (broadcast from rank 0); `ok` is the verdict."""
    import torch
    import torch.distributed as dist
    from .api import Rxb
    dev = torch.device("cuda", local_rank)
    r, grid, n0 = setup_distributed(H, rank, world, local_rank, cells, tol=tol, thermo=1, T=T, p2p=p2p, dt=dt)
    th0 = r.md_thermo()
    tags0 = set(r.local_tags().tolist())
    natoms = 384 * cells[0] * cells[1] * cells[2]
    if species:
        r.species_config(1, 5, 5, natoms=natoms)
    r.md_run(steps)
    sp_log = r.species_log() if species else []
    bt = r.bond_table()
    bt_entries = int(sum_over_ranks(dist, [float(len(bt["nbr"]))], dev)[0])
    th = r.md_thermo()
    n_now = int(r.counts()[0])
    arrived = len(set(r.local_tags().tolist()) - tags0)      # atoms this rank received from other bricks
    arrived_total = int(sum_over_ranks(dist, [float(arrived)], dev)[0])
    allr = gather_by_tag(dist, dev, world, r)
    r.close()
    res = np.zeros(10)
    if rank == 0:
        assert len(allr) == natoms and np.array_equal(allr[:, 0].astype(np.int64), np.arange(1, natoms + 1)), "atoms lost or duplicated"
        box, x, t, tag = H.tatb_cell(*cells)
        v = velocities_by_tag(H, t, tag, T, 12345)
        s = Rxb(local_rank)
        s.pair_settings(H.CONTROL); s.pair_coeff(H.FFIELD, H.ELEMENTS); s.fix_qeq(0.0, 10.0, tol)
        s.md_setup(box, x, v, t, tag, H.MASS, dt=dt, every=5, thermo=1)
        s0 = s.md_thermo()
        if species:
            s.species_config(1, 5, 5, natoms=natoms)
        s.md_run(steps)
        ref_log = s.species_log() if species else []
        ok_sp = len(ref_log) == len(sp_log) and all(
            a["step"] == b["step"] and a["nmole"] == b["nmole"] and np.array_equal(a["composition"], b["composition"])
            for a, b in zip(ref_log, sp_log))
        ok_sp = ok_sp and (not species or len(ref_log) == steps // 5) and len(s.bond_table()["nbr"]) == bt_entries
        ref = s.md_get(); sth = s.md_thermo()
        s.close()
        # positions may differ by a box vector after wrapping: compare through the lamda-space minimum image
        dx = allr[:, 1:4] - ref["x"]
        a = np.array([box[0], 0, 0]); b = np.array([box[3], box[1], 0]); cc = np.array([box[4], box[5], box[2]])
        Hm = np.stack([a, b, cc], axis=1)
        lam = np.linalg.solve(Hm, dx.T).T
        dx = (Hm @ (lam - np.round(lam)).T).T
        res[:] = [np.abs(dx).max(), np.abs(allr[:, 4:7] - ref["f"]).max() / np.abs(ref["f"]).max(),
                  np.abs(allr[:, 7] - ref["q"]).max(), abs(th0["pe"] - s0["pe"]) / abs(s0["pe"]),
                  abs(th["pe"] - sth["pe"]) / abs(sth["pe"]), abs(th["ke"] - sth["ke"]) / abs(sth["ke"]),
                  1.0 if ok_sp else 0.0, float((np.abs(lam) > 0.5).any(axis=1).sum()), float(arrived_total), float(bt_entries)]
    tt = torch.tensor(res, dtype=torch.float64, device=dev)
    dist.broadcast(tt, src=0)
    res = tt.cpu().numpy()
    out = {"against": f"single-GPU run of the same system: TATB {cells[0]}x{cells[1]}x{cells[2]} ({natoms} atoms), {steps} steps, "
                      f"T {T:g} K, dt {dt:g} fs, qeq tol {tol:g}, reneighbour every 5, grid {grid[0]}x{grid[1]}x{grid[2]}",
           "dx": float(res[0]), "f_rel": float(res[1]), "dq": float(res[2]), "pe0_rel": float(res[3]), "pe_rel": float(res[4]),
           "ke_rel": float(res[5]), "species_and_bond_table_identical": bool(res[6] == 1.0), "atoms_wrapped": int(res[7]),
           "atoms_migrated_between_bricks": int(res[8]), "bond_table_entries": int(res[9])}
    out["ok"] = bool(out["dx"] < 1e-8 and out["f_rel"] < 1e-8 and out["dq"] < 1e-8 and out["pe0_rel"] < 1e-9 and out["pe_rel"] < 1e-8
                     and out["ke_rel"] < 1e-6 and out["species_and_bond_table_identical"])
    return out


def _timed(dist, dev, r, natoms_total, steps, warm):
    import torch
    r.md_run(warm)
    torch.cuda.synchronize(); dist.barrier()
    c0 = r.counts()
    r.md_run(steps)
    ms_local = r.md_last_run_ms()
    torch.cuda.synchronize(); dist.barrier()
    c1 = r.counts()
    ms = max_over_ranks(dist, ms_local, dev)
    return natoms_total * steps / (ms * 1e-3), ms / steps, float(c1[7] - c0[7]) / steps


def extra_configs(H, rank, world, local_rank, steps):
    """BASELINE.json configs[2..4] at N ranks: C3 strong (1,572,864 atoms over N GPUs), C4 weak (393,216 atoms per GPU,
    tol 1e-8), C5 hot-compressed 3000 K with fix reax/c/bonds 25 + fix reax/c/species 1 25 25 (196,608 atoms per GPU)."""
    import torch
    import torch.distributed as dist
    from bench import DT_ALT, DT_SCRIPT, STRONG_C3, WEAK, WEAK_C4, workload_string
    dev = torch.device("cuda", local_rank)
    out = {}
    k = max(5, min(steps, 10))
    for name, cells, tol in (("C3_strong_1.57M", STRONG_C3, 1e-6), ("C4_weak_393k_tol1e-8", WEAK_C4[world], 1e-8)):
        r, grid, _ = setup_distributed(H, rank, world, local_rank, cells, tol=tol, dt=DT_SCRIPT)
        nat = 384 * cells[0] * cells[1] * cells[2]
        val, ms, its = _timed(dist, dev, r, nat, k, 5)
        out[name] = {"value": val, "ms_per_step": ms, "atoms": nat, "n_gpus": world, "steps": k, "qeq_iterations_per_step": its,
                     "workload": workload_string(cells, DT_SCRIPT, tol)}
        r.close(); del r
    cells = WEAK[world]
    nat = 384 * cells[0] * cells[1] * cells[2]
    r, grid, _ = setup_distributed(H, rank, world, local_rank, cells, tol=1e-6, dt=DT_ALT, T=3000.0, scale=0.90)
    r.species_config(1, 25, 25, natoms=nat)
    r.md_run(25)
    torch.cuda.synchronize(); dist.barrier()
    c0 = r.counts()
    r.md_run(25)
    ms = max_over_ranks(dist, r.md_last_run_ms(), dev)
    bt = r.bond_table()
    c1 = r.counts()
    log = r.species_log()
    ent = int(sum_over_ranks(dist, [float(len(bt["nbr"]))], dev)[0])
    out["C5_hot_compressed_bonds_species"] = {
        "value": nat * 25 / (ms * 1e-3), "ms_per_step": ms / 25, "atoms": nat, "n_gpus": world, "steps": 25,
        "qeq_iterations_per_step": float(c1[7] - c0[7]) / 25, "bond_table_entries": ent, "species_outputs": len(log),
        "molecules": int(log[-1]["nmole"]) if log else None,
        "workload": f"TATB {cells[0]}x{cells[1]}x{cells[2]} ({nat} atoms) compressed to 0.90 of the lattice constant, 3000 K, "
                    "dt 0.0625 fs, qeq tol 1e-6, fix reax/c/bonds 25 + fix reax/c/species 1 25 25"}
    r.close(); del r
    return out


def run_distributed(args, rank, world, local_rank, cells, H):
    """bench.py body for N > 1: weak scaling, 196,608 atoms per GPU, spatial decomposition."""
    import torch
    import torch.distributed as dist
    from bench import ClockSampler, METRIC, TOL, config_for, peaks
    dev = torch.device("cuda", local_rank)
    # parity first: nothing is timed on a decomposition that does not reproduce the single-GPU answer
    parity = None
    if not getattr(args, "no_parity", False):
        parity = parity_check(H, rank, world, local_rank)
        if not parity["ok"]:
            raise RuntimeError(f"N-rank parity check failed, nothing timed: {parity}")
    dt = args.dt
    r, grid, n0 = setup_distributed(H, rank, world, local_rank, cells, dt=dt)
    natoms_total = 384 * cells[0] * cells[1] * cells[2]
    warm = max(args.warmup, 3)
    r.md_run(warm)
    torch.cuda.synchronize()
    dist.barrier()
    c0 = r.counts()
    with ClockSampler(local_rank) as cs:
        t0 = time.perf_counter()
        r.md_run(args.steps)
        ms_local = r.md_last_run_ms()
        torch.cuda.synchronize()
        dist.barrier()
        wall = time.perf_counter() - t0
    c1 = r.counts()
    ms = max_over_ranks(dist, ms_local, dev)
    launches, qeq_it = sum_over_ranks(dist, [float(c1[6] - c0[6]), float(c1[7] - c0[7])], dev)
    value = natoms_total * args.steps / (ms * 1e-3)
    # per-kernel profile on rank 0's GPU
    r.profile(1)
    nprof = min(args.steps, 20)
    k0 = r.counters()
    r.md_run(nprof)
    prof = r.profile(0)
    k1 = r.counters()
    cnt = r.counts()
    hbm_peak, peak_src = peaks()
    spmv_ms, _ = prof["spmv"]
    spmv_calls = max(int(k1["spmv_active"] - k0["spmv_active"]), 1)   # launches that really multiplied (not gated off)
    hfmt = r.h_format()
    spmv_bytes = float(hfmt["bytes_per_entry"]) * int(cnt[5]) + 16.0 * int(cnt[1]) + 24.0 * int(cnt[0])
    spmv_avg = spmv_ms * 1e-3 / max(spmv_calls, 1)
    achieved = spmv_bytes / spmv_avg / 1e9
    th = r.md_thermo()
    r.close(); del r
    configs = extra_configs(H, rank, world, local_rank, args.steps) if not (getattr(args, "quick", False) or args.cells) else {}
    out = {
        "metric": METRIC, "value": value, "unit": "atom-timesteps/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if getattr(args, "strong", False) else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (TATB 384-atom cell replicated by lattice translation, per-tag Gaussian velocities 300 K)",
        "config": {**config_for(cells, dt, TOL), "decomposition": f"{grid[0]}x{grid[1]}x{grid[2]} bricks, {natoms_total // world} atoms per GPU, "
                   "ghost shell 12.5 A; halo exchange and CG reductions inside the library (csrc/rxb_dist.cu)",
                   "l2": "inputs larger than L2", "timing": "CUDA events on each rank's launch stream, max over ranks"},
        "clocks": cs.summary(), "gpu_launches": int(launches), "qeq_iterations_per_step": qeq_it / world / args.steps,
        "wall_ms_per_step": 1e3 * wall / args.steps,
        "e2e": {"value": natoms_total * args.steps / wall, "unit": "atom-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "note": "multi-GPU runs are device resident (the host-buffer plugin path is measured at N=1); this is the "
                        "wall-clock rate around the same K steps including launch overhead and the barrier"},
        "roofline": {"kernel": "k_spmv2 on rank 0", "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved / hbm_peak, "peak_source": peak_src, "traffic": None, "h_entry_format": hfmt["name"]},
        "kernel_ms_per_step": {k: round(prof[k][0] / nprof, 4) for k in prof},
        "potential_energy_per_atom": th["pe"] / natoms_total,
        "parity": parity, "configs": configs,
    }
    return out
