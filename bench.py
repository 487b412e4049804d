#!/usr/bin/env python
"""bench.py — atom-timesteps/s of the TATB ReaxFF+QEq hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU semantics (oracle) on host cores

A "step" is one full MD timestep of the hot path: fix nve initial -> (reneighbour every 5 | ghost forward) ->
fix qeq/reax pre_force (H build + dual-RHS pipelined CG) -> pair reax/c compute -> reverse -> fix nve final, with the
settings of in.reaxc.lattice:825-837: skin 2.5, every 5, qeq tol 1e-6, thermo 5, **timestep 0.625 fs** (the script's live
value; its commented alternative 0.0625 fs is reported next to it under `secondary`).
`value`  : whole-job atom-timesteps/s with everything resident in HBM (device time, CUDA events on the launch stream).
`e2e`    : same metric through the LAMMPS-facing C ABI with HOST buffers (x in / f out every step, copies inside the
           timed region, host-side integration and ghost forward/reverse as the LAMMPS core would do).
`configs`: the other BASELINE.json configurations (C3 strong 1.57 M atoms, C4 weak 393,216 atoms/GPU at tol 1e-8, C5
           hot-compressed 3000 K with fix reax/c/bonds + fix reax/c/species), short runs of the same resident path.
`parity` : at N > 1 the N-rank run is compared with a single-GPU run of the same system before anything is timed.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "atom-timesteps/s (TATB ReaxFF+QEq)"
DT_SCRIPT = 0.625      # in.reaxc.lattice:837 `timestep 0.625 #0.0625`
DT_ALT = 0.0625
TOL = 1e-6


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def committed_traffic(kernel, cells):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/*_traffic.json, written by
    tests/summarise_profiles.py from the raw CSV); None when no capture of this kernel on this workload is committed."""
    import glob
    best = None
    for p in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json"))):
        try:
            for e in json.load(open(p)):
                if e.get("kernel") == kernel and tuple(e.get("cells", ())) == tuple(cells):
                    best = (float(e["dram_bytes_per_launch"]), os.path.relpath(p, ROOT) + " <- " + e.get("source", "?"),
                            e.get("l1_data_pipe_pct"))
        except Exception:  # noqa: BLE001
            pass
    return best


class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self.ok = False
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv = nv
            self.h = nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = str(e)
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.02)

    def __enter__(self):
        if self.ok:
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.ok:
            self.t.join(timeout=2)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


WEAK = {1: (8, 8, 8), 2: (16, 8, 8), 4: (16, 16, 8), 8: (16, 16, 16)}            # 196,608 atoms per GPU (configs[1] at N=1)
WEAK_C4 = {1: (16, 8, 8), 2: (16, 16, 8), 4: (16, 16, 16), 8: (32, 16, 16)}      # 393,216 atoms per GPU (configs[3])
STRONG_C3 = (16, 16, 16)                                                          # 1,572,864 atoms (configs[2])


def workload(n_gpus, cells=None):
    """configs[1] of BASELINE.json at N=1 (TATB 8x8x8 = 196,608 atoms); weak scaling keeps 196,608 atoms per GPU."""
    return tuple(cells) if cells else WEAK[n_gpus]


def workload_string(cells, dt=DT_SCRIPT, tol=TOL):
    n = 384 * cells[0] * cells[1] * cells[2]
    return (f"TATB {cells[0]}x{cells[1]}x{cells[2]} ({n} atoms) ReaxFF+QEq fp64, NVE dt {dt:g} fs, qeq/reax tol {tol:g}, "
            "skin 2.5, reneighbour every 5, thermo 5 (in.reaxc.lattice settings)")


def config_for(cells, dt=DT_SCRIPT, tol=TOL):
    return {"workload": workload_string(cells, dt, tol)}


def make_system(H, cells, seed=12345, T=300.0, scale=1.0):
    box, x, t, tag = H.tatb_cell(*cells, scale=scale)
    v = H.maxwell_velocities(t, T, seed)
    return box, x, v, t, tag


# ----------------------------------------------------------------------------------------------------------------
# CPU legs: the oracle (CPU restatement of the reference's serial semantics, OpenMP over atoms).  The Sunway reference
# itself cannot be built for x86 (athread/DMA runtime), so kind = "port".
def oracle_with_threads(H):
    """-> (oracle handle, threads the library really uses).  The thread count is forced through omp_set_num_threads:
    launchers such as torchrun export OMP_NUM_THREADS=1, and the environment variable is read only once per process."""
    want = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(want)
    o = H.Oracle(omp=True)
    got = o.omp_threads(want)
    return o, got


def oracle_probe(H, dt, tol, cells=(2, 2, 2)):
    """seconds per atom-step of the oracle on a small sample (one reneighbouring period after the first QEq)."""
    box, x, v, t, tag = make_system(H, cells)
    o, thr = oracle_with_threads(H)
    o.md_init(box, x, v, t, tag, dt=dt, qeq_tol=tol)
    o.md_run(2)
    t0 = time.time(); o.md_run(5); per = (time.time() - t0) / 5
    return per / len(x), thr


def choose_sample(cells, per_atom_step, nsteps, budget_s):
    """Largest replication <= cells whose (nsteps + setup ~ 4 steps) fit the budget, shrinking the longest axis first."""
    c = list(cells)
    while True:
        n = 384 * c[0] * c[1] * c[2]
        if per_atom_step * 1.25 * n * (nsteps + 4) <= budget_s or max(c) == 1:
            return tuple(c)
        k = int(np.argmax(c))
        c[k] = max(1, c[k] // 2)


def cpu_baseline(H, cells, dt, tol, budget_s=25.0):
    """`cpu_baseline` of the b200 line: a bounded sample of the same workload (about budget_s of CPU work)."""
    per, thr = oracle_probe(H, dt, tol)
    sample = choose_sample(cells, per, 3, budget_s)
    box, x, v, t, tag = make_system(H, sample)
    o, thr = oracle_with_threads(H)
    o.md_init(box, x, v, t, tag, dt=dt, qeq_tol=tol)
    o.md_run(5)                         # reneighbour once, warm the QEq history
    n = len(x)
    steps = int(max(3, min(200, budget_s / max(per * n, 1e-4))))
    t0 = time.time(); o.md_run(steps); el = time.time() - t0
    return {"value": n * steps / el, "unit": "atom-timesteps/s", "cores": thr, "kind": "port",
            "sample": f"TATB {sample[0]}x{sample[1]}x{sample[2]} ({n} atoms), dt {dt:g} fs, {steps} steps after 6 warm-up steps, "
                      f"oracle/liboracle_omp.so on {thr} OpenMP threads (as reported by the library; CPU restatement of the "
                      "reference - the Sunway reference itself cannot be built here)"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    import helpers as H
    H.build_oracle()
    big = workload(args.gpus, args.cells)
    dt, tol = args.dt, TOL
    per, thr = oracle_probe(H, dt, tol)
    # the same cell count as the GPU arm whenever K + W steps of it fit a few minutes of host time; otherwise the largest
    # replication that does, and then `config` names that sample (never the workload that was not run)
    cells = choose_sample(big, per, args.steps + args.warmup, args.ref_budget)
    box, x, v, t, tag = make_system(H, cells)
    o, thr = oracle_with_threads(H)
    o.md_init(box, x, v, t, tag, dt=dt, qeq_tol=tol)
    o.md_run(args.warmup)
    t0 = time.time(); o.md_run(args.steps); el = time.time() - t0
    n = len(x)
    val = n * args.steps / el
    same = tuple(cells) == tuple(big)
    cfg = config_for(cells, dt, tol)
    if not same:
        cfg["sample_of"] = workload_string(big, dt, tol)
        cfg["note"] = (f"bounded sample: {args.steps + args.warmup} oracle steps of the full workload would exceed the "
                       f"{args.ref_budget:.0f} s host budget (measured {per * 1e6:.2f} us per atom-step on {thr} threads)")
    sample = (f"TATB {cells[0]}x{cells[1]}x{cells[2]} ({n} atoms), {args.steps} timed steps after {args.warmup} warm-up steps, "
              f"{thr} OpenMP threads (reported by the library, os.cpu_count() = {os.cpu_count()})")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "atom-timesteps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (TATB 384-atom cell replicated by lattice translation, Maxwell velocities 300 K seed 12345)",
        "config": cfg, "same_cells_as_gpu_arm": same,
        "cpu_baseline": {"value": val, "unit": "atom-timesteps/s", "cores": thr, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "atom-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ----------------------------------------------------------------------------------------------------------------
def new_handle(H, device, tol=TOL):
    from sw_reaxff_b200 import Rxb
    r = Rxb(device)
    r.pair_settings(H.CONTROL)
    r.pair_coeff(H.FFIELD, H.ELEMENTS)
    r.fix_qeq(0.0, 10.0, tol)
    return r


def timed_run(r, natoms, steps, warm):
    """-> (atom-steps/s, ms/step, CG iterations/step) of `steps` resident steps after `warm` warm-up steps."""
    import torch
    r.md_run(warm)
    torch.cuda.synchronize()
    c0 = r.counts()
    r.md_run(steps)
    ms = r.md_last_run_ms()
    c1 = r.counts()
    return natoms * steps / (ms * 1e-3), ms / steps, float(c1[7] - c0[7]) / steps


def extra_configs_1gpu(H, device, steps):
    """BASELINE.json configs[2..4] on one GPU (short runs of the same resident path; parity for each is in tests/)."""
    import torch
    out = {}
    k = max(5, min(steps, 10))
    # C3: 1.57 M atoms (the strong-scaling system) on this GPU
    box, x, v, t, tag = make_system(H, STRONG_C3)
    r = new_handle(H, device)
    r.md_setup(box, x, v, t, tag, H.MASS, dt=DT_SCRIPT, every=5, thermo=5)
    val, ms, its = timed_run(r, len(x), k, 5)
    out["C3_strong_1.57M"] = {"value": val, "ms_per_step": ms, "atoms": len(x), "n_gpus": 1, "steps": k, "qeq_iterations_per_step": its,
                              "workload": workload_string(STRONG_C3)}
    r.close(); del r
    # C4: 393,216 atoms per GPU, QEq tolerance 1e-8
    c4 = WEAK_C4[1]
    box, x, v, t, tag = make_system(H, c4)
    r = new_handle(H, device, tol=1e-8)
    r.md_setup(box, x, v, t, tag, H.MASS, dt=DT_SCRIPT, every=5, thermo=5)
    val, ms, its = timed_run(r, len(x), k, 5)
    out["C4_weak_393k_tol1e-8"] = {"value": val, "ms_per_step": ms, "atoms": len(x), "n_gpus": 1, "steps": k, "qeq_iterations_per_step": its,
                                   "workload": workload_string(c4, DT_SCRIPT, 1e-8)}
    r.close(); del r
    # C5: hot-compressed (0.90 linear scale, 3000 K) with fix reax/c/bonds 25 + fix reax/c/species 1 25 25
    box, x, v, t, tag = make_system(H, (8, 8, 8), T=3000.0, scale=0.90)
    r = new_handle(H, device)
    r.md_setup(box, x, v, t, tag, H.MASS, dt=DT_ALT, every=5, thermo=5)
    r.species_config(1, 25, 25, natoms=len(x))
    r.md_run(25)
    torch.cuda.synchronize()
    c0 = r.counts()
    t0 = time.perf_counter()
    r.md_run(25)
    ms = r.md_last_run_ms()
    bt = r.bond_table()                                   # what fix reax/c/bonds 25 fetches at its output step
    wall = time.perf_counter() - t0
    c1 = r.counts()
    log = r.species_log()
    out["C5_hot_compressed_bonds_species"] = {
        "value": len(x) * 25 / (ms * 1e-3), "ms_per_step": ms / 25, "atoms": len(x), "n_gpus": 1, "steps": 25,
        "qeq_iterations_per_step": float(c1[7] - c0[7]) / 25, "wall_ms_per_step_incl_bond_table_d2h": 1e3 * wall / 25,
        "bond_table_entries": int(len(bt["nbr"])), "species_outputs": len(log), "molecules": int(log[-1]["nmole"]) if log else None,
        "workload": "TATB 8x8x8 (196608 atoms) compressed to 0.90 of the lattice constant, 3000 K, dt 0.0625 fs, qeq tol 1e-6, "
                    "fix reax/c/bonds 25 + fix reax/c/species 1 25 25 (species sampled every step on the device)"}
    r.close(); del r
    return out


def parity_1gpu(H, device):
    """One 384-atom force + charge evaluation against the oracle (the same check as __graft_entry__.smoke())."""
    cfg = H.static_config(1, 1, 1, perturb=0.05, seed=7, qeq=False)
    o = cfg["oracle"]
    n, x, ty, tg, owner = cfg["n"], cfg["x"], cfg["type"], cfg["tag"], cfg["owner"]
    q0 = np.zeros(len(x))
    o.set_atoms(n, x, ty, tg, q0); o.build_neighbors(12.5); o.qeq_init(0.0, 10.0, 1e-10)
    o.qeq_set_hist(np.zeros((n, 5)), np.zeros((n, 5))); o.qeq_pre_force(owner); o.compute()
    r = new_handle(H, device, tol=1e-10)
    r.set_atoms(n, x, ty, tg, q0, owner); r.neigh_build(); r.qeq_pre_force()
    qg = r.get_charges()
    res = r.pair_compute(True, True)
    fo, qo = o.forces(), o.q()
    eo, _ = o.energies()
    out = {"against": "oracle, 384-atom TATB cell perturbed 0.05 A, qeq tol 1e-10",
           "f_rel": float(np.abs(res["f"] - fo).max() / np.abs(fo).max()), "dq": float(np.abs(qg - qo).max()),
           "pe_rel": float(abs(res["eng"].sum() - eo.sum()) / abs(eo.sum()))}
    r.close()
    if not (out["f_rel"] < 1e-8 and out["dq"] < 1e-8 and out["pe_rel"] < 1e-10):
        raise RuntimeError(f"parity check failed: {out}")
    return out


def run_b200(args, rank, world, local_rank):
    import torch
    import helpers as H
    from sw_reaxff_b200 import Rxb
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    cells = workload(args.gpus, args.cells)
    if world > 1:
        from sw_reaxff_b200.dist import run_distributed
        out = run_distributed(args, rank, world, local_rank, cells, H)
        if rank == 0:
            print(json.dumps(out))
        dist.barrier()
        dist.destroy_process_group()
        return

    dt = args.dt
    parity = parity_1gpu(H, local_rank) if not args.no_parity else None
    box, x, v, t, tag = make_system(H, cells)
    natoms = len(x)
    r = new_handle(H, local_rank)
    r.md_setup(box, x, v, t, tag, H.MASS, dt=dt, every=5, thermo=5)
    warm = max(args.warmup, 3)
    r.md_run(warm)
    torch.cuda.synchronize()
    c0 = r.counts()
    ncu_range = bool(os.environ.get("RXB_NCU_RANGE"))   # `ncu --profile-from-start off`: capture exactly the timed region
    if ncu_range:
        r.profiler_range(1)
    with ClockSampler(local_rank) as cs:
        hs0 = r.host_syncs()
        r.md_run(args.steps)                      # timed region: CUDA events on the launch stream inside
        host_syncs = r.host_syncs() - hs0         # every stream / event wait of the library inside the timed region
        ms = r.md_last_run_ms()
    torch.cuda.synchronize()
    if ncu_range:
        r.profiler_range(0)
    c1 = r.counts()
    value = natoms * args.steps / (ms * 1e-3)
    launches = int(c1[6] - c0[6])
    qeq_iters = int(c1[7] - c0[7])

    # ---- per-kernel device times over the same kind of steps (lazy CUDA events, no sync inside a step) ----
    r.profile(1)
    nprof = min(args.steps, 20)
    k0 = r.counters()
    r.md_run(nprof)
    prof = r.profile(0)
    k1 = r.counters()
    cnt = r.counts()
    nnz_far, nall = int(cnt[5]), int(cnt[1])
    spmv_ms, spmv_launched = prof["spmv"]
    # SpMV launches past convergence are gated off on the device (they return at once): the average launch duration is
    # taken over the launches that really multiplied, counted on the device
    spmv_calls = max(int(k1["spmv_active"] - k0["spmv_active"]), 1)
    qeq_replays = int(k1["qeq_replays"])
    hbm_peak, peak_src = peaks()
    hfmt = r.h_format()
    # dual-RHS SpMV = (bytes per stored H entry) nnz10 + 8k(N+G) gathered + 8kN written + 8N diagonal, k = 2;
    # 12 B/entry for fp64 value + int32 column, 8 B/entry for the packed 42-bit fixed-point value + 22-bit column
    spmv_bytes = float(hfmt["bytes_per_entry"]) * nnz_far + 16.0 * nall + 16.0 * natoms + 8.0 * natoms
    spmv_avg = spmv_ms * 1e-3 / max(spmv_calls, 1)
    traffic = committed_traffic("k_spmv2", cells)
    achieved = spmv_bytes / spmv_avg / 1e9
    step_ms = sum(prof[k][0] for k in ("neigh", "qeq_farH", "qeq_cg", "bond_list", "bond_orders", "bonded", "nonbonded", "dbond")) / nprof
    breakdown = {k: round(prof[k][0] / nprof, 4) for k in prof}
    roofline = {"kernel": "k_spmv2 (dual-RHS QEq SpMV, the largest single-kernel share of the step)", "bound": "hbm",
                "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "peak_source": peak_src, "traffic": traffic[0] if traffic else None,
                "second_roof": ({"bound": "l1 data pipe (LSU wavefronts)", "busy_pct_under_ncu": traffic[2],
                                 "note": "the 16-byte CG-vector gathers cost >= 4 wavefronts per 32 entries even when coherent: with "
                                         "8-byte H entries the kernel sits on this roof as much as on HBM (DESIGN.md section 3)"}
                                if traffic and traffic[2] is not None else None),
                "traffic_source": traffic[1] if traffic else None,
                "algorithmic_bytes_per_launch": spmv_bytes, "h_entry_format": hfmt["name"],
                "avg_launch_us": spmv_avg * 1e6, "launches_per_step": spmv_calls / nprof,
                "gated_launches_per_step": (spmv_launched - spmv_calls) / nprof, "qeq_solves_continued_after_status": qeq_replays,
                "share_of_step": (spmv_ms / nprof) / step_ms}

    # the two other kernels north_star names, against the bound each one has (reported next to the contract's `roofline`)
    nnz_vl = int(cnt[2])
    farh_ms, farh_calls = prof["qeq_farH"]
    nb_ms, nb_calls = prof["nonbonded"]
    farh_bytes = 4.0 * nnz_vl + float(hfmt["bytes_per_entry"]) * nnz_far + 48.0 * nall   # Verlet indices read, far list + H written
    farh_avg = farh_ms * 1e-3 / max(farh_calls, 1)
    nb_avg = nb_ms * 1e-3 / max(nb_calls, 1)
    fp64_meas = Rxb.measure_fp64_tflops(local_rank)
    fp64_peak = fp64_meas if fp64_meas > 0 else 148 * 64 * 2 * 1.965e9 / 1e12
    nb_dp_per_pair = 143.0                                           # fp64 instructions per pair, ncu source page (profiles/)
    other = [
        {"kernel": "k_far_H (far list + H matrix)", "bound": "hbm", "achieved": farh_bytes / farh_avg / 1e9, "peak": hbm_peak,
         "unit": "GB/s", "frac": farh_bytes / farh_avg / 1e9 / hbm_peak, "avg_launch_us": farh_avg * 1e6, "peak_source": peak_src},
        {"kernel": "k_nonbonded (tapered vdW + Coulomb)", "bound": "fp64", "achieved": nb_dp_per_pair * nnz_far * 2 / nb_avg / 1e12,
         "peak": fp64_peak, "unit": "TFLOP/s (fp64, every DP instruction counted as one FMA)",
         "peak_source": "measured (rxb_measure_fp64_tflops: pure DFMA kernel on this GPU)" if fp64_meas > 0 else "nominal 148 SM x 64 lanes x 2 x 1.965 GHz",
         "frac": nb_dp_per_pair * nnz_far * 2 / nb_avg / 1e12 / fp64_peak, "avg_launch_us": nb_avg * 1e6,
         "pairs_per_launch": nnz_far},
    ]
    r.close(); del r

    # ---- the script's commented alternative dt (0.0625 fs): fewer CG iterations per step ----
    secondary = {}
    if not args.quick:
        alt = DT_ALT if abs(dt - DT_ALT) > 1e-12 else DT_SCRIPT
        r2 = new_handle(H, local_rank)
        r2.md_setup(box, x, v, t, tag, H.MASS, dt=alt, every=5, thermo=5)
        k2 = min(args.steps, 40)
        val2, ms2, its2 = timed_run(r2, natoms, k2, warm)
        secondary[f"dt_{alt:g}fs"] = {"value": val2, "ms_per_step": ms2, "qeq_iterations_per_step": its2, "steps": k2,
                                      "workload": workload_string(cells, alt, TOL)}
        r2.close(); del r2

    # ---- end to end through the LAMMPS-facing plugin calls with HOST buffers (C++ styles in sw_reaxff_b200/host) ----
    import ctypes as C
    hostlib = C.CDLL(os.path.join(ROOT, "sw_reaxff_b200", "librxb200_host.so"))
    script = os.path.join(H.DATA, "in.tatb.b200")
    if cells[0] == cells[1] == cells[2]:
        kv = {"S": cells[0], "t": 0, "T": 300.0, "D": H.DATA, "dt": dt, "tol": TOL}
        names = (C.c_char_p * len(kv))(*[k.encode() for k in kv]); vals = (C.c_char_p * len(kv))(*[str(v).encode() for v in kv.values()])
        out4 = (C.c_double * 4)(); err = C.create_string_buffer(512)
        e2e_steps = min(args.steps, 50)
        rc = hostlib.rxh_bench_script(script.encode(), len(kv), names, vals, local_rank, warm, e2e_steps, out4, err, 512)
        if rc != 0:
            raise RuntimeError("e2e run failed: " + err.value.decode())
        nall2, e2e_dt = int(out4[1]), float(out4[2])
        e2e = {"value": natoms * e2e_steps / e2e_dt, "unit": "atom-timesteps/s", "h2d_bytes_per_step": nall2 * 24,
               "d2h_bytes_per_step": nall2 * 24 + 22 * 8, "steps": e2e_steps, "ms_per_step": 1e3 * e2e_dt / e2e_steps,
               "note": "C++ host styles (PairReaxCB200::compute, FixQEqReaxB200::pre_force, FixNVEB200) on the LAMMPS-core stand-in, "
                       f"script sw_reaxff_b200/data/tatb/in.tatb.b200 (dt {dt:g}): host x uploaded (rxb_set_positions / rxb_set_atoms + "
                       "rxb_neigh_build on reneighbouring steps), host f downloaded (rxb_pair_compute) every step; host-side "
                       "integration, borders/forward/reverse comm inside the timed region (wall clock)"}
    else:
        e2e = {"value": None, "unit": "atom-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "note": "non-cubic replication: not measured"}

    configs = extra_configs_1gpu(H, local_rank, args.steps) if not (args.quick or args.cells) else {}
    cpu = (cpu_baseline(H, cells, dt, TOL) if not args.no_cpu_baseline
           else {"value": None, "unit": "atom-timesteps/s", "cores": 0, "kind": "port", "sample": "skipped"})
    out = {
        "metric": METRIC, "value": value, "unit": "atom-timesteps/s", "n_gpus": 1, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (TATB 384-atom cell replicated by lattice translation, Maxwell velocities 300 K seed 12345)",
        "config": {**config_for(cells, dt, TOL), "ghost_atoms": nall - natoms,
                   "l2": "inputs larger than L2 (H matrix + Verlet list > 1 GB per step vs 126 MB L2)",
                   "timing": "CUDA events on the launch stream around the K steps"},
        "clocks": cs.summary(), "e2e": e2e, "gpu_launches": launches, "host_syncs_per_step": host_syncs / args.steps,
        "qeq_cg_iterations_per_s": qeq_iters / (ms * 1e-3),
        "qeq_iterations_per_step": qeq_iters / args.steps, "roofline": roofline, "cpu_baseline": cpu,
        "kernel_ms_per_step": breakdown, "other_kernels": other, "secondary": secondary, "configs": configs, "parity": parity,
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cells", type=int, nargs=3, default=None, help="override the replication (parity-size runs)")
    ap.add_argument("--dt", type=float, default=DT_SCRIPT, help="timestep in fs (default: the script's live 0.625)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--quick", action="store_true", help="headline + roofline + e2e only (no secondary dt, no extra configs)")
    ap.add_argument("--ref-budget", type=float, default=240.0, help="host seconds the --impl reference run may take")
    ap.add_argument("--strong", action="store_true", help="strong scaling: TATB 16x16x16 (1,572,864 atoms) on any N (configs[2])")
    args = ap.parse_args()
    if args.strong and not args.cells:
        args.cells = list(STRONG_C3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
