#!/usr/bin/env python
"""bench.py — atom-timesteps/s of the TATB ReaxFF+QEq hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU semantics (oracle) on host cores

A "step" is one full MD timestep of the hot path: fix nve initial -> (reneighbour every 5 | ghost forward) ->
fix qeq/reax pre_force (H build + dual-RHS pipelined CG) -> pair reax/c compute -> reverse -> fix nve final,
script settings of in.reaxc.lattice:825-837 (skin 2.5, every 5, qeq tol 1e-6, thermo 5, dt 0.0625 fs).
`value`  : whole-job atom-timesteps/s with everything resident in HBM (device time, CUDA events on the launch stream).
`e2e`    : same metric through the LAMMPS-facing C ABI with HOST buffers (x in / f out every step, copies inside the
           timed region, host-side integration and ghost forward/reverse as the LAMMPS core would do).
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
FTM2V = 1.0 / 48.88821291 / 48.88821291


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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


def workload(n_gpus, cells=None):
    """configs[1] of BASELINE.json at N=1 (TATB 8x8x8 = 196,608 atoms); weak scaling keeps 196,608 atoms per GPU."""
    if cells:
        return tuple(cells)
    grid = {1: (8, 8, 8), 2: (16, 8, 8), 4: (16, 16, 8), 8: (16, 16, 16)}
    return grid[n_gpus]


def config_for(cells):
    n = 384 * cells[0] * cells[1] * cells[2]
    return {"workload": f"TATB {cells[0]}x{cells[1]}x{cells[2]} ({n} atoms) ReaxFF+QEq fp64, NVE dt 0.0625 fs, qeq/reax tol 1e-6, "
                        "skin 2.5, reneighbour every 5, thermo 5 (in.reaxc.lattice settings)"}


def make_system(H, cells, seed=12345, T=300.0):
    box, x, t, tag = H.tatb_cell(*cells)
    v = H.maxwell_velocities(t, T, seed)
    return box, x, v, t, tag


# ----------------------------------------------------------------------------------------------------------------
def cpu_baseline(H, budget_s=15.0, cells=(2, 2, 2), steps=None):
    """The oracle (CPU restatement of the reference's serial semantics, OpenMP over atoms) on a bounded sample."""
    ncores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(ncores))
    box, x, v, t, tag = make_system(H, cells)
    o = H.Oracle(omp=True)
    o.md_init(box, x, v, t, tag, dt=0.0625, qeq_tol=1e-6)
    o.md_run(5)                         # reneighbour once, warm the QEq history
    t0 = time.time(); o.md_run(1); per = time.time() - t0
    if steps is None:
        steps = int(max(3, min(200, budget_s / max(per, 1e-3))))
    t0 = time.time(); o.md_run(steps); dt = time.time() - t0
    n = len(x)
    return {"value": n * steps / dt, "unit": "atom-timesteps/s", "cores": ncores, "kind": "port",
            "sample": f"TATB {cells[0]}x{cells[1]}x{cells[2]} ({n} atoms), {steps} steps after 6 warm-up steps, oracle/liboracle_omp.so "
                      f"(CPU restatement of the reference; the Sunway reference itself cannot be built here)"}, steps, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    import helpers as H
    H.build_oracle()
    ncores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(ncores))
    cells = (2, 2, 2)
    box, x, v, t, tag = make_system(H, cells)
    o = H.Oracle(omp=True)
    o.md_init(box, x, v, t, tag, dt=0.0625, qeq_tol=1e-6)
    t0 = time.time(); o.md_run(1); per = time.time() - t0
    if per * (args.steps + args.warmup) > 240.0:        # keep the whole run within a few minutes
        cells = (1, 1, 1)
        box, x, v, t, tag = make_system(H, cells)
        o = H.Oracle(omp=True)
        o.md_init(box, x, v, t, tag, dt=0.0625, qeq_tol=1e-6)
    o.md_run(args.warmup)
    t0 = time.time(); o.md_run(args.steps); dt = time.time() - t0
    n = len(x)
    val = n * args.steps / dt
    big = workload(args.gpus, args.cells)
    sample = (f"each step = one MD timestep of a bounded sample of the workload: TATB {cells[0]}x{cells[1]}x{cells[2]} "
              f"({n} atoms), {ncores} OpenMP threads")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "atom-timesteps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic (TATB lattice replicated, Maxwell 300 K)",
        "config": config_for(big),
        "cpu_baseline": {"value": val, "unit": "atom-timesteps/s", "cores": ncores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "atom-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


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

    box, x, v, t, tag = make_system(H, cells)
    natoms = len(x)
    r = Rxb(local_rank)
    r.pair_settings(H.CONTROL)
    r.pair_coeff(H.FFIELD, H.ELEMENTS)
    r.fix_qeq(0.0, 10.0, 1e-6)
    r.md_setup(box, x, v, t, tag, H.MASS, dt=0.0625, every=5, thermo=5)
    r.md_run(max(args.warmup, 3))
    torch.cuda.synchronize()
    c0 = r.counts()
    ncu_range = bool(os.environ.get("RXB_NCU_RANGE"))   # `ncu --profile-from-start off`: capture exactly the timed region
    if ncu_range:
        r.profiler_range(1)
    with ClockSampler(local_rank) as cs:
        r.md_run(args.steps)                      # timed region: CUDA events on the launch stream inside
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
    r.md_run(nprof)
    prof = r.profile(0)
    cnt = r.counts()
    nnz_far, nall = int(cnt[5]), int(cnt[1])
    spmv_ms, spmv_calls = prof["spmv"]
    hbm_peak, peak_src = peaks()
    # dual-RHS SpMV = 12 nnz10 + 8k(N+G) + 8kN + 8N bytes, k = 2
    spmv_bytes = 12.0 * nnz_far + 16.0 * nall + 16.0 * natoms + 8.0 * natoms
    spmv_avg = spmv_ms * 1e-3 / max(spmv_calls, 1)
    # DRAM bytes of one k_spmv2 launch (dram__bytes_read.sum + dram__bytes_write.sum) from the committed ncu --set full
    # capture of this workload, profiles/r01e_ncu_full_k_spmv2.csv: 1.1140 GB + 6.5 MB
    spmv_traffic = 1.1205e9 if tuple(cells) == (8, 8, 8) else None
    achieved = spmv_bytes / spmv_avg / 1e9
    step_ms = sum(prof[k][0] for k in ("neigh", "qeq_farH", "qeq_cg", "bond_list", "bond_orders", "bonded", "nonbonded", "dbond")) / nprof
    breakdown = {k: round(prof[k][0] / nprof, 4) for k in prof}
    roofline = {"kernel": "k_spmv2 (dual-RHS QEq SpMV, the largest single-kernel share of the step)", "bound": "hbm",
                "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "peak_source": peak_src, "traffic": spmv_traffic,
                "traffic_source": "ncu --set full, profiles/r01e_ncu_full_k_spmv2.csv (dram read + write per launch)" if spmv_traffic else None,
                "algorithmic_bytes_per_launch": spmv_bytes,
                "avg_launch_us": spmv_avg * 1e6, "launches_per_step": spmv_calls / nprof,
                "share_of_step": (spmv_ms / nprof) / step_ms}

    # the two other kernels north_star names, against the bound each one has (reported next to the contract's `roofline`)
    nnz_vl = int(cnt[2])
    farh_ms, farh_calls = prof["qeq_farH"]
    nb_ms, nb_calls = prof["nonbonded"]
    farh_bytes = 4.0 * nnz_vl + 12.0 * nnz_far + 48.0 * nall          # Verlet indices read, far list + H written
    farh_avg = farh_ms * 1e-3 / max(farh_calls, 1)
    nb_avg = nb_ms * 1e-3 / max(nb_calls, 1)
    fp64_peak = 148 * 64 * 2 * 1.965e9 / 1e12                        # DFMA lanes x 2 flop x max SM clock (TFLOP/s)
    nb_dp_per_pair = 143.0                                           # fp64 instructions per pair, ncu source page (profiles/)
    other = [
        {"kernel": "k_far_H (far list + H matrix)", "bound": "hbm", "achieved": farh_bytes / farh_avg / 1e9, "peak": hbm_peak,
         "unit": "GB/s", "frac": farh_bytes / farh_avg / 1e9 / hbm_peak, "avg_launch_us": farh_avg * 1e6,
         "note": "gather-bound: ncu shows the L1 data pipe 77 % busy (32-byte position gathers), DRAM 23 %"},
        {"kernel": "k_nonbonded (tapered vdW + Coulomb)", "bound": "fp64", "achieved": nb_dp_per_pair * nnz_far * 2 / nb_avg / 1e12,
         "peak": fp64_peak, "unit": "TFLOP/s (fp64, every DP instruction counted as one FMA)",
         "frac": nb_dp_per_pair * nnz_far * 2 / nb_avg / 1e12 / fp64_peak, "avg_launch_us": nb_avg * 1e6,
         "pairs_per_launch": nnz_far},
    ]

    # ---- end to end through the LAMMPS-facing plugin calls with HOST buffers (C++ styles in sw_reaxff_b200/host) ----
    import ctypes as C
    hostlib = C.CDLL(os.path.join(ROOT, "sw_reaxff_b200", "librxb200_host.so"))
    script = os.path.join(H.DATA, "in.tatb.b200")
    if cells[0] == cells[1] == cells[2]:
        kv = {"S": cells[0], "t": 0, "T": 300.0, "D": H.DATA}
        names = (C.c_char_p * len(kv))(*[k.encode() for k in kv]); vals = (C.c_char_p * len(kv))(*[str(v).encode() for v in kv.values()])
        out4 = (C.c_double * 4)(); err = C.create_string_buffer(512)
        e2e_steps = min(args.steps, 50)
        rc = hostlib.rxh_bench_script(script.encode(), len(kv), names, vals, local_rank, max(args.warmup, 3), e2e_steps, out4, err, 512)
        if rc != 0:
            raise RuntimeError("e2e run failed: " + err.value.decode())
        nall2, e2e_dt = int(out4[1]), float(out4[2])
        e2e = {"value": natoms * e2e_steps / e2e_dt, "unit": "atom-timesteps/s", "h2d_bytes_per_step": nall2 * 24,
               "d2h_bytes_per_step": nall2 * 24 + 22 * 8, "steps": e2e_steps, "ms_per_step": 1e3 * e2e_dt / e2e_steps,
               "note": "C++ host styles (PairReaxCB200::compute, FixQEqReaxB200::pre_force, FixNVEB200) on the LAMMPS-core stand-in, "
                       "script sw_reaxff_b200/data/tatb/in.tatb.b200: host x uploaded (rxb_set_positions / rxb_set_atoms + "
                       "rxb_neigh_build on reneighbouring steps), host f downloaded (rxb_pair_compute) every step; host-side "
                       "integration, borders/forward/reverse comm inside the timed region (wall clock)"}
    else:
        e2e = {"value": None, "unit": "atom-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "note": "non-cubic replication: not measured"}

    cpu, _, _ = cpu_baseline(H) if not args.no_cpu_baseline else ({"value": None, "unit": "atom-timesteps/s", "cores": 0, "kind": "port", "sample": "skipped"}, 0, 0)
    out = {
        "metric": METRIC, "value": value, "unit": "atom-timesteps/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (TATB 384-atom cell replicated by lattice translation, Maxwell velocities 300 K seed 12345)",
        "config": {**config_for(cells), "ghost_atoms": nall - natoms,
                   "l2": "inputs larger than L2 (H matrix 1.07 GB, Verlet list 0.69 GB per step vs 126 MB L2)",
                   "timing": "CUDA events on the launch stream around the K steps"},
        "clocks": cs.summary(), "e2e": e2e, "gpu_launches": launches, "qeq_cg_iterations_per_s": qeq_iters / (ms * 1e-3),
        "qeq_iterations_per_step": qeq_iters / args.steps, "roofline": roofline, "cpu_baseline": cpu,
        "kernel_ms_per_step": breakdown, "other_kernels": other,
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cells", type=int, nargs=3, default=None, help="override the replication (parity-size runs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--strong", action="store_true", help="strong scaling: TATB 16x16x16 (1,572,864 atoms) on any N (configs[2])")
    args = ap.parse_args()
    if args.strong and not args.cells:
        args.cells = [16, 16, 16]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
