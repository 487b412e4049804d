"""ctypes binding of include/rxb200.h (one method per C entry point, same names minus the rxb_ prefix)."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "librxb200.so")
DATA_DIR = os.path.join(HERE, "data")

SYMBOLS = [
    "rxb_create", "rxb_destroy", "rxb_last_error", "rxb_pair_settings", "rxb_pair_coeff", "rxb_pair_extract",
    "rxb_fix_qeq", "rxb_neighbor_skin", "rxb_params_dump", "rxb_set_atoms", "rxb_set_positions", "rxb_set_charges",
    "rxb_neigh_build", "rxb_qeq_pre_force", "rxb_qeq_set_history", "rxb_qeq_get_history", "rxb_get_charges",
    "rxb_pair_compute", "rxb_md_setup", "rxb_md_run", "rxb_md_get", "rxb_md_thermo", "rxb_get_counts",
    "rxb_get_neighbors", "rxb_get_bonds", "rxb_get_workspace", "rxb_get_far", "rxb_profile", "rxb_profiler_range", "rxb_md_last_run_ms", "rxb_parse_dump",
    "rxb_dist_unique_id", "rxb_dist_init", "rxb_dist_set_p2p", "rxb_md_get_tags",
    "rxb_bond_table", "rxb_bond_table_get", "rxb_species_config", "rxb_species_step", "rxb_species_result",
    "rxb_species_cluster", "rxb_species_log_size", "rxb_species_log_get", "rxb_host_register", "rxb_host_unregister", "rxb_lookup_dump", "rxb_get_cutoffs", "rxb_measure_fp64_tflops", "rxb_get_h_format",
    "rxb_qeq_matvecs", "rxb_set_h_exact", "rxb_debug_set_caps", "rxb_debug_get_caps", "rxb_get_hbond_pairs",
    "rxb_fix_qeq_params", "rxb_spec_atom_abo", "rxb_get_counters", "rxb_comm_init", "rxb_comm_set_ghosts",
    "rxb_species_avg_qxyz", "rxb_host_sync_count",
]

E_NAMES = ["e_bond", "e_ov", "e_un", "e_lp", "e_ang", "e_pen", "e_coa", "e_hb", "e_tor", "e_con", "e_vdW", "e_ele", "e_pol"]


class RxbError(RuntimeError):
    pass


_lib = None


def load_library(path=LIB_PATH):
    """Load librxb200.so.  Raises if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise RxbError(f"{path} is missing: build it first (__graft_entry__.build()); there is no CPU fallback")
    lib = C.CDLL(path)
    lib.rxb_last_error.restype = C.c_char_p
    lib.rxb_params_dump.restype = C.c_long
    lib.rxb_md_last_run_ms.restype = C.c_double
    lib.rxb_parse_dump.restype = C.c_long
    lib.rxb_lookup_dump.restype = C.c_long
    lib.rxb_measure_fp64_tflops.restype = C.c_double
    _lib = lib
    return lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Rxb:
    """One handle == one GPU-resident ReaxFF system."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = C.c_void_p()
        self._chk(self.lib.rxb_create(int(device), C.byref(h)))
        self.h = h
        self.nlocal = self.nall = 0
        self.ntypes = 0

    def _chk(self, rc):
        if rc != 0:
            raise RxbError(self.lib.rxb_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.rxb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- configuration ----
    def pair_settings(self, control_file, lgvdw=False, enobonds=True):
        self._chk(self.lib.rxb_pair_settings(self.h, control_file.encode() if control_file else None, int(lgvdw), int(enobonds)))

    def pair_coeff(self, ffield, elements):
        arr = (C.c_char_p * len(elements))(*[e.encode() for e in elements])
        self.ntypes = len(elements)
        self._chk(self.lib.rxb_pair_coeff(self.h, ffield.encode(), len(elements), arr))

    def pair_extract(self, name):
        out = np.zeros(self.ntypes + 1)
        self._chk(self.lib.rxb_pair_extract(self.h, name.encode(), _p(out), self.ntypes))
        return out

    def fix_qeq(self, swa=0.0, swb=10.0, tol=1e-6, max_iter=200):
        self._chk(self.lib.rxb_fix_qeq(self.h, C.c_double(swa), C.c_double(swb), C.c_double(tol), int(max_iter)))

    def fix_qeq_params(self, chi=None, eta=None, gamma=None):
        """fix qeq/reax <param file>: chi/eta/gamma per LAMMPS type (arrays of ntypes + 1, index 0 unused); None = reax/c."""
        if chi is None:
            self._chk(self.lib.rxb_fix_qeq_params(self.h, 0, None, None, None))
            return
        chi = _f(chi); eta = _f(eta); gamma = _f(gamma)
        self._chk(self.lib.rxb_fix_qeq_params(self.h, len(chi) - 1, _p(chi), _p(eta), _p(gamma)))

    def neighbor_skin(self, skin):
        self._chk(self.lib.rxb_neighbor_skin(self.h, C.c_double(skin)))

    def params_dump(self):
        n = self.lib.rxb_params_dump(self.h, None, C.c_long(0))
        out = np.zeros(n)
        self.lib.rxb_params_dump(self.h, _p(out), C.c_long(n))
        return out

    # ---- atoms / lists ----
    def set_atoms(self, nlocal, x, types, tags, q=None, ghost_owner=None):
        x = _f(x); types = _i(types); tags = _i(tags)
        nall = len(types)
        q = _f(q) if q is not None else np.zeros(nall)
        go = _i(ghost_owner) if ghost_owner is not None else None
        self.nlocal, self.nall = int(nlocal), int(nall)
        self._chk(self.lib.rxb_set_atoms(self.h, int(nlocal), int(nall - nlocal), _p(x), _p(types), _p(tags), _p(q), _p(go)))

    def set_positions(self, x):
        x = _f(x)
        self._chk(self.lib.rxb_set_positions(self.h, int(x.size // 3), _p(x)))

    def set_charges(self, q):
        q = _f(q)
        self._chk(self.lib.rxb_set_charges(self.h, _p(q)))

    def neigh_build(self):
        self._chk(self.lib.rxb_neigh_build(self.h))

    # ---- QEq ----
    def qeq_pre_force(self):
        mv = np.zeros(2, dtype=np.int32)
        self._chk(self.lib.rxb_qeq_pre_force(self.h, _p(mv)))
        return int(mv[0]), int(mv[1])

    def qeq_pre_force_async(self):
        """Enqueue the solve without a host round trip (settled inside pair_compute); counts via qeq_matvecs()."""
        self._chk(self.lib.rxb_qeq_pre_force(self.h, None))

    def qeq_matvecs(self):
        mv = np.zeros(2, dtype=np.int32)
        self._chk(self.lib.rxb_qeq_matvecs(self.h, _p(mv)))
        return int(mv[0]), int(mv[1])

    def debug_set_caps(self, row_cap=0, strong_cap=0, cap_bonds=0, cap_ang=0, cap_tor=0, cap_hb=0):
        self._chk(self.lib.rxb_debug_set_caps(self.h, int(row_cap), int(strong_cap), int(cap_bonds), int(cap_ang), int(cap_tor), int(cap_hb)))

    def debug_get_caps(self):
        o = np.zeros(6, dtype=np.int32)
        self._chk(self.lib.rxb_debug_get_caps(self.h, _p(o)))
        return dict(zip(["row_cap", "strong_cap", "cap_bonds", "cap_ang", "cap_tor", "cap_hb"], o.tolist()))

    def set_h_exact(self, on=True):
        self._chk(self.lib.rxb_set_h_exact(self.h, int(on)))

    def qeq_set_history(self, s_hist, t_hist):
        s = _f(s_hist); t = _f(t_hist)
        self._chk(self.lib.rxb_qeq_set_history(self.h, _p(s), _p(t)))

    def qeq_get_history(self):
        s = np.zeros((self.nlocal, 5)); t = np.zeros((self.nlocal, 5))
        self._chk(self.lib.rxb_qeq_get_history(self.h, _p(s), _p(t)))
        return s, t

    def get_charges(self):
        q = np.zeros(self.nall)
        self._chk(self.lib.rxb_get_charges(self.h, _p(q)))
        return q

    # ---- pair compute ----
    def pair_compute(self, eflag=True, vflag=True, want_forces=True, f_out=None):
        f = f_out if f_out is not None else (np.zeros((self.nall, 3)) if want_forces else None)
        pv = np.zeros(14); eng = np.zeros(2); vir = np.zeros(6)
        self._chk(self.lib.rxb_pair_compute(self.h, int(self.nall if f is None else f.size // 3), int(eflag), int(vflag), _p(f), _p(pv), _p(eng), _p(vir)))
        return dict(f=f, pvector=pv, eng=eng, virial=vir)

    # ---- resident MD ----
    def md_setup(self, box6, x, v, types, tags, mass, dt=0.0625, every=5, thermo=5, qeq=True):
        x = _f(x); v = _f(v); types = _i(types); tags = _i(tags); mass = _f(mass); box6 = _f(box6)
        self.nlocal = len(types)
        self._chk(self.lib.rxb_md_setup(self.h, _p(box6), len(types), _p(x), _p(v), _p(types), _p(tags), _p(mass),
                                        len(mass) - 1, C.c_double(dt), int(every), int(thermo), int(qeq)))
        self.nall = int(self.counts()[1])

    def md_run(self, nsteps):
        self._chk(self.lib.rxb_md_run(self.h, int(nsteps)))
        self.nall = int(self.counts()[1])

    def md_get(self):
        n = self.nlocal = int(self.counts()[0])
        x = np.zeros((n, 3)); v = np.zeros((n, 3)); f = np.zeros((n, 3)); q = np.zeros(n)
        self._chk(self.lib.rxb_md_get(self.h, _p(x), _p(v), _p(f), _p(q)))
        return dict(x=x, v=v, f=f, q=q)

    def local_tags(self):
        n = int(self.counts()[0])
        t = np.zeros(n, dtype=np.int32)
        self._chk(self.lib.rxb_md_get_tags(self.h, _p(t)))
        return t

    def md_thermo(self):
        pv = np.zeros(14); pe = C.c_double(); ke = C.c_double()
        self._chk(self.lib.rxb_md_thermo(self.h, _p(pv), C.byref(pe), C.byref(ke)))
        return dict(pvector=pv, pe=pe.value, ke=ke.value)

    # ---- fix reax/c/bonds / fix reax/c/species ----
    def bond_table(self, bo_cut=-1.0):
        """Connection table of the local atoms (fix_reaxc_bonds_sunway.cpp:187-260); bo_cut < 0 = control bg_cut."""
        n = C.c_int(); m = C.c_int(); mx = C.c_int()
        self._chk(self.lib.rxb_bond_table(self.h, C.c_double(bo_cut), C.byref(n), C.byref(m), C.byref(mx)))
        n, m = n.value, m.value
        t = dict(tag=np.zeros(n, dtype=np.int32), type=np.zeros(n, dtype=np.int32), off=np.zeros(n + 1, dtype=np.int32),
                 nbr=np.zeros(m, dtype=np.int32), bo=np.zeros(m), abo=np.zeros(n), nlp=np.zeros(n), q=np.zeros(n),
                 max_nb=mx.value)
        self._chk(self.lib.rxb_bond_table_get(self.h, _p(t["tag"]), _p(t["type"]), _p(t["off"]), _p(t["nbr"]), _p(t["bo"]),
                                              _p(t["abo"]), _p(t["nlp"]), _p(t["q"])))
        return t

    def spec_atom_abo(self):
        """compute SPEC/ATOM abo01..abo12: one sample [nlocal][12] of the bond orders FindBond would store."""
        a = np.zeros((int(self.counts()[0]), 12))
        self._chk(self.lib.rxb_spec_atom_abo(self.h, _p(a)))
        return a

    def species_config(self, nevery, nrepeat, nfreq, natoms, ntypes=4, bocut=None, ntimestep=-1):
        """fix reax/c/species nevery nrepeat nfreq; bocut = (ntypes+1)^2 BOCut matrix (default 0.30).  Returns True
        when the reneighbouring period of the resident run was reset (the reference's warning)."""
        bc = np.full((ntypes + 1, ntypes + 1), 0.30) if bocut is None else np.ascontiguousarray(bocut, dtype=np.float64)
        self._sp_ntypes = ntypes
        r = C.c_int()
        self._chk(self.lib.rxb_species_config(self.h, int(nevery), int(nrepeat), int(nfreq), int(ntypes), _p(bc),
                                              C.c_long(int(natoms)), C.c_long(int(ntimestep)), C.byref(r)))
        return bool(r.value)

    def species_step(self, ntimestep):
        f = C.c_int()
        self._chk(self.lib.rxb_species_step(self.h, C.c_long(int(ntimestep)), C.byref(f)))
        return bool(f.value)

    def species_result(self):
        nm = C.c_int()
        self._chk(self.lib.rxb_species_result(self.h, C.byref(nm), None, C.c_long(0)))
        comp = np.zeros((nm.value, self._sp_ntypes), dtype=np.int32)
        self._chk(self.lib.rxb_species_result(self.h, C.byref(nm), _p(comp), C.c_long(comp.size)))
        return dict(nmole=nm.value, composition=comp)

    def species_cluster(self):
        c = np.zeros(int(self.counts()[0]), dtype=np.int32)
        self._chk(self.lib.rxb_species_cluster(self.h, _p(c)))
        return c

    def species_avg_qxyz(self):
        """Averaged q, x, y, z columns of the window that just ended ([nlocal][4]); input of the `position` output."""
        a = np.zeros((int(self.counts()[0]), 4))
        self._chk(self.lib.rxb_species_avg_qxyz(self.h, _p(a)))
        return a

    def species_log(self):
        out = []
        for k in range(self.lib.rxb_species_log_size(self.h)):
            st = C.c_long(); nm = C.c_int()
            self._chk(self.lib.rxb_species_log_get(self.h, k, C.byref(st), C.byref(nm), None, C.c_long(0)))
            comp = np.zeros((nm.value, self._sp_ntypes), dtype=np.int32)
            self._chk(self.lib.rxb_species_log_get(self.h, k, C.byref(st), C.byref(nm), _p(comp), C.c_long(comp.size)))
            out.append(dict(step=st.value, nmole=nm.value, composition=comp))
        return out

    # ---- multi-GPU ----
    @staticmethod
    def dist_unique_id():
        lib = load_library()
        buf = C.create_string_buffer(128)
        if lib.rxb_dist_unique_id(buf) != 0:
            raise RxbError(lib.rxb_last_error().decode())
        return buf.raw

    def dist_init(self, rank, world, uid, grid):
        assert len(uid) == 128
        self._chk(self.lib.rxb_dist_init(self.h, int(rank), int(world), C.c_char_p(uid), int(grid[0]), int(grid[1]), int(grid[2])))

    def dist_set_p2p(self, on):
        self._chk(self.lib.rxb_dist_set_p2p(self.h, int(on)))

    def comm_init(self, rank, world, uid):
        """Host-planned halo (a multi-rank LAMMPS drives one handle per rank through the plugin calls)."""
        assert len(uid) == 128
        self._chk(self.lib.rxb_comm_init(self.h, int(rank), int(world), C.c_char_p(uid)))

    def comm_set_ghosts(self, owner_rank, owner_index):
        """Collective, after every set_atoms and before neigh_build: per ghost, the owning rank and the local index there."""
        r = _i(owner_rank); k = _i(owner_index)
        assert len(r) == len(k)
        self._chk(self.lib.rxb_comm_set_ghosts(self.h, len(r), _p(r), _p(k)))

    # ---- introspection ----
    def cutoffs(self):
        c = np.zeros(3)
        self._chk(self.lib.rxb_get_cutoffs(self.h, _p(c)))
        return dict(verlet=c[0], bond_candidates=c[1], bond_reach=c[2])

    def counts(self):
        c = np.zeros(8, dtype=np.int64)
        self._chk(self.lib.rxb_get_counts(self.h, _p(c)))
        return c

    def neighbors(self, which=0):
        c = self.counts()
        nrows = int(c[0]) if which == 0 else int(c[1])
        nnz = int(c[2]) if which == 0 else int(c[3])
        off = np.zeros(nrows + 1, dtype=np.int64); idx = np.zeros(max(nnz, 1), dtype=np.int32)
        self._chk(self.lib.rxb_get_neighbors(self.h, which, _p(off), _p(idx)))
        return off, idx[:nnz]

    def bonds(self):
        c = self.counts()
        N, nb = int(c[1]), int(c[4])
        bs = np.zeros(N, dtype=np.int32); bc = np.zeros(N, dtype=np.int32)
        nbr = np.zeros(max(nb, 1), dtype=np.int32); sym = np.zeros(max(nb, 1), dtype=np.int32)
        fld = np.zeros((max(nb, 1), 31))
        self._chk(self.lib.rxb_get_bonds(self.h, _p(bs), _p(bc), _p(nbr), _p(sym), _p(fld)))
        return bs, bc, nbr[:nb], sym[:nb], fld[:nb]

    def counters(self):
        """dict: spmv_active (launches that really multiplied), qeq_replays, qeq_iterations, kernel_launches."""
        o = np.zeros(4, dtype=np.int64)
        self._chk(self.lib.rxb_get_counters(self.h, _p(o)))
        return dict(zip(["spmv_active", "qeq_replays", "qeq_iterations", "kernel_launches"], o.tolist()))

    def host_syncs(self):
        """Host-side waits on a CUDA stream / event issued by the library in this process so far."""
        self.lib.rxb_host_sync_count.restype = C.c_longlong
        return int(self.lib.rxb_host_sync_count())

    def hbond_pairs(self):
        n = C.c_int()
        self._chk(self.lib.rxb_get_hbond_pairs(self.h, C.byref(n), None, 0))
        p = np.zeros((max(n.value, 1), 2), dtype=np.int32)
        self._chk(self.lib.rxb_get_hbond_pairs(self.h, C.byref(n), _p(p), n.value))
        return p[:n.value]

    def workspace(self):
        N = int(self.counts()[1])
        w = np.zeros((N, 16))
        self._chk(self.lib.rxb_get_workspace(self.h, _p(w)))
        return w

    def far(self):
        c = self.counts()
        n, nnz = int(c[0]), int(c[2])
        num = np.zeros(n, dtype=np.int32); idx = np.zeros(max(nnz, 1), dtype=np.int32); val = np.zeros(max(nnz, 1))
        self._chk(self.lib.rxb_get_far(self.h, _p(num), _p(idx), _p(val)))
        return num, idx, val

    def profiler_range(self, start):
        self._chk(self.lib.rxb_profiler_range(int(start)))

    PHASES = ["neigh", "qeq_farH", "qeq_cg", "bond_list", "bond_orders", "bonded", "nonbonded", "dbond", "spmv",
              "hbond_items", "angle_torsion_items", "multi_body", "enum", "spmv_boundary"]

    def profile(self, enable=None):
        """-> dict phase -> (total ms, calls) accumulated since the last profile(1).  "spmv" = whole SpMVs: in multi-GPU
        runs a SpMV is two launches (interior rows while the halo is in flight, then boundary rows); the time of the
        second ("spmv_boundary") is added to "spmv" here, the call count is the number of SpMVs."""
        nph = len(self.PHASES)
        out = np.zeros(2 * nph)
        self._chk(self.lib.rxb_profile(self.h, -1 if enable is None else int(enable), _p(out)))
        d = {nm: (out[k], int(out[nph + k])) for k, nm in enumerate(self.PHASES)}
        d["spmv"] = (d["spmv"][0] + d["spmv_boundary"][0], d["spmv"][1])
        return d

    def h_format(self):
        b = C.c_int(); nm = C.create_string_buffer(128)
        self._chk(self.lib.rxb_get_h_format(self.h, C.byref(b), nm, 128))
        return {"bytes_per_entry": b.value, "name": nm.value.decode()}

    @staticmethod
    def measure_fp64_tflops(device=0):
        return float(load_library().rxb_measure_fp64_tflops(int(device)))

    def md_last_run_ms(self):
        return float(self.lib.rxb_md_last_run_ms(self.h))
