"""Test helpers: ctypes wrapper of the CPU oracle (oracle/liboracle.so) and TATB system builders.

The oracle is test infrastructure: it is only ever imported from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(ROOT, "sw_reaxff_b200", "data", "tatb")
FFIELD = os.path.join(DATA, "ffield.reax")
CONTROL = os.path.join(DATA, "control.reax_c.tatb")
DATAFILE = os.path.join(DATA, "data.tatb")
ELEMENTS = ["C", "H", "O", "N"]
MASS = np.array([0.0, 12.0, 1.008, 15.999, 14.0])

E_NAMES = ["e_bond", "e_ov", "e_un", "e_lp", "e_ang", "e_pen", "e_coa", "e_hb", "e_tor", "e_con", "e_vdW", "e_ele", "e_pol"]

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_lp = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"])


def read_data_tatb(path=DATAFILE):
    """-> box6 (xprd,yprd,zprd,xy,xz,yz), x[384,3], type[384] (1-based), tag[384]"""
    with open(path) as f:
        lines = f.read().splitlines()
    lo_hi = {}
    tilt = None
    natoms = None
    i = 0
    atoms_at = None
    for i, ln in enumerate(lines):
        t = ln.split()
        if len(t) >= 2 and t[1] == "atoms":
            natoms = int(t[0])
        if len(t) >= 4 and t[2] in ("xlo", "ylo", "zlo"):
            lo_hi[t[2][0]] = (float(t[0]), float(t[1].rstrip("E")))
        if len(t) >= 6 and t[3] == "xy":
            tilt = (float(t[0]), float(t[1]), float(t[2]))
        if t and t[0] == "Atoms":
            atoms_at = i
    rows = []
    for ln in lines[atoms_at + 1:]:
        t = ln.split()
        if len(t) >= 6:
            rows.append([float(v) for v in t[:6]])
        if len(rows) == natoms:
            break
    a = np.array(rows)
    order = np.argsort(a[:, 0])
    a = a[order]
    box6 = np.array([lo_hi["x"][1] - lo_hi["x"][0], lo_hi["y"][1] - lo_hi["y"][0], lo_hi["z"][1] - lo_hi["z"][0],
                     tilt[0], tilt[1], tilt[2]])
    return box6, np.ascontiguousarray(a[:, 3:6]), a[:, 1].astype(np.int32), a[:, 0].astype(np.int32)


def tatb_cell(nx=1, ny=1, nz=1, perturb=0.0, seed=0, scale=1.0):
    """Replicate the 384-atom TATB cell (deterministic lattice translation)."""
    box6, x0, t0, _ = read_data_tatb()
    a = np.array([box6[0], 0, 0]); b = np.array([box6[3], box6[1], 0]); c = np.array([box6[4], box6[5], box6[2]])
    xs, ts = [], []
    for iz in range(nz):
        for iy in range(ny):
            for ix in range(nx):
                xs.append(x0 + ix * a + iy * b + iz * c)
                ts.append(t0)
    x = np.concatenate(xs)
    t = np.concatenate(ts).astype(np.int32)
    box = np.array([box6[0] * nx, box6[1] * ny, box6[2] * nz, box6[3] * ny, box6[4] * nz, box6[5] * nz])
    if perturb > 0:
        rng = np.random.default_rng(seed)
        x = x + rng.uniform(-perturb, perturb, size=x.shape)
    if scale != 1.0:
        x = x * scale
        box = box * scale
    tag = np.arange(1, len(x) + 1, dtype=np.int32)
    return box, np.ascontiguousarray(x), t, tag


def maxwell_velocities(types, T, seed):
    """Gaussian velocities at temperature T (K), units real (A/fs), zero net momentum."""
    rng = np.random.default_rng(seed)
    kB = 0.0019872067
    mvv2e = 48.88821291 ** 2
    m = MASS[types]
    v = rng.normal(size=(len(types), 3)) * np.sqrt(kB * T / (m * mvv2e))[:, None]
    p = (v * m[:, None]).sum(0) / m.sum()
    v -= p
    ke = 0.5 * mvv2e * (m[:, None] * v * v).sum()
    tcur = 2 * ke / (3 * (len(types) - 1) * kB)
    v *= np.sqrt(T / tcur)
    return np.ascontiguousarray(v)


class Oracle:
    def __init__(self, ffield=FFIELD, control=CONTROL, elements=ELEMENTS, omp=False, lgflag=0, enobonds=1):
        name = "liboracle_omp.so" if omp else "liboracle.so"
        path = os.path.join(ROOT, "oracle", name)
        if not os.path.exists(path):
            build_oracle()
        L = self.L = C.CDLL(path)
        L.orc_create.restype = C.c_void_p
        L.orc_params_dump.restype = C.c_long
        L.orc_num_neighbors.restype = C.c_long
        L.orc_qeq_get_H.restype = C.c_long
        err = C.create_string_buffer(512)
        arr = (C.c_char_p * len(elements))(*[e.encode() for e in elements])
        self.h = C.c_void_p(L.orc_create(ffield.encode(), control.encode() if control else None, len(elements), arr,
                                         lgflag, enobonds, err, 512))
        if not self.h:
            raise RuntimeError("oracle: " + err.value.decode())
        self.n = self.N = 0

    def __del__(self):
        try:
            if self.h:
                self.L.orc_destroy(self.h)
        except Exception:
            pass

    def params_dump(self):
        n = self.L.orc_params_dump(self.h, None, C.c_long(0))
        out = np.zeros(n)
        self.L.orc_params_dump(self.h, out.ctypes.data_as(C.c_void_p), C.c_long(n))
        return out

    # ---- static configuration ----
    def set_atoms(self, n, x, ltype, tag, q):
        N = len(x)
        self.n, self.N = n, N
        self.L.orc_set_atoms(self.h, n, N, _c(x), _ci(ltype), _ci(tag), _c(q))

    def build_neighbors(self, cutneigh=12.5):
        self.L.orc_build_neighbors(self.h, C.c_double(cutneigh))

    def get_neighbors(self):
        nn = self.L.orc_num_neighbors(self.h)
        off = np.zeros(self.N + 1, dtype=np.int64)
        nb = np.zeros(nn, dtype=np.int32)
        self.L.orc_get_neighbors(self.h, off.ctypes.data_as(C.c_void_p), nb.ctypes.data_as(C.c_void_p))
        return off, nb

    def compute(self):
        self.L.orc_compute(self.h)

    def phase(self, which):
        self.L.orc_phase(self.h, which)

    def forces(self):
        f = np.zeros((self.N, 3))
        self.L.orc_get_forces(self.h, f.ctypes.data_as(C.c_void_p))
        return f

    def cddelta(self):
        c = np.zeros(self.N)
        self.L.orc_get_cddelta(self.h, c.ctypes.data_as(C.c_void_p))
        return c

    def energies(self):
        e = np.zeros(13)
        v = np.zeros(6)
        self.L.orc_get_energies(self.h, e.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p))
        return e, v

    def bonds(self):
        nb = self.L.orc_num_bonds(self.h)
        bs = np.zeros(self.N, dtype=np.int32); be = np.zeros(self.N, dtype=np.int32)
        nbr = np.zeros(nb, dtype=np.int32); sym = np.zeros(nb, dtype=np.int32)
        fld = np.zeros((nb, 31))
        self.L.orc_get_bonds(self.h, _p(bs), _p(be), _p(nbr), _p(sym), _p(fld))
        return bs, be, nbr, sym, fld

    def workspace(self):
        w = np.zeros((self.N, 16))
        self.L.orc_get_workspace(self.h, _p(w))
        return w

    def ddeltap_self(self):
        d = np.zeros((self.N, 3))
        self.L.orc_get_ddeltap_self(self.h, _p(d))
        return d

    def hbonds(self):
        nh = self.L.orc_num_hbonds(self.h)
        Hindex = np.zeros(self.N, dtype=np.int32)
        numH = self.n
        hs = np.zeros(numH, dtype=np.int32); he = np.zeros(numH, dtype=np.int32)
        nbr = np.zeros(nh, dtype=np.int32)
        self.L.orc_get_hbonds(self.h, _p(Hindex), _p(hs), _p(he), _p(nbr))
        return Hindex, hs, he, nbr

    # ---- QEq ----
    def qeq_init(self, swa=0.0, swb=10.0, tol=1e-6):
        self.L.orc_qeq_init(self.h, C.c_double(swa), C.c_double(swb), C.c_double(tol))

    def qeq_set_hist(self, s_hist, t_hist):
        self.L.orc_qeq_set_hist(self.h, _c(s_hist), _c(t_hist))

    def qeq_get_hist(self):
        s = np.zeros((self.n, 5)); t = np.zeros((self.n, 5))
        self.L.orc_qeq_get_hist(self.h, _p(s), _p(t))
        return s, t

    def qeq_pre_force(self, ghost_owner):
        mv = np.zeros(2, dtype=np.int32)
        self.L.orc_qeq_pre_force(self.h, _ci(ghost_owner), _p(mv))
        return int(mv[0]), int(mv[1])

    def q(self):
        q = np.zeros(self.N)
        self.L.orc_get_q(self.h, _p(q))
        return q

    def qeq_st(self):
        s = np.zeros(self.N); t = np.zeros(self.N)
        self.L.orc_qeq_get_st(self.h, _p(s), _p(t))
        return s, t

    def qeq_H(self):
        nnz = self.L.orc_qeq_get_H(self.h, None, None, None, None)
        off = np.zeros(self.n + 1, dtype=np.int64); num = np.zeros(self.n, dtype=np.int32)
        col = np.zeros(nnz, dtype=np.int32); val = np.zeros(nnz)
        self.L.orc_qeq_get_H(self.h, _p(off), _p(num), _p(col), _p(val))
        return off, num, col, val

    # ---- mini MD ----
    def md_init(self, box6, x, v, ltype, tag, dt=0.0625, skin=2.5, every=5, qeq=True, qeq_tol=1e-6, mass=MASS):
        self.nlocal = len(x)
        self.L.orc_md_init(self.h, _c(box6), len(x), _c(x), _c(v), _ci(ltype), _ci(tag), _c(mass), len(mass) - 1,
                           C.c_double(dt), C.c_double(skin), every, int(qeq), C.c_double(qeq_tol))
        self.n = self.nlocal
        self.N = self.L.orc_md_nall(self.h)

    def md_run(self, nsteps):
        self.L.orc_md_run(self.h, nsteps)
        self.N = self.L.orc_md_nall(self.h)

    def md_get(self):
        n = self.nlocal
        x = np.zeros((n, 3)); v = np.zeros((n, 3)); f = np.zeros((n, 3)); q = np.zeros(n); e = np.zeros(13); pk = np.zeros(2)
        self.L.orc_md_get(self.h, _p(x), _p(v), _p(f), _p(q), _p(e), _p(pk))
        return dict(x=x, v=v, f=f, q=q, e=e, pe=pk[0], ke=pk[1])

    def md_ghosts(self):
        N = self.L.orc_md_nall(self.h)
        xall = np.zeros((N, 3)); ty = np.zeros(N, dtype=np.int32); tg = np.zeros(N, dtype=np.int32)
        owner = np.zeros(N - self.nlocal, dtype=np.int32)
        self.L.orc_md_get_ghosts(self.h, _p(xall), _p(ty), _p(tg), _p(owner))
        return xall, ty, tg, owner

    def md_matvecs(self):
        return self.L.orc_md_matvecs(self.h, 0), self.L.orc_md_matvecs(self.h, 1)

    def omp_threads(self, n=0):
        """Set (n > 0) and return the OpenMP thread count the oracle library really runs with (torchrun exports
        OMP_NUM_THREADS=1, so the environment variable alone is not to be trusted)."""
        return int(self.L.orc_omp_threads(int(n)))

    # ---- fix reax/c/bonds, fix reax/c/species ----
    def _text(self, fn, step):
        self.L[fn].restype = C.c_long
        need = self.L[fn](self.h, C.c_long(step), None, C.c_long(0))
        buf = C.create_string_buffer(need + 1)
        self.L[fn](self.h, C.c_long(step), buf, C.c_long(need + 1))
        return buf.value.decode()

    def md_bonds_text(self, step):
        return self._text("orc_md_bonds_text", step)

    def md_species_init(self, nevery, nrepeat, nfreq, ntypes=4, bocut=None):
        bc = np.full((ntypes + 1, ntypes + 1), 0.30) if bocut is None else np.ascontiguousarray(bocut, dtype=np.float64)
        self._sp_ntypes = ntypes
        self.L.orc_md_species_init(self.h, nevery, nrepeat, nfreq, _c(bc))

    def md_species_step(self, step):
        r = self.L.orc_md_species_step(self.h, C.c_long(step))
        assert r >= 0, "oracle species error"
        return r == 1

    def md_species_get(self):
        nm = self.L.orc_md_species_nmole(self.h)
        comp = np.zeros((nm, self._sp_ntypes), dtype=np.int32); cl = np.zeros(self.nlocal, dtype=np.int32)
        self.L.orc_md_species_get(self.h, _p(comp), _p(cl))
        return dict(nmole=nm, composition=comp, cluster=cl)

    def md_species_raw(self):
        ids = np.zeros((self.nlocal, 12), dtype=np.int32); avg = np.zeros((self.nlocal, 12))
        self.L.orc_md_species_raw(self.h, _p(ids), _p(avg))
        return ids, avg

    def md_species_text(self, step):
        return self._text("orc_md_species_text", step)

    def md_species_pos(self, step, box6):
        """-> (text of the `position` file block, averaged q/x/y/z columns [nlocal][4]); box6 = boxlo[3] + boxhi[3]."""
        fn = self.L.orc_md_species_pos_text
        fn.restype = C.c_long
        b = np.ascontiguousarray(box6, dtype=np.float64)
        avg = np.zeros((self.nlocal, 4))
        need = fn(self.h, C.c_long(step), _p(b), _p(avg), None, C.c_long(0))
        buf = C.create_string_buffer(need + 1)
        fn(self.h, C.c_long(step), _p(b), None, buf, C.c_long(need + 1))
        return buf.value.decode(), avg


def _c(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    _keep.append(a)
    return a.ctypes.data_as(C.c_void_p)


def _ci(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    _keep.append(a)
    return a.ctypes.data_as(C.c_void_p)


def _p(a):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


_keep = []  # keep temporaries alive across the ctypes call (bounded: cleared opportunistically)


def control_variant(path, tabulate):
    """Copy of the TATB control file with `tabulate_long_range N` (spline-table mode of the long-range terms)."""
    src = open(CONTROL).read().splitlines()
    src = [("tabulate_long_range     %d" % tabulate) if l.startswith("tabulate_long_range") else l for l in src]
    with open(path, "w") as f:
        f.write("\n".join(src) + "\n")
    return str(path)


def ffield_variant(path, vdw_type):
    """Write a copy of the TATB force field whose element block selects another van der Waals form
    (reaxc_ffield_sunway.cpp:240-293): 3 = shielding + inner wall (rcore2/ecore2/acore2 set), 2 = inner wall only
    (gamma_w <= 0.5 as well).  Exercises the branches the shipped force field (type 1) never takes."""
    src = open(FFIELD).read().splitlines()
    ia = next(i for i, l in enumerate(src) if "Nr of atoms" in l)
    nel = int(src[ia].split()[0])
    for e in range(nel):
        l2 = ia + 4 + 4 * e + 1          # alfa; gammavdW; ...
        l4 = ia + 4 + 4 * e + 3          # ov/un; val1; n.u.; val3; vval4; rcore2; ecore2; acore2
        w4 = src[l4].split()
        w4[5:8] = ["%.4f" % (1.2 + 0.15 * e), "%.4f" % (0.08 + 0.02 * e), "%.4f" % (10.0 + e)]
        src[l4] = "     " + "  ".join(w4)
        if vdw_type == 2:
            w2 = src[l2].split()
            w2[1] = "0.4000"
            src[l2] = "     " + "  ".join(w2)
    with open(path, "w") as f:
        f.write("\n".join(src) + "\n")
    return str(path)


def static_config(nx=1, ny=1, nz=1, perturb=0.0, seed=0, scale=1.0, qeq=True, oracle=None):
    """Build (n, xall, typeall, tagall, qall, ghost_owner) exactly as the LAMMPS core would hand it to the pair style,
    using the oracle's mini-MD setup (remap + periodic ghosts) and, optionally, equilibrated charges."""
    o = oracle or Oracle()
    box, x, t, tag = tatb_cell(nx, ny, nz, perturb, seed, scale)
    v = np.zeros_like(x)
    o.md_init(box, x, v, t, tag, qeq=qeq)
    xall, ty, tg, owner = o.md_ghosts()
    q = o.q()
    del _keep[:]
    return dict(n=len(x), box=box, x=xall, type=ty, tag=tg, q=q, owner=owner, oracle=o)
