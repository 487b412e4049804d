"""Test infrastructure: the part of the LAMMPS core that a MULTI-RANK run adds on the host side of the plugin boundary.

A real LAMMPS keeps its own spatial decomposition (Domain/Comm: bricks in lamda space), migrates atoms (exchange), builds
ghost atoms (borders), forward-communicates positions, reverse-communicates forces and all-reduces the thermo sums; the
reference's pair style and fix only ever see "my local atoms + my ghosts" (pair_reaxc_sunway.cpp:560-640,
fix_qeq_reax_sunway.cpp:1043-1140).  `HostComm` reproduces that contract with numpy for any px*py*pz grid, including the
ghost ORDER being the host's own (shuffled here on purpose), and `host_md` is the run loop (velocity Verlet, reneighbouring
every `every` steps with migration of the QEq history, like fix qeq/reax's pack_exchange) that drives one library handle
per rank exclusively through the plugin calls of include/rxb200.h:

    rxb_set_atoms -> rxb_comm_set_ghosts -> rxb_neigh_build          (reneighbouring steps)
    rxb_set_positions -> rxb_qeq_pre_force -> rxb_pair_compute       (every step)

With world == 1 and comm=False the same loop drives a plain single-rank handle (ghost_owner path): the trajectory the
multi-rank run must reproduce.  The global state is replicated on every rank (the systems are a few thousand atoms); the
only host communication is the gather of per-rank results (`allgather`, a torch.distributed gloo all_gather_object in the
N > 1 script, the identity for one rank).
"""
import itertools

import numpy as np

KB = 0.0019872067
MVV2E = 48.88821291 ** 2
FTM2V = 1.0 / MVV2E


def box_h(box6):
    lx, ly, lz, xy, xz, yz = [float(b) for b in box6]
    return np.array([[lx, xy, xz], [0.0, ly, yz], [0.0, 0.0, lz]])   # x = h @ lamda


class HostComm:
    def __init__(self, box6, grid, cut):
        self.h = box_h(box6)
        self.hinv = np.linalg.inv(self.h)
        self.grid = tuple(int(g) for g in grid)
        self.world = self.grid[0] * self.grid[1] * self.grid[2]
        self.cg = cut * np.linalg.norm(self.hinv, axis=1)             # ghost cut-off in lamda units, per dimension
        self.m = np.ceil(self.cg).astype(int)                        # periodic image range per dimension

    def lamda(self, x):
        return x @ self.hinv.T

    def wrap(self, x):
        l = self.lamda(x)
        l -= np.floor(l)
        l[l >= 1.0] = 0.0
        return l @ self.h.T

    def assign(self, x):
        l = self.lamda(x)
        c = [np.clip(np.floor(l[:, d] * self.grid[d]).astype(int), 0, self.grid[d] - 1) for d in range(3)]
        return c[0] + self.grid[0] * (c[1] + self.grid[1] * c[2])

    def brick(self, rank):
        gx, gy, gz = self.grid
        c = (rank % gx, (rank // gx) % gy, rank // (gx * gy))
        lo = np.array([c[d] / self.grid[d] for d in range(3)])
        hi = np.array([(c[d] + 1) / self.grid[d] for d in range(3)])
        return lo, hi

    def borders(self, x, owner, rank, seed=None):
        """-> (local: global indices of my atoms, src: global index behind each ghost, shift: its image vector (Cartesian)).
        Ghosts = every periodic image of every atom inside my brick extended by the ghost cut-off, except my own atoms
        unshifted - the set LAMMPS' multi-hop swaps produce.  seed: shuffle the ghost order (the host's order is its own)."""
        local = np.nonzero(owner == rank)[0]
        lo, hi = self.brick(rank)
        l = self.lamda(x)
        src, shift = [], []
        for s in itertools.product(*[range(-m, m + 1) for m in self.m]):
            ls = l + np.array(s, dtype=float)
            inside = np.all((ls >= lo - self.cg) & (ls < hi + self.cg), axis=1)
            if s == (0, 0, 0):
                inside &= owner != rank
            k = np.nonzero(inside)[0]
            src.append(k)
            shift.append(np.tile(self.h @ np.array(s, dtype=float), (len(k), 1)))
        src = np.concatenate(src)
        shift = np.concatenate(shift) if len(src) else np.zeros((0, 3))
        if seed is not None and len(src):
            p = np.random.default_rng(seed).permutation(len(src))
            src, shift = src[p], shift[p]
        return local, src, shift


def host_md(Rxb, H, comm, rank, device, box6, x, v, types, tags, steps, every=4, dt=0.25, tol=1e-8, uid=None, use_comm=True,
            allgather=lambda o: [o], shuffle=True, exact_h=False, async_qeq=False):
    """Velocity-Verlet NVE of the replicated global system, forces from one handle per rank through the plugin calls.
    Returns dict(pe[steps+1], ke[...], x, v, q (global, by atom), f (global), matvecs[...], nghost, peer)."""
    r = Rxb(device)
    r.pair_settings(H.CONTROL)
    r.pair_coeff(H.FFIELD, H.ELEMENTS)
    r.fix_qeq(0.0, 10.0, tol)
    if exact_h:
        r.set_h_exact(True)
    if use_comm:
        r.comm_init(rank, comm.world, uid)
    natoms = len(x)
    x = comm.wrap(np.array(x, dtype=float)); v = np.array(v, dtype=float)
    mass = H.MASS[types]
    s_hist = np.zeros((natoms, 5)); t_hist = np.zeros((natoms, 5))
    have_hist = False
    state = {}

    def reneighbour(k):
        nonlocal x, have_hist
        if have_hist:       # the QEq history travels with the atoms (fix_qeq_reax pack_exchange / unpack_exchange)
            s, t = r.qeq_get_history()
            for loc, ss, tt in allgather((state["local"], s, t)):
                s_hist[loc] = ss; t_hist[loc] = tt
        x = comm.wrap(x)
        owner = comm.assign(x)
        local, src, shift = comm.borders(x, owner, rank, seed=(1000 * k + rank) if shuffle else None)
        local_index = np.empty(natoms, dtype=np.int64)
        for rr in range(comm.world):
            loc = np.nonzero(owner == rr)[0]
            local_index[loc] = np.arange(len(loc))
        idx = np.concatenate([local, src])
        state.update(local=local, src=src, shift=shift, idx=idx, owner=owner)
        xa = np.concatenate([x[local], x[src] + shift])
        if use_comm:
            r.set_atoms(len(local), xa, types[idx], tags[idx], q=state.get("q", np.zeros(natoms))[idx])
            r.comm_set_ghosts(owner[src], local_index[src])
        else:
            r.set_atoms(len(local), xa, types[idx], tags[idx], q=state.get("q", np.zeros(natoms))[idx],
                        ghost_owner=local_index[src])
        r.neigh_build()
        if have_hist:
            r.qeq_set_history(s_hist[local], t_hist[local])

    def forces(first):
        local, src, shift, idx = state["local"], state["src"], state["shift"], state["idx"]
        if not first:
            r.set_positions(np.concatenate([x[local], x[src] + shift]))
        if async_qeq:                  # enqueue only; the solve is settled inside pair_compute (non-thermo steps of the host styles)
            r.qeq_pre_force_async()
            out = r.pair_compute(True, True)
            mv = r.qeq_matvecs()
        else:
            mv = r.qeq_pre_force()
            out = r.pair_compute(True, True)
        q = r.get_charges()
        f = np.zeros((natoms, 3)); qg = np.zeros(natoms); pe = 0.0
        # reverse communication of the forces (local + ghost contributions summed onto the real atom) and thermo sums
        for gi, fr, loc, ql, e in allgather((idx, out["f"], local, q[:len(local)], float(out["eng"].sum()))):
            np.add.at(f, gi, fr)
            qg[loc] = ql
            pe += e
        # ghost charges must equal their owners' (forward communication of q at the end of pre_force)
        dq_ghost = float(np.abs(q[len(local):] - qg[src]).max()) if len(src) else 0.0
        state["q"] = qg
        return f, pe, qg, mv, dq_ghost

    pe_t, ke_t, mv_t, dqg = [], [], [], 0.0
    reneighbour(0)
    f, pe, q, mv, d = forces(True)
    have_hist = True
    dqg = max(dqg, d)

    def ke_of(v):
        return 0.5 * MVV2E * float((mass[:, None] * v * v).sum())
    pe_t.append(pe); ke_t.append(ke_of(v)); mv_t.append(mv)
    for step in range(1, steps + 1):
        v += 0.5 * dt * FTM2V * f / mass[:, None]
        x += dt * v
        if step % every == 0:
            reneighbour(step)
            f, pe, q, mv, d = forces(True)
        else:
            f, pe, q, mv, d = forces(False)
        dqg = max(dqg, d)
        v += 0.5 * dt * FTM2V * f / mass[:, None]
        pe_t.append(pe); ke_t.append(ke_of(v)); mv_t.append(mv)
    owner0 = state["owner"]
    res = dict(pe=np.array(pe_t), ke=np.array(ke_t), x=x.copy(), v=v.copy(), q=q, f=f, matvecs=mv_t, nghost=len(state["src"]),
               nlocal=len(state["local"]), ghost_q_err=dqg, owner=owner0, h_format=r.h_format())
    del r
    return res


def compare(a, b):
    """a = multi-rank / comm run, b = the plain single-rank run of the same trajectory."""
    fs = max(float(np.abs(b["f"]).max()), 1e-300)
    return dict(pe_rel=float(np.abs((a["pe"] - b["pe"]) / b["pe"]).max()),
                ke_rel=float(np.abs((a["ke"] - b["ke"]) / np.maximum(np.abs(b["ke"]), 1e-300)).max()),
                dx=float(np.abs(a["x"] - b["x"]).max()), f_rel=float(np.abs(a["f"] - b["f"]).max() / fs),
                dq=float(np.abs(a["q"] - b["q"]).max()), ghost_q_err=a["ghost_q_err"],
                matvecs_equal=bool(a["matvecs"] == b["matvecs"]))
