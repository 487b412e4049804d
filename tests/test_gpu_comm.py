"""Host-planned halo (rxb_comm_init / rxb_comm_set_ghosts): the multi-rank LAMMPS boundary.  On one GPU the communicator
has one rank and every ghost is an image of an own atom, but the whole path is the multi-rank one (exchange plan, ghost
permutation, NCCL/peer transport, q forward); the N > 1 run is tests/gpu_comm_check.py under torchrun."""
import numpy as np
import pytest

import helpers as H
import lammps_comm as LC


def test_host_comm_ghost_shell_is_complete():
    """The stand-in's borders(): every atom image within the cut-off of a local atom is local or a ghost (brute force)."""
    box, x, t, tag = H.tatb_cell(1, 1, 1)
    cut = 6.0
    for grid in ((1, 1, 1), (2, 1, 1), (1, 2, 2)):
        comm = LC.HostComm(box, grid, cut)
        xw = comm.wrap(x)
        owner = comm.assign(xw)
        assert sorted(np.unique(owner).tolist()) == list(range(comm.world))
        hmat = LC.box_h(box)
        # brute-force neighbour counts over periodic images
        shifts = np.array([hmat @ np.array(s, float) for s in np.ndindex(5, 5, 5)]) - hmat @ np.array([2.0, 2.0, 2.0])
        for rank in range(comm.world):
            local, src, shift = comm.borders(xw, owner, rank, seed=7)
            xa = np.concatenate([xw[local], xw[src] + shift])
            for i in local[:: max(1, len(local) // 25)]:
                d_all = np.linalg.norm((xw[None, :, :] + shifts[:, None, :]) - xw[i], axis=2)
                want = int((d_all <= cut).sum())
                got = int((np.linalg.norm(xa - xw[i], axis=1) <= cut).sum())
                assert got == want, (grid, rank, i, got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("peer,async_qeq", [("1", False), ("0", False), ("1", True)])
def test_comm_mode_one_rank_matches_plain_plugin_run(peer, async_qeq, monkeypatch):
    from sw_reaxff_b200 import Rxb
    monkeypatch.setenv("RXB_PEER", peer)
    box, x, t, tag = H.tatb_cell(2, 2, 2)
    v = H.maxwell_velocities(t, 1500.0, 4242)
    comm = LC.HostComm(box, (1, 1, 1), 12.5)
    a = LC.host_md(Rxb, H, comm, 0, 0, box, x, v, t, tag, 9, uid=Rxb.dist_unique_id(), use_comm=True, tol=1e-10, async_qeq=async_qeq)
    b = LC.host_md(Rxb, H, comm, 0, 0, box, x, v, t, tag, 9, use_comm=False, shuffle=False, tol=1e-10)
    c = LC.compare(a, b)
    assert c["ghost_q_err"] == 0.0
    assert c["pe_rel"] < 1e-11 and c["ke_rel"] < 1e-9, c
    assert c["f_rel"] < 1e-9 and c["dq"] < 1e-9 and c["dx"] < 1e-10, c
    assert all(abs(p[0] - q[0]) <= 2 and abs(p[1] - q[1]) <= 2 for p, q in zip(a["matvecs"], b["matvecs"])), (a["matvecs"], b["matvecs"])


@pytest.mark.gpu
def test_comm_mode_argument_errors():
    from sw_reaxff_b200 import Rxb, RxbError
    box, x, t, tag = H.tatb_cell(1, 1, 1)
    comm = LC.HostComm(box, (1, 1, 1), 12.5)
    xw = comm.wrap(x)
    owner = comm.assign(xw)
    local, src, shift = comm.borders(xw, owner, 0)
    idx = np.concatenate([local, src])
    xa = np.concatenate([xw[local], xw[src] + shift])
    r = Rxb(0)
    r.pair_settings(H.CONTROL); r.pair_coeff(H.FFIELD, H.ELEMENTS); r.fix_qeq(0.0, 10.0, 1e-6)
    with pytest.raises(RxbError, match="rxb_comm_init first"):
        r.comm_set_ghosts(np.zeros(3, np.int32), np.zeros(3, np.int32))
    r.comm_init(0, 1, Rxb.dist_unique_id())
    with pytest.raises(RxbError, match="already belongs"):
        r.comm_init(0, 1, Rxb.dist_unique_id())
    r.set_atoms(len(local), xa, t[idx], tag[idx])
    with pytest.raises(RxbError, match="nghost"):
        r.comm_set_ghosts(np.zeros(3, np.int32), np.zeros(3, np.int32))
    bad = np.zeros(len(src), np.int32); bad[0] = 5
    with pytest.raises(RxbError, match="outside"):
        r.comm_set_ghosts(bad, src)
    with pytest.raises(RxbError, match="not a local atom"):
        r.comm_set_ghosts(np.zeros(len(src), np.int32), np.full(len(src), len(local), np.int32))
    with pytest.raises(RxbError, match="must follow every rxb_set_atoms"):
        r.neigh_build()                                  # no plan yet for this atom set
    r.comm_set_ghosts(np.zeros(len(src), np.int32), src)  # (1 rank: local index == global index)
    r.neigh_build()
    r.qeq_pre_force()
    r.pair_compute()
    with pytest.raises(RxbError, match="host-planned"):
        r.md_setup(box, x, np.zeros_like(x), t, tag, H.MASS)
