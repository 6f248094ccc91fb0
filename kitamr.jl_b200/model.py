"""Flat host model == the arrays that cross the C-ABI (include/kamr.h).

`HostMesh` is what a Julia shim would assemble from `ka` after every `amr_recover!`
(src/Solver/AMR.jl:54); here it is assembled from the synthetic forest.  `HostState` holds the
per-point / per-cell fields in the reference's column-major per-cell blocks.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import abi
from .synth.forest import FACE_BACKHANGING, FACE_DOMAIN, FACE_FULL, FACE_HANGING, Forest


@dataclass
class HostMesh:
    dim: int
    ndf: int
    n_local: int
    n_ghost: int
    n_solidnbr: int
    ds: np.ndarray
    mid: np.ndarray
    bound_enc: np.ndarray
    ps_level: np.ndarray
    cell_grid: np.ndarray
    grid_off: np.ndarray
    v_level: np.ndarray
    v_weight: np.ndarray
    v_mid: np.ndarray
    nb_state: np.ndarray
    nb_off: np.ndarray
    nb_ids: np.ndarray
    ps_maxlevel: int
    ps_minlevel: int
    face_kind: np.ndarray
    face_here: np.ndarray
    face_there: np.ndarray
    face_dir: np.ndarray
    face_rot: np.ndarray
    face_mid: np.ndarray
    face_there_mid: np.ndarray
    bc_type: np.ndarray
    bc_prim: np.ndarray
    peer_rank: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    send_off: np.ndarray = field(default_factory=lambda: np.zeros(1, np.int32))
    send_cells: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    recv_off: np.ndarray = field(default_factory=lambda: np.zeros(1, np.int32))
    global_ids: np.ndarray = None      # [n_local + n_ghost] global (forest) cell index
    ib: object = None

    # ------------------------------------------------------------------ sizes
    @property
    def n_cell(self):
        return self.n_local + self.n_ghost + self.n_solidnbr

    @property
    def n_grid(self):
        return len(self.grid_off) - 1

    def cell_n(self):
        return (self.grid_off[1:] - self.grid_off[:-1])[self.cell_grid]

    def vs_off(self):
        return np.concatenate([[0], np.cumsum(self.cell_n())]).astype(np.int64)

    def n_phase_local(self):
        """sum of vs_num over local fluid cells (the reference's 'Total number of phase grids')."""
        sel = self.bound_enc[: self.n_local] >= 0
        return int(self.cell_n()[: self.n_local][sel].sum())

    # ------------------------------------------------------------------ C view
    def c_struct(self):
        m = abi.KamrMesh()
        p = abi.ptr
        m.n_local, m.n_ghost, m.n_solidnbr = self.n_local, self.n_ghost, self.n_solidnbr
        m.ds = p(self.ds, C.c_double); m.mid = p(self.mid, C.c_double)
        m.bound_enc = p(self.bound_enc, C.c_int32); m.ps_level = p(self.ps_level, C.c_int32)
        m.cell_grid = p(self.cell_grid, C.c_int32)
        m.n_grid = self.n_grid
        m.grid_off = p(self.grid_off, C.c_int64); m.v_level = p(self.v_level, C.c_int8)
        m.v_weight = p(self.v_weight, C.c_double); m.v_mid = p(self.v_mid, C.c_double)
        m.nb_state = p(self.nb_state, C.c_int32); m.nb_off = p(self.nb_off, C.c_int32)
        m.nb_ids = p(self.nb_ids, C.c_int32)
        m.ps_maxlevel, m.ps_minlevel = self.ps_maxlevel, self.ps_minlevel
        m.n_face = len(self.face_kind)
        m.face_kind = p(self.face_kind, C.c_int32); m.face_here = p(self.face_here, C.c_int32)
        m.face_there = p(self.face_there, C.c_int32); m.face_dir = p(self.face_dir, C.c_int32)
        m.face_rot = p(self.face_rot, C.c_double); m.face_mid = p(self.face_mid, C.c_double)
        m.face_there_mid = p(self.face_there_mid, C.c_double)
        m.n_bc = len(self.bc_type)
        m.bc_type = p(self.bc_type, C.c_int32); m.bc_prim = p(self.bc_prim, C.c_double)
        m.n_peer = len(self.peer_rank)
        m.peer_rank = p(self.peer_rank, C.c_int32); m.send_off = p(self.send_off, C.c_int32)
        m.send_cells = p(self.send_cells, C.c_int32); m.recv_off = p(self.recv_off, C.c_int32)
        m.ib = self.ib.c_struct_ptr() if self.ib is not None else None
        return m


@dataclass
class HostState:
    df: np.ndarray
    sdf: np.ndarray
    flux: np.ndarray
    w: np.ndarray
    prim: np.ndarray
    mflux: np.ndarray
    qf: np.ndarray
    sw: np.ndarray

    @staticmethod
    def zeros(mesh: HostMesh):
        D, K = mesh.dim, mesh.ndf
        npts = int(mesh.vs_off()[-1])
        nc = mesh.n_cell
        z = np.zeros
        return HostState(z(npts * K), z(npts * K * D), z(npts * K), z(nc * (D + 2)), z(nc * (D + 2)),
                         z(nc * (D + 2)), z(nc * D), z(nc * (D + 2) * D))

    def copy(self):
        return HostState(*[getattr(self, f).copy() for f in self.__dataclass_fields__])

    def cell_df(self, mesh, c, off=None):
        off = mesh.vs_off() if off is None else off
        n = int(off[c + 1] - off[c])
        return self.df[off[c] * mesh.ndf: off[c + 1] * mesh.ndf].reshape(mesh.ndf, n)


def pack_grids(grids, dim):
    """Concatenate VGrid objects into grid_off / v_level / v_weight / v_mid (planes per grid)."""
    off = np.zeros(len(grids) + 1, dtype=np.int64)
    for g, gr in enumerate(grids):
        off[g + 1] = off[g] + gr.n
    v_level = np.concatenate([g.level for g in grids]).astype(np.int8)
    v_weight = np.concatenate([g.weight for g in grids]).astype(np.float64)
    v_mid = np.concatenate([np.ascontiguousarray(g.mid.T).ravel() for g in grids]).astype(np.float64)
    return off, v_level, v_weight, v_mid


def build_rank_view(forest: Forest, grids, cell_grid_global, bc_type, bc_prim, ndf,
                    owner=None, rank=0, bound_enc_global=None) -> HostMesh:
    """Flatten one rank's partition of the forest (the re-flatten step of the host shim).

    Faces follow the decision tree of initialize_faces! (src/Solver/Initialize.jl:5-158); since
    p4est's callback order is unavailable the canonical order is: local cells ascending, faceid
    ascending, a local/local full face emitted from the lower-index side.
    """
    D = forest.dim
    N = forest.n
    owner = np.zeros(N, dtype=np.int32) if owner is None else owner
    benc = np.zeros(N, dtype=np.int32) if bound_enc_global is None else bound_enc_global
    local = np.nonzero(owner == rank)[0]
    n_local = len(local)
    loc_of = {int(g): i for i, g in enumerate(local)}
    is_local = owner == rank

    # ghost layer (CONNECT_FULL) and mirrors
    ghost_set = set()
    mirror_of_peer = {}
    multi = (owner != rank).any()
    if multi:
        for g in local:
            for j in forest.adjacent(int(g)):
                if not is_local[j]:
                    ghost_set.add(j)
                    mirror_of_peer.setdefault(int(owner[j]), set()).add(int(g))
    ghosts = sorted(ghost_set, key=lambda j: (int(owner[j]), j))
    n_ghost = len(ghosts)
    for i, g in enumerate(ghosts):
        loc_of[int(g)] = n_local + i
    peers = sorted(mirror_of_peer.keys())
    send_off = [0]; send_cells = []; recv_off = [0]
    for p in peers:
        send_cells += [loc_of[g] for g in sorted(mirror_of_peer[p])]
        send_off.append(len(send_cells))
        recv_off.append(recv_off[-1] + sum(1 for g in ghosts if owner[g] == p))

    gids = np.array(list(local) + ghosts, dtype=np.int64)
    ds = forest.ds[gids].copy(); mid = forest.mid[gids].copy()
    lvl = forest.level[gids].astype(np.int32)
    cg_global = np.asarray(cell_grid_global)
    used = sorted(set(int(x) for x in cg_global[gids]))
    gmap = {g: i for i, g in enumerate(used)}
    cell_grid = np.array([gmap[int(x)] for x in cg_global[gids]], dtype=np.int32)
    grid_off, v_level, v_weight, v_mid = pack_grids([grids[g] for g in used], D)

    nb_state = np.zeros(n_local * 2 * D, dtype=np.int32)
    nb_off = np.zeros(n_local * 2 * D + 1, dtype=np.int32)
    nb_ids = []
    kinds, heres, theres, dirs, rots, fmids, tmids = [], [], [], [], [], [], []
    geo = forest.geometry

    def on_edge(x, d):
        return x == geo[2 * d] or x == geo[2 * d + 1]

    for ci, g in enumerate(local):
        g = int(g)
        for f in range(2 * D):
            state, nbs = forest.face_neighbors(g, f)
            e = ci * 2 * D + f
            nb_state[e] = state
            nb_ids += [loc_of[j] for j in nbs]
            nb_off[e + 1] = len(nb_ids)
            d = f // 2
            rot = 1.0 if f % 2 == 0 else -1.0
            if benc[g] < 0:
                continue  # solid cells never own a face record here (IB flattening adds theirs)
            if state == 0:
                fm = forest.mid[g].copy(); fm[d] -= 0.5 * rot * forest.ds[g][d]
                kinds.append(FACE_DOMAIN); heres.append(ci); theres.append(f); dirs.append(d); rots.append(rot)
                fmids.append(fm); tmids.append(fm.copy())
            elif state == 1:
                j = nbs[0]
                if is_local[j] and j < g and benc[j] >= 0:
                    continue  # emitted from the lower-index side
                if is_local[j] and j == g and f % 2 == 1:
                    continue  # a cell that is its own periodic neighbour: one face per direction
                fm = forest.mid[g].copy(); fm[d] -= 0.5 * rot * forest.ds[g][d]
                tm = forest.mid[j].copy()
                if on_edge(fm[d], d):
                    tm = fm.copy(); tm[d] -= 0.5 * rot * forest.ds[g][d]
                kinds.append(FACE_FULL); heres.append(ci); theres.append(loc_of[j]); dirs.append(d); rots.append(rot)
                fmids.append(fm); tmids.append(tm)
            elif state == -1:
                j = nbs[0]
                if is_local[j] and benc[j] >= 0:
                    continue  # the coarse local cell emits the HangingFace
                fm = forest.mid[g].copy(); fm[d] -= 0.5 * rot * forest.ds[g][d]
                tm = forest.mid[j].copy()
                if on_edge(fm[d], d):
                    tm[d] = fm[d] - rot * forest.ds[g][d]
                kinds.append(FACE_BACKHANGING); heres.append(ci); theres.append(loc_of[j]); dirs.append(d)
                rots.append(rot); fmids.append(fm); tmids.append(tm)
            else:
                for j in nbs:
                    fm = forest.mid[j].copy(); fm[d] = forest.mid[g][d] - 0.5 * rot * forest.ds[g][d]
                    tm = forest.mid[j].copy()
                    if on_edge(fm[d], d):
                        tm = fm.copy(); tm[d] -= 0.5 * rot * forest.ds[j][d]
                    kinds.append(FACE_HANGING); heres.append(ci); theres.append(loc_of[j]); dirs.append(d)
                    rots.append(rot); fmids.append(fm); tmids.append(tm)

    fluid_local = benc[local] >= 0
    minlevel = int(forest.level[benc >= 0].min()) if (benc >= 0).any() else 0
    i32 = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.int32))
    f64 = lambda a, shape=None: np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))
    return HostMesh(
        dim=D, ndf=ndf, n_local=n_local, n_ghost=n_ghost, n_solidnbr=0,
        ds=f64(ds), mid=f64(mid), bound_enc=i32(benc[gids]), ps_level=i32(lvl), cell_grid=cell_grid,
        grid_off=grid_off, v_level=v_level, v_weight=v_weight, v_mid=v_mid,
        nb_state=nb_state, nb_off=nb_off, nb_ids=i32(nb_ids),
        ps_maxlevel=int(forest.maxlevel), ps_minlevel=minlevel,
        face_kind=i32(kinds), face_here=i32(heres), face_there=i32(theres), face_dir=i32(dirs),
        face_rot=f64(rots), face_mid=f64(np.array(fmids).reshape(-1, D) if fmids else np.zeros((0, D))),
        face_there_mid=f64(np.array(tmids).reshape(-1, D) if tmids else np.zeros((0, D))),
        bc_type=i32(bc_type), bc_prim=f64(bc_prim),
        peer_rank=i32(peers), send_off=i32(send_off), send_cells=i32(send_cells), recv_off=i32(recv_off),
        global_ids=gids,
    )
