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
                    owner=None, rank=0, bound_enc_global=None, cell_class=None, ib_shape=None) -> HostMesh:
    """Flatten one rank's partition of the forest (the re-flatten step of the host shim).

    Faces follow the decision tree of initialize_faces! (src/Solver/Initialize.jl:5-158); since
    p4est's callback order is unavailable the canonical order is: local cells ascending, faceid
    ascending, a local/local full face emitted from the lower-index side.

    Immersed boundary (cell_class + ib_shape): INSIDE_SOLID cells are not listed (InsideSolidData placeholders,
    Solver/Initialize.jl:358-376); SOLID_GHOST cells are listed with bound_enc = -1; a fluid cell with a solid ghost
    face-neighbour becomes a donor (bound_enc = +1) and that neighbour slot / face `there` is replaced by a
    SolidNeighbor pseudo-cell (Boundary/Immersed_boundary.jl:282-305) appended after the ghosts.
    """
    from .synth import ib as ibm
    D = forest.dim
    N = forest.n
    owner = np.zeros(N, dtype=np.int32) if owner is None else owner
    if cell_class is not None:
        benc = np.where(cell_class == ibm.SOLID_GHOST, -1, 0).astype(np.int32)
        dropped = cell_class == ibm.INSIDE_SOLID
    else:
        benc = np.zeros(N, dtype=np.int32) if bound_enc_global is None else np.asarray(bound_enc_global).copy()
        dropped = np.zeros(N, dtype=bool)
    if ib_shape is not None:   # donors: fluid cells with a solid ghost face neighbour (Immersed_boundary.jl:285-287)
        for g in np.nonzero(benc < 0)[0]:
            for f in range(2 * D):
                state, nbs = forest.face_neighbors(int(g), f)
                for j in nbs:
                    if benc[j] >= 0 and not dropped[j]:
                        if state != 1:
                            raise RuntimeError("immersed boundary crosses a refinement-level jump")
                        benc[j] = 1
    local = np.nonzero((owner == rank) & ~dropped)[0]
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
                if not is_local[j] and not dropped[j]:
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
    benc_l = benc[gids].astype(np.int32)

    nb_state = np.zeros(n_local * 2 * D, dtype=np.int32)
    nb_off = np.zeros(n_local * 2 * D + 1, dtype=np.int32)
    nb_ids = []
    kinds, heres, theres, dirs, rots, fmids, tmids = [], [], [], [], [], [], []
    geo = forest.geometry
    # SolidNeighbor records
    sn = dict(donor=[], solid=[], faceid=[], aux=[], normal=[], bc=[], nb=[], ds=[], mid=[], grid=[], lvl=[])
    n_lg = n_local + n_ghost

    def on_edge(x, d):
        return x == geo[2 * d] or x == geo[2 * d + 1]

    for ci, g in enumerate(local):
        g = int(g)
        for f in range(2 * D):
            state, nbs = forest.face_neighbors(g, f)
            if any(dropped[j] for j in nbs):
                if benc[g] >= 0:
                    raise RuntimeError("fluid cell next to an inside-solid cell")
                state, nbs = 0, []   # a solid ghost cell looking into the body: InsideSolidData, never read
            e = ci * 2 * D + f
            nb_state[e] = state
            d = f // 2
            rot = 1.0 if f % 2 == 0 else -1.0
            if benc[g] < 0:
                nb_ids += [loc_of[j] for j in nbs]
                nb_off[e + 1] = len(nb_ids)
                continue  # solid cells never own a face record: the fluid side emits it (Initialize.jl:67-71,92-105)
            if state == 1 and benc[nbs[0]] < 0 and ib_shape is not None:
                j = nbs[0]
                sid = n_lg + len(sn["donor"])
                aux, normal = ib_shape.calc_intersect(forest.mid[g], forest.mid[j])
                sn["donor"].append(ci); sn["solid"].append(loc_of[j]); sn["faceid"].append(f)
                sn["aux"].append(aux); sn["normal"].append(normal); sn["bc"].append(np.asarray(ib_shape.bc, float))
                sn["ds"].append(forest.ds[j]); sn["mid"].append(forest.mid[j]); sn["grid"].append(cell_grid[ci])
                sn["lvl"].append(int(forest.level[j]))
                nb_ids.append(sid)
                nb_off[e + 1] = len(nb_ids)
                fm = forest.mid[g].copy(); fm[d] -= 0.5 * rot * forest.ds[g][d]
                kinds.append(FACE_FULL); heres.append(ci); theres.append(sid); dirs.append(d); rots.append(rot)
                fmids.append(fm); tmids.append(forest.mid[j].copy())
                continue
            nb_ids += [loc_of[j] for j in nbs]
            nb_off[e + 1] = len(nb_ids)
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

    n_sn = len(sn["donor"])
    host_ib = None
    if ib_shape is not None:
        # fluid neighbours of solid ghost cells: first element of every face / corner list (Immersed_boundary.jl:207-213)
        solid_cells, s_off, s_ids = [], [0], []
        for ci, g in enumerate(local):
            if benc[g] >= 0:
                continue
            fl = [j for j in forest.direction_neighbors(int(g)) if j is not None and not dropped[j] and benc[j] >= 0]
            if not fl:
                continue   # no fluid neighbour in reach (only possible on coarse test meshes): nothing reads this cell
            solid_cells.append(ci)
            s_ids += [loc_of[j] for j in fl]
            s_off.append(len(s_ids))
        # donor's fluid face neighbours (image_df, Immersed_boundary.jl:366-373); the donor itself goes last
        n_off, n_ids = [0], []
        c_off, c_idx, c_gw, c_sw = [0], [], [], []
        for s in range(n_sn):
            ci = sn["donor"][s]; g = int(local[ci])
            for f in range(2 * D):
                state, nbs = forest.face_neighbors(g, f)
                if nbs and not dropped[nbs[0]] and benc[nbs[0]] >= 0:
                    n_ids.append(loc_of[nbs[0]])
            n_off.append(len(n_ids))
            idx, gw, sw = ibm.cut_cells(sn["normal"][s], grids[int(cg_global[g])])
            c_idx += list(idx); c_gw += list(gw); c_sw += list(sw)
            c_off.append(len(c_idx))
        i32a = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.int32))
        f64a = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))
        host_ib = ibm.HostIB(i32a(solid_cells), i32a(s_off), i32a(s_ids), i32a(sn["donor"]), i32a(sn["solid"]),
                             i32a(sn["faceid"]), f64a(sn["aux"]), f64a(sn["normal"]), f64a(sn["bc"]), i32a(n_off),
                             i32a(n_ids), i32a(c_off), i32a(c_idx), f64a(c_gw), f64a(c_sw))
        if n_sn:
            ds = np.concatenate([ds, np.array(sn["ds"])]); mid = np.concatenate([mid, np.array(sn["mid"])])
            lvl = np.concatenate([lvl, np.array(sn["lvl"], dtype=np.int32)])
            cell_grid = np.concatenate([cell_grid, np.array(sn["grid"], dtype=np.int32)])
            benc_l = np.concatenate([benc_l, np.full(n_sn, -1, dtype=np.int32)])

    fluid_sel = (benc >= 0) & ~dropped
    minlevel = int(forest.level[fluid_sel].min()) if fluid_sel.any() else 0
    i32 = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.int32))
    f64 = lambda a, shape=None: np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))
    return HostMesh(
        dim=D, ndf=ndf, n_local=n_local, n_ghost=n_ghost, n_solidnbr=n_sn,
        ds=f64(ds), mid=f64(mid), bound_enc=i32(benc_l), ps_level=i32(lvl), cell_grid=i32(cell_grid),
        grid_off=grid_off, v_level=v_level, v_weight=v_weight, v_mid=v_mid,
        nb_state=nb_state, nb_off=nb_off, nb_ids=i32(nb_ids),
        ps_maxlevel=int(forest.maxlevel), ps_minlevel=minlevel,
        face_kind=i32(kinds), face_here=i32(heres), face_there=i32(theres), face_dir=i32(dirs),
        face_rot=f64(rots), face_mid=f64(np.array(fmids).reshape(-1, D) if fmids else np.zeros((0, D))),
        face_there_mid=f64(np.array(tmids).reshape(-1, D) if tmids else np.zeros((0, D))),
        bc_type=i32(bc_type), bc_prim=f64(bc_prim),
        peer_rank=i32(peers), send_off=i32(send_off), send_cells=i32(send_cells), recv_off=i32(recv_off),
        global_ids=gids, ib=host_ib,
    )


def uniquify_grids(mesh: HostMesh) -> HostMesh:
    """The same mesh with one velocity-grid copy per (local or ghost) cell, as a host holds them when every PsData
    owns its VsData (Velocity_space/Types.jl:5-20) and nothing deduplicates: `cell_grid` becomes the identity.  A
    SolidNeighbor keeps sharing its donor's grid (Boundary/Immersed_boundary.jl:290-298).  The library recognises
    identical grids by content, so index maps are unchanged; only the statics stop being cache-resident."""
    import dataclasses
    D = mesh.dim
    nlg = mesh.n_local + mesh.n_ghost
    go = mesh.grid_off
    g = mesh.cell_grid[:nlg].astype(np.int64)
    n = (go[1:] - go[:-1])[g]
    new_off = np.concatenate([[0], np.cumsum(n)]).astype(np.int64)
    total = int(new_off[-1])
    within = np.arange(total, dtype=np.int64) - np.repeat(new_off[:-1], n)
    src = np.repeat(go[g], n) + within
    v_level = mesh.v_level[src]
    v_weight = mesh.v_weight[src]
    v_mid = np.empty(total * D, dtype=np.float64)
    base_new = np.repeat(new_off[:-1] * D, n)
    base_old = np.repeat(go[g] * D, n)
    nn = np.repeat(n, n)
    for d in range(D):
        v_mid[base_new + d * nn + within] = mesh.v_mid[base_old + d * nn + within]
    cell_grid = np.arange(mesh.n_cell, dtype=np.int32)
    if mesh.n_solidnbr:
        cell_grid[nlg:] = np.asarray(mesh.ib.sn_donor, dtype=np.int32)
    return dataclasses.replace(mesh, cell_grid=np.ascontiguousarray(cell_grid), grid_off=new_off,
                               v_level=np.ascontiguousarray(v_level), v_weight=np.ascontiguousarray(v_weight),
                               v_mid=v_mid)
