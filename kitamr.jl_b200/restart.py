"""The reference's restart payload (src/IO/Restart.jl:79-118,192-199) <-> the flat host state of the C-ABI.

`save_for_restart` writes one `restart_<rank>.jld2` per rank with six arrays, in Julia's column-major layout over ALL
quadrants of the rank in p4est order (InsideSolidData placeholders included, with vs_num = 0):

    vs_nums      Int64   [N]            velocity points per quadrant
    bound_encs   Int64   [N]            PsData.bound_enc
    ws           Float64 [N, DIM+2]     conserved variables                       (column-major: DIM+2 columns of N)
    vs_levels    Int8    [Nv]           velocity refinement level of every point, quadrants back to back
    vs_midpoints Float64 [Nv, DIM]      velocity coordinates                      (column-major: DIM columns of Nv)
    vs_df        Float64 [Nv, NDF]      distribution function                     (column-major: NDF columns of Nv)

The C-ABI (include/kamr.h) keeps the same data per CELL: a cell's df block is NDF planes of its own n points, grids are
per-cell planes.  This module converts both ways, so a restart file of the real reference is a ready-made fixture
 and a device state can be written back as a payload the reference restarts from.

The arrays are exchanged as a plain dict / `.npz` (h5py is not in this image; JLD2 files are HDF5 — dumping the six
datasets with `h5py.File(...)[name][...]` or `Kamr.dump_reference_step` is all a Julia host needs to do).
"""
from __future__ import annotations

import numpy as np

PAYLOAD_KEYS = ("vs_nums", "bound_encs", "ws", "vs_levels", "vs_midpoints", "vs_df")


def payload_from_state(mesh, state, weights_out=False):
    """HostMesh + HostState (local cells) -> the six payload arrays (numpy, Fortran-ordered like Julia's)."""
    D, K, M = mesh.dim, mesh.ndf, mesh.dim + 2
    nl = mesh.n_local
    n = mesh.cell_n()[:nl].astype(np.int64)
    off = mesh.vs_off()
    Nv = int(off[nl])
    ws = np.asfortranarray(state.w[: nl * M].reshape(nl, M))
    vs_levels = np.empty(Nv, dtype=np.int8)
    vs_mid = np.empty((Nv, D), order="F")
    vs_df = np.empty((Nv, K), order="F")
    go = mesh.grid_off
    for c in range(nl):
        g = int(mesh.cell_grid[c]); nc = int(n[c]); a, b = int(off[c]), int(off[c + 1])
        vs_levels[a:b] = mesh.v_level[go[g]: go[g + 1]]
        vs_mid[a:b, :] = mesh.v_mid[go[g] * D: go[g + 1] * D].reshape(D, nc).T
        vs_df[a:b, :] = state.df[a * K: b * K].reshape(K, nc).T
    return {"vs_nums": n.copy(), "bound_encs": mesh.bound_enc[:nl].astype(np.int64), "ws": ws, "vs_levels": vs_levels,
            "vs_midpoints": vs_mid, "vs_df": vs_df}


def state_from_payload(p, root_weight=None):
    """The six payload arrays -> per-cell ABI arrays of the quadrants that own velocity data:
    dict(keep, vs_off, df, w, bound_enc, cell_grid, grid_off, v_level, v_mid[, v_weight]).  `keep` are the indices of
    the quadrants with vs_num > 0 (InsideSolidData placeholders are dropped, as the flattener drops them).  Identical
    velocity grids are stored once.  `root_weight`: the weight of a level-0 velocity cell (velocity-domain volume /
    number of roots); with it v_weight = root_weight / 2^(DIM level), as the reference rebuilds it after a migration
    (Parallel/Partition.jl:624-644)."""
    vs_nums = np.asarray(p["vs_nums"], dtype=np.int64)
    ws = np.asarray(p["ws"], dtype=np.float64)
    lev = np.asarray(p["vs_levels"], dtype=np.int8)
    mid = np.asarray(p["vs_midpoints"], dtype=np.float64)
    vdf = np.asarray(p["vs_df"], dtype=np.float64)
    D, K = mid.shape[1], vdf.shape[1]
    assert ws.shape == (len(vs_nums), D + 2) and len(lev) == mid.shape[0] == vdf.shape[0] == int(vs_nums.sum())
    keep = np.nonzero(vs_nums > 0)[0]
    n = vs_nums[keep]
    src_off = np.concatenate([[0], np.cumsum(vs_nums)])[keep]          # placeholders take no points
    vs_off = np.concatenate([[0], np.cumsum(n)]).astype(np.int64)
    df = np.empty(int(vs_off[-1]) * K)
    grids, cell_grid = {}, np.zeros(len(keep), dtype=np.int32)
    g_level, g_mid, grid_off = [], [], [0]
    for c, (a, nc) in enumerate(zip(src_off, n)):
        a, nc = int(a), int(nc)
        df[vs_off[c] * K: vs_off[c + 1] * K] = vdf[a:a + nc, :].T.ravel()
        key = (lev[a:a + nc].tobytes(), mid[a:a + nc, :].tobytes())
        g = grids.get(key)
        if g is None:
            g = grids[key] = len(grid_off) - 1
            g_level.append(lev[a:a + nc].copy()); g_mid.append(np.ascontiguousarray(mid[a:a + nc, :].T).ravel())
            grid_off.append(grid_off[-1] + nc)
        cell_grid[c] = g
    out = {"keep": keep, "vs_off": vs_off, "df": df, "w": np.ascontiguousarray(ws[keep]).ravel(),
           "bound_enc": np.asarray(p["bound_encs"], dtype=np.int64)[keep].astype(np.int32), "cell_grid": cell_grid,
           "grid_off": np.asarray(grid_off, dtype=np.int64), "v_level": np.concatenate(g_level).astype(np.int8),
           "v_mid": np.concatenate(g_mid), "dim": D, "ndf": K}
    if root_weight is not None:
        out["v_weight"] = root_weight / 2.0 ** (D * out["v_level"].astype(np.float64))
    return out


def save_npz(path, payload):
    np.savez_compressed(path, **{k: np.asarray(payload[k]) for k in PAYLOAD_KEYS})


def load_npz(path):
    z = np.load(path)
    return {k: z[k] for k in PAYLOAD_KEYS}


# ---------------------------------------------------------------------------------------------------------------
def read_reference_dump(path):
    """Reads a directory written by `Kamr.dump_reference_step` (kitamr.jl_b200/julia/Kamr.jl): the flat mesh and state
    of a REAL KitAMR run before, and df / w / prim after `steps` reference steps.  Returns (HostMesh, HostState, after,
    meta); feeding (mesh, state) to the oracle and comparing with `after` pins the oracle against the reference."""
    import os
    from .model import HostMesh, HostState
    from .synth import ib as ibm
    meta, arrays = {}, {}
    jl = {"Float64": np.float64, "Int32": np.int32, "Int64": np.int64, "Int8": np.int8}
    for line in open(os.path.join(path, "index.txt")):
        t = line.split()
        if not t:
            continue
        if t[0] == "array":
            a = np.fromfile(os.path.join(path, t[1] + ".bin"), dtype=jl[t[2]])
            assert len(a) == int(t[3]), t
            arrays[t[1]] = a
        else:
            meta[t[0]] = float(t[1]) if "." in t[1] or "e" in t[1] else int(t[1])
    A = arrays
    D, K = meta["dim"], meta["ndf"]
    hib = None
    if len(A["sn_donor"]) or len(A["solid_cell"]):
        hib = ibm.HostIB(A["solid_cell"], A["solid_nb_off"], A["solid_nb_ids"], A["sn_donor"], A["sn_solid"], A["sn_faceid"],
                         A["sn_aux"], A["sn_normal"], A["sn_bc"], A["sn_nb_off"], A["sn_nb_ids"], A["cvc_off"], A["cvc_index"],
                         A["cvc_gas_w"], A["cvc_solid_w"])
    mesh = HostMesh(dim=D, ndf=K, n_local=meta["n_local"], n_ghost=meta["n_ghost"], n_solidnbr=meta["n_solidnbr"],
                    ds=A["ds"], mid=A["mid"], bound_enc=A["bound_enc"], ps_level=A["ps_level"], cell_grid=A["cell_grid"],
                    grid_off=A["grid_off"], v_level=A["v_level"], v_weight=A["v_weight"], v_mid=A["v_mid"],
                    nb_state=A["nb_state"], nb_off=A["nb_off"], nb_ids=A["nb_ids"], ps_maxlevel=meta["ps_maxlevel"],
                    ps_minlevel=meta["ps_minlevel"], face_kind=A["face_kind"], face_here=A["face_here"],
                    face_there=A["face_there"], face_dir=A["face_dir"], face_rot=A["face_rot"], face_mid=A["face_mid"],
                    face_there_mid=A["face_there_mid"], bc_type=A["bc_type"], bc_prim=A["bc_prim"],
                    peer_rank=A["peer_rank"], send_off=A["send_off"], send_cells=A["send_cells"], recv_off=A["recv_off"],
                    global_ids=np.arange(meta["n_local"] + meta["n_ghost"], dtype=np.int64), ib=hib)
    st = HostState.zeros(mesh)
    st.df[:] = A["df"]; st.w[:] = A["w"]; st.prim[:] = A["prim"]
    after = {k: np.fromfile(os.path.join(path, f"after_{k}.bin"), dtype=np.float64) for k in ("df", "w", "prim")}
    return mesh, st, after, meta
