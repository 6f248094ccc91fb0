"""Host-side reference of the halo exchange (test infrastructure): packs mirror cells' blocks in
mirror_proc_mirrors order and lands them in the ghost tail, exactly the index maps libkamr builds its NCCL
send/recv plan from (src/Parallel/Ghost.jl:133-145, 203-284, 757-808, 896).  Transport: torch.distributed (gloo)."""
import numpy as np
import torch
import torch.distributed as dist


def _cells_block(mesh, arr, comps, cells):
    off = mesh.vs_off()
    return np.concatenate([arr[off[c] * comps: off[c + 1] * comps] for c in cells]) if len(cells) else np.zeros(0)


def exchange(mesh, arr, comps, level=None, solid_only=False):
    """Exchange per-point blocks (`comps` planes per cell) of mirror cells -> ghost cells.  level: only cells of
    that physical level and not solid (slope_exchange_level!), None: all (data_exchange!)."""
    off = mesh.vs_off()
    reqs, recvs = [], []
    for p, peer in enumerate(mesh.peer_rank):
        send = [int(c) for c in mesh.send_cells[mesh.send_off[p]: mesh.send_off[p + 1]]]
        ghosts = list(range(mesh.n_local + int(mesh.recv_off[p]), mesh.n_local + int(mesh.recv_off[p + 1])))
        if level is not None:
            send = [c for c in send if mesh.ps_level[c] == level and mesh.bound_enc[c] >= 0]
            ghosts = [c for c in ghosts if mesh.ps_level[c] == level and mesh.bound_enc[c] >= 0]
        if solid_only:   # solid_exchange_begin!/finish!, Boundary/Parallel.jl:138-259
            send = [c for c in send if mesh.bound_enc[c] < 0]
            ghosts = [c for c in ghosts if mesh.bound_enc[c] < 0]
        sb = torch.from_numpy(_cells_block(mesh, arr, comps, send))
        rb = torch.zeros(int(sum(off[c + 1] - off[c] for c in ghosts)) * comps, dtype=torch.float64)
        if sb.numel():
            reqs.append(dist.isend(sb, int(peer)))
        if rb.numel():
            reqs.append(dist.irecv(rb, int(peer)))
        recvs.append((ghosts, rb))
    for r in reqs:
        r.wait()
    for ghosts, rb in recvs:
        pos = 0
        rbn = rb.numpy()
        for c in ghosts:
            ln = int(off[c + 1] - off[c]) * comps
            arr[off[c] * comps: off[c] * comps + ln] = rbn[pos: pos + ln]
            pos += ln


def exchange_cells(mesh, arr, width):
    """per-cell arrays (w: DIM+2 per cell) of mirrors -> ghosts"""
    reqs, recvs = [], []
    for p, peer in enumerate(mesh.peer_rank):
        send = mesh.send_cells[mesh.send_off[p]: mesh.send_off[p + 1]]
        g0, g1 = mesh.n_local + int(mesh.recv_off[p]), mesh.n_local + int(mesh.recv_off[p + 1])
        sb = torch.from_numpy(np.ascontiguousarray(arr.reshape(-1, width)[send]).ravel())
        rb = torch.zeros((g1 - g0) * width, dtype=torch.float64)
        reqs.append(dist.isend(sb, int(peer)))
        reqs.append(dist.irecv(rb, int(peer)))
        recvs.append((g0, g1, rb))
    for r in reqs:
        r.wait()
    for g0, g1, rb in recvs:
        arr.reshape(-1, width)[g0:g1] = rb.numpy().reshape(-1, width)


def oracle_step_distributed(orc, cfg, mesh, st, dt, want_residual=False):
    """slope! / flux! / iterate! with the reference's exchange points (Slope.jl:1055-1068, Iterate.jl:8-9)."""
    D, K = mesh.dim, mesh.ndf
    orc.slope_level(cfg, mesh, st, mesh.ps_minlevel, 0)
    exchange(mesh, st.sdf, K * D, level=mesh.ps_minlevel)
    for L in range(mesh.ps_minlevel + 1, mesh.ps_maxlevel + 1):
        orc.slope_level(cfg, mesh, st, L, 1)
        exchange(mesh, st.sdf, K * D, level=L)
    orc.macro_slope(cfg, mesh, st)
    orc.ib_solid_cells(cfg, mesh, st)
    exchange(mesh, st.df, K, solid_only=True)
    orc.ib_solid_neighbors(cfg, mesh, st)
    orc.flux(cfg, mesh, st, dt)
    res = orc.iterate(cfg, mesh, st, dt, want_residual)
    exchange(mesh, st.df, K)
    exchange_cells(mesh, st.w, D + 2)
    return res
