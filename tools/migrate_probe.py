#!/usr/bin/env python
"""Diagnostic for the device-to-device migration on 2 ranks (torchrun): where does a difference come from —
the partition itself (fresh context on the skewed split), the migration (state compared right after it), or the steps
that follow?"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from kitamr_jl_b200 import abi, api
    from kitamr_jl_b200.synth import cases
    from kitamr_jl_b200.synth.forest import partition
    from oracle import orc
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    case = cases.cylinder_s2(trees=5, ps_maxlevel=4, box_level=2, vtrees=8, vs_maxlevel=2)
    full = case.rank_mesh()
    cfg1 = case.config()
    K, M = full.ndf, case.dim + 2
    off_g = full.vs_off()
    index_of = {int(g): i for i, g in enumerate(full.global_ids[: full.n_local])}
    n_of = np.array([g.n for g in case.grids])[case.cell_grid].astype(np.float64)
    owner_b = partition(n_of * np.linspace(0.4, 1.6, len(n_of)), world)
    mesh_a = case.rank_mesh(rank, world)
    meshes_a = [case.rank_mesh(r, world) for r in range(world)]
    case.owner = lambda nranks, _o=owner_b: _o
    mesh_b = case.rank_mesh(rank, world)
    dt = case.dt()

    def make_ctx(mesh, st):
        ctx = api.Context(case.config(device=local, rank=rank, nranks=world))
        box = [ctx.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ctx.comm_init(box[0])
        ctx.upload_topology(mesh)
        ctx.upload_state(st)
        ctx.exchange_df()
        return ctx

    def err_vs(ref, mesh, out):
        off = mesh.vs_off()
        num = den = 0.0
        worst = (0.0, -1)
        for i in range(mesh.n_local):
            g = index_of[int(mesh.global_ids[i])]
            a, b = out.df[off[i] * K: off[i + 1] * K], ref.df[off_g[g] * K: off_g[g + 1] * K]
            e = float(np.sum((a - b) ** 2)); num += e; den += float(np.sum(b ** 2))
            if e > worst[0]:
                worst = (e, int(mesh.global_ids[i]))
        t = torch.tensor([num, den], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        return float(torch.sqrt(t[0] / t[1])), worst

    # (a) fresh context on the skewed partition
    ref = case.init_state(full)
    ctx = make_ctx(mesh_b, case.init_state(mesh_b))
    for _ in range(8):
        ctx.step(dt, False); orc.step(cfg1, full, ref, dt, False)
    e, w = err_vs(ref, mesh_b, ctx.download_state(case.init_state(mesh_b), abi.DL_DF))
    if rank == 0:
        print(f"(a) fresh context on the skewed partition, 8 steps: rel L2 = {e:.3e}", flush=True)
    ctx.close()
    # (b) partition A, 5 steps, migrate, compare at once; (c) 3 more steps
    ref = case.init_state(full)
    ctx = make_ctx(mesh_a, case.init_state(mesh_a))
    for _ in range(5):
        ctx.step(dt, False); orc.step(cfg1, full, ref, dt, False)
    nl = mesh_a.n_local
    pre = ctx.download_state(case.init_state(mesh_a), abi.DL_DF | abi.DL_W | abi.DL_PRIM)
    off_a = mesh_a.vs_off()
    mine = {int(mesh_a.global_ids[i]): (pre.df[off_a[i] * K: off_a[i + 1] * K].copy(), pre.w[i * M:(i + 1) * M].copy(),
                                         pre.prim[i * M:(i + 1) * M].copy()) for i in range(nl)}
    rows = [None] * world
    dist.all_gather_object(rows, mine)
    by_gid = {}
    for r in rows:
        by_gid.update(r)
    new_id = {int(g): i for i, g in enumerate(mesh_b.global_ids[: mesh_b.n_local])}
    dest = owner_b[mesh_a.global_ids[:nl]]
    src_rank, src_cells, src_points, recv_cells = [], [], [], []
    nb_of = mesh_b.cell_n()
    for r in range(world):
        gids = meshes_a[r].global_ids[: meshes_a[r].n_local]
        lst = [new_id[int(g)] for g in gids if owner_b[g] == rank]
        if lst:
            src_rank.append(r); src_cells.append(len(lst)); src_points.append(int(nb_of[lst].sum())); recv_cells += lst
    # mid-run state laid out for partition B (host side), for the variants below
    st_mid = case.init_state(mesh_b)
    off_b0 = mesh_b.vs_off()
    for i in range(mesh_b.n_local):
        d, w_, p_ = by_gid[int(mesh_b.global_ids[i])]
        st_mid.df[off_b0[i] * K: off_b0[i + 1] * K] = d
        st_mid.w[i * M:(i + 1) * M] = w_; st_mid.prim[i * M:(i + 1) * M] = p_
    ref6 = ref.copy()
    orc.step(cfg1, full, ref6, dt, False)
    # (d) fresh context on B, mid-run state uploaded, one step
    c2 = make_ctx(mesh_b, st_mid)
    c2.step(dt, False)
    e, _ = err_vs(ref6, mesh_b, c2.download_state(case.init_state(mesh_b), abi.DL_DF))
    if rank == 0:
        print(f"(d) fresh context on B + upload_state(mid-run state) + 1 step: rel L2 = {e:.3e}", flush=True)
    c2.close()
    # (e) old context: re-flatten to B, upload_state(mid-run state), one step  (no migration calls)
    c3 = make_ctx(mesh_a, case.init_state(mesh_a))
    for _ in range(5):
        c3.step(dt, False)
    c3.upload_topology(mesh_b)
    c3.upload_state(st_mid)
    c3.exchange_df()
    c3.step(dt, False)
    e, _ = err_vs(ref6, mesh_b, c3.download_state(case.init_state(mesh_b), abi.DL_DF))
    if rank == 0:
        print(f"(e) 5 steps on A, re-flatten to B, upload_state(mid-run state), 1 step: rel L2 = {e:.3e}", flush=True)
    # (e2) the same context: re-flatten to B again, upload again, kamr_slope first, then the step
    c3.upload_topology(mesh_b)
    c3.upload_state(st_mid)
    c3.exchange_df()
    c3.slope()
    c3.step(dt, False)
    e, _ = err_vs(ref6, mesh_b, c3.download_state(case.init_state(mesh_b), abi.DL_DF))
    if rank == 0:
        print(f"(e2) ... with kamr_slope before the step: rel L2 = {e:.3e}", flush=True)
    c3.close()
    ctx.migrate_begin(np.arange(nl), dest, src_rank, src_cells, src_points)
    ctx.upload_topology(mesh_b)
    ctx.migrate_finish(recv_cells)
    ctx.exchange_df()
    post = ctx.download_state(case.init_state(mesh_b), abi.DL_DF | abi.DL_W | abi.DL_PRIM)
    off_b = mesh_b.vs_off()
    bad = [0, 0, 0]
    for i in range(mesh_b.n_local):
        d, w_, p_ = by_gid[int(mesh_b.global_ids[i])]
        bad[0] += int(not np.array_equal(post.df[off_b[i] * K: off_b[i + 1] * K], d))
        bad[1] += int(not np.array_equal(post.w[i * M:(i + 1) * M], w_))
        bad[2] += int(not np.array_equal(post.prim[i * M:(i + 1) * M], p_))
    t = torch.tensor(bad, dtype=torch.float64, device="cuda"); dist.all_reduce(t)
    e0, _ = err_vs(ref, mesh_b, post)
    if rank == 0:
        print(f"(b) right after the migration: cells whose df / w / prim differ from the pre-migration device state: "
              f"{[int(x) for x in t]}; rel L2 vs oracle = {e0:.3e}", flush=True)
    for k in range(3):
        ctx.step(dt, False); orc.step(cfg1, full, ref, dt, False)
        e, w = err_vs(ref, mesh_b, ctx.download_state(case.init_state(mesh_b), abi.DL_DF))
        lw = [None] * world
        dist.all_gather_object(lw, w)
        if rank == 0:
            print(f"(c) {k + 1} step(s) after: rel L2 = {e:.3e}; worst cells (err^2, gid, old owner -> new owner): "
                  f"{[(f'{x[0]:.2e}', x[1], int(case.__class__.owner(case, world)[x[1]]) if False else None) for x in lw]}",
                  flush=True)
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
